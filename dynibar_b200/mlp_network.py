"""Parameter containers for the three networks on the hot path.

Mirror of the reference's constructor surface and `state_dict` key names
(ibrnet/mlp_network.py:129-234 DynibarDynamic, :319-421 DynibarStatic,
:558-603 MotionMLP) so published checkpoints load with `load_state_dict` and
the weight packer sees the names SURVEY.md 8(b) lists.  There is no PyTorch
math here: `forward` hands the parameters to the CUDA library
(`dynibar_b200.render_ray`), and raises if it is missing.
"""

import torch
import torch.nn as nn


def _seq(*dims_and_acts):
  """Build nn.Sequential with Linear layers at the reference's indices; the
  activation slots hold parameter-free placeholders (indices must match the
  reference's Sequential so `base_fc.2.weight` etc. line up)."""
  layers = []
  for item in dims_and_acts:
    if isinstance(item, tuple):
      layers.append(nn.Linear(item[0], item[1]))
    else:
      layers.append(nn.Identity())  # activation slot (applied inside the kernels)
  return nn.Sequential(*layers)


class _PackedNet(nn.Module):
  """Base class of the containers (packing/caching lives in weights.py)."""

  kind = None  # 'dynamic' | 'static' | 'motion'


class DynibarDynamic(_PackedNet):
  """Time-varying model container (ibrnet/mlp_network.py:129-234)."""

  kind = "dynamic"

  def __init__(self, args, in_feat_ch=32, n_samples=64, shift=0.0, **kwargs):
    super().__init__()
    if not args.input_dir:
      raise NotImplementedError  # mlp_network.py:215-216
    self.args = args
    self.anti_alias_pooling = False  # hard-coded off, mlp_network.py:135
    self.input_dir = args.input_dir
    self.shift = shift
    self.n_samples = n_samples
    self.in_feat_ch = in_feat_ch
    F = in_feat_ch + 3
    self.ray_dir_fc = _seq((21, 256), "a", (256, F), "a")
    self.base_fc = _seq((F * 3, 256), "a", (256, 128), "a")
    self.vis_fc = _seq((128, 128), "a", (128, 129), "a")
    self.vis_fc2 = _seq((128, 128), "a", (128, 1), "a")
    self.geometry_fc = _seq((257, 256), "a", (256, 128), "a")
    self.ray_attention = MultiHeadAttention(4, 128, 32, 32)
    self.ref_pts_fc = _seq((33 + 128, 256), "a", (256, 128), "a")
    self.out_geometry_fc = _seq((128, 128), "a", (128, 1))
    self.rgb_fc = _seq((128 + 27, 128), "a", (128, 64), "a", (64, 3), "a")

  def forward(self, pts_xyz, rgb_feat, glb_ray_dir, ray_diff, time_diff, mask,
              time):
    """Same signature as mlp_network.py:236-238 -> raw [R,S,4]."""
    from dynibar_b200 import render_ray as rr
    return rr.net_dynamic_forward(self, pts_xyz, rgb_feat, glb_ray_dir, mask, time)


class DynibarStatic(_PackedNet):
  """Time-invariant model container (ibrnet/mlp_network.py:319-421)."""

  kind = "static"

  def __init__(self, args, in_feat_ch=32, n_samples=64, **kwargs):
    super().__init__()
    if not args.input_dir:
      raise NotImplementedError("input_dir=False head (mlp_network.py:396-403) "
                                "is not on the benchmarked path")
    self.args = args
    self.anti_alias_pooling = args.anti_alias_pooling
    self.mask_rgb = args.mask_rgb
    self.input_dir = args.input_dir
    self.n_samples = n_samples
    self.in_feat_ch = in_feat_ch
    F = in_feat_ch + 3
    if self.anti_alias_pooling:
      self.s = nn.Parameter(torch.tensor(0.2), requires_grad=True)
    self.ray_dir_fc = _seq((4 + 33 + 66, 256), "a", (256, F))
    self.ref_feature_fc = _seq((66, F))
    self.base_fc = _seq((F * 6, 256), "a", (256, 128), "a")
    self.vis_fc = _seq((128, 128), "a", (128, 129), "a")
    self.vis_fc2 = _seq((128, 128), "a", (128, 1), "a")
    self.geometry_fc = _seq((257, 256), "a", (256, 128), "a")
    self.ray_attention = MultiHeadAttention(4, 128, 32, 32)
    self.out_geometry_fc = _seq((128, 128), "a", (128, 1))
    self.rgb_fc = _seq((128 * 2 + 1 + 4, 128), "a", (128, 64), "a", (64, 1))

  def forward(self, pts, ref_rays_coords, src_rays_coords, rgb_feat,
              glb_ray_dir, ray_diff, mask):
    """Same signature as mlp_network.py:423-432 -> raw [R,S,4]."""
    from dynibar_b200 import render_ray as rr
    return rr.net_static_forward(self, pts, ref_rays_coords, src_rays_coords,
                                 rgb_feat, ray_diff, mask)


class MultiHeadAttention(nn.Module):
  """Ray-transformer parameters (ibrnet/mlp_network.py:56-74)."""

  def __init__(self, n_head, d_model, d_k, d_v, dropout=0.1):
    super().__init__()
    self.n_head, self.d_k, self.d_v = n_head, d_k, d_v
    self.w_qs = nn.Linear(d_model, n_head * d_k, bias=False)
    self.w_ks = nn.Linear(d_model, n_head * d_k, bias=False)
    self.w_vs = nn.Linear(d_model, n_head * d_v, bias=False)
    self.fc = nn.Linear(n_head * d_v, d_model, bias=False)
    self.layer_norm = nn.LayerNorm(d_model, eps=1e-6)


class MotionMLP(_PackedNet):
  """Motion-trajectory MLP container (ibrnet/mlp_network.py:558-603)."""

  kind = "motion"

  def __init__(self, num_basis=4, D=8, W=256, input_ch=4, num_freqs=16,
               skips=(4,), sf_mag_div=1.0):
    super().__init__()
    if (D, W, input_ch, num_freqs, tuple(skips)) != (8, 256, 4, 16, (4,)):
      raise NotImplementedError("kernels are specialised to the shipped "
                                "MotionMLP shape (8x256, skip 4, 16 freqs)")
    self.D, self.W = D, W
    self.num_basis = num_basis
    self.input_ch = input_ch + input_ch * num_freqs * 2
    self.skips = list(skips)
    self.sf_mag_div = sf_mag_div
    self.pts_linears = nn.ModuleList(
        [nn.Linear(self.input_ch, W)]
        + [nn.Linear(W + self.input_ch, W) if i in self.skips else nn.Linear(W, W)
           for i in range(D - 1)])
    self.coeff_linear = nn.Linear(W, num_basis * 3)
    self.coeff_linear.weight.data.fill_(0.0)  # mlp_network.py:602-603
    self.coeff_linear.bias.data.fill_(0.0)

  def forward(self, x):
    """x [...,4] xyzt -> coeffs [...,3*num_basis] (mlp_network.py:605-618)."""
    from dynibar_b200 import render_ray as rr
    return rr.motion_mlp_forward(self, x)
