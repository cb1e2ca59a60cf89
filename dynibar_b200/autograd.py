"""First slice of the training backward (row f2): `torch.autograd.Function`s over the CUDA forward/backward
kernels of the two non-MLP ends of the path.

  * `composite`         raw2outputs (ibrnet/render_ray.py:214-330): gradients of rgb / rgb_static / rgb_dy / depth /
                        alpha_dy / weights_dy / weights_st / alpha / weights w.r.t. raw_dy, raw_st.
  * `project_gather`    Projector.compute_with_motions (ibrnet/projection.py:103-176): gradient of rgb_feat w.r.t.
                        the source feature maps and the motion-displaced points (ray_diff and mask are detached /
                        non-differentiable, as in the reference).

  * `motion_mlp`        MotionMLP.forward (ibrnet/mlp_network.py:605-618): gradients w.r.t. every parameter of the
                        module and w.r.t. the xyzt rows (fp32 GEMMs, csrc/motion_train.cu).

The backward of the aggregation nets (per-view stage, ray transformer) is not built yet, so the orchestrators (`render_rays_*`) still refuse inputs
that require grad; these functions are the tested building blocks of that step (tests/test_backward_gpu.py checks
them against torch autograd through the oracle's restatement of the same functions).
"""

from collections import OrderedDict

import torch

from dynibar_b200._lib import lib, ptr, f32c, check, stream, dev_of, Args


class _Composite(torch.autograd.Function):
  @staticmethod
  def forward(ctx, raw_dy, raw_st, z_vals, mask_dy, mask_st, min_dy, min_st):
    R, S = z_vals.shape
    dev = dev_of(z_vals)
    V_dy, V_st = mask_dy.shape[2], mask_st.shape[2]
    rays = torch.empty(R, 11, device=dev)
    samp = torch.empty(5, R, S, device=dev)
    rd, rs, zz = f32c(raw_dy), f32c(raw_st), f32c(z_vals)
    with torch.cuda.device(dev):
      check(lib.dyn_composite(ptr(rd), ptr(rs), ptr(zz), ptr(f32c(mask_dy)), V_dy, int(min_dy),
                              ptr(f32c(mask_st)), V_st, int(min_st), R, S, ptr(rays), ptr(samp), stream()))
    ctx.save_for_backward(rd, rs, zz)
    ctx.mark_non_differentiable(rays[:, 10])
    return rays, samp

  @staticmethod
  def backward(ctx, g_rays, g_samp):
    rd, rs, zz = ctx.saved_tensors
    R, S = zz.shape
    g_dy, g_st = torch.empty_like(rd), torch.empty_like(rs)
    gr = f32c(g_rays) if g_rays is not None else torch.zeros(R, 11, device=zz.device)
    gs = f32c(g_samp) if g_samp is not None else None
    with torch.cuda.device(zz.device):
      check(lib.dyn_composite_backward(ptr(rd), ptr(rs), ptr(zz), ptr(gr), ptr(gs) if gs is not None else None,
                                       R, S, ptr(g_dy), ptr(g_st), stream()))
    return g_dy, g_st, None, None, None, None, None


def composite(raw_dy, raw_st, z_vals, mask_dy, mask_st, min_views_dy=1, min_views_st=1):
  """Differentiable raw2outputs: same 11-key dict as render_ray._composite."""
  rays, samp = _Composite.apply(raw_dy, raw_st, z_vals, mask_dy, mask_st, min_views_dy, min_views_st)
  return OrderedDict([
      ("rgb", rays[:, 0:3]), ("rgb_static", rays[:, 3:6]), ("rgb_dy", rays[:, 6:9]),
      ("depth", rays[:, 9]), ("alpha_dy", samp[0]), ("weights_dy", samp[1]),
      ("weights_st", samp[2]), ("alpha", samp[3]), ("weights", samp[4]),
      ("mask", rays[:, 10].detach() > 0.5), ("z_vals", z_vals),
  ])


class _ProjectGather(torch.autograd.Function):
  @staticmethod
  def forward(ctx, xyz_st, xyz, featmaps, query_camera, train_imgs, train_cameras):
    from dynibar_b200.projection import project_gather
    rgb_feat, ray_diff, mask = project_gather(xyz_st, xyz, query_camera, train_imgs, train_cameras, featmaps)
    ctx.save_for_backward(f32c(xyz_st), f32c(xyz), f32c(featmaps), f32c(train_imgs))
    ctx.cams = train_cameras.detach().float().cpu().contiguous()
    ctx.mark_non_differentiable(ray_diff, mask)
    return rgb_feat, ray_diff, mask

  @staticmethod
  def backward(ctx, g_feat, _g_rd, _g_mask):
    xyz_st, xyz, fm, imgs = ctx.saved_tensors
    V, R, S = xyz.shape[:3]
    _, _, H, W, _ = imgs.shape
    _, Cc, h, w = fm.shape
    need_maps, need_xyz = ctx.needs_input_grad[2], ctx.needs_input_grad[1]
    g_maps = torch.empty_like(fm) if need_maps else None
    g_xyz = torch.empty_like(xyz) if need_xyz else None
    gf = f32c(g_feat)
    with torch.cuda.device(fm.device):
      check(lib.dyn_project_gather_backward(ptr(xyz_st), ptr(xyz), ptr(imgs), ctx.cams.data_ptr(), ptr(fm), ptr(gf),
                                            V, R, S, H, W, Cc, h, w,
                                            ptr(g_maps) if need_maps else None,
                                            ptr(g_xyz) if need_xyz else None, stream()))
    return None, g_xyz, g_maps, None, None, None


def project_gather(xyz_st, xyz, query_camera, train_imgs, train_cameras, featmaps):
  """Differentiable Projector.compute_with_motions -> (rgb_feat, ray_diff, mask)."""
  return _ProjectGather.apply(xyz_st, xyz, featmaps, query_camera, train_imgs, train_cameras)


class _MotionMLP(torch.autograd.Function):
  @staticmethod
  def forward(ctx, xyzt, module, *params):
    from dynibar_b200 import weights
    dev = dev_of(xyzt)
    net = weights.packed_of(module, dev)
    x = f32c(xyzt.detach()).reshape(-1, 4)
    N = x.shape[0]
    out = torch.empty(N, 3 * net.num_basis, device=dev)
    nbytes = int(lib.dyn_motion_train_workspace_bytes(N))
    saved = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
      check(lib.dyn_motion_mlp_train_forward(net.handle, ptr(x), N, ptr(out), saved.data_ptr(), nbytes, stream()))
    ctx.net, ctx.ws, ctx.nbytes, ctx.x = net, saved, nbytes, x
    ctx.in_shape = xyzt.shape
    ctx.shapes = [p.shape for p in params]
    return out.reshape(xyzt.shape[:-1] + (out.shape[-1],))

  @staticmethod
  def backward(ctx, g):
    net, x = ctx.net, ctx.x
    N = x.shape[0]
    gc = f32c(g).reshape(N, -1)
    d_params = torch.zeros(net.blob.numel(), device=x.device)
    d_x = torch.empty_like(x) if ctx.needs_input_grad[0] else None
    with torch.cuda.device(x.device):
      check(lib.dyn_motion_mlp_backward(net.handle, ptr(x), ptr(gc), N, ctx.ws.data_ptr(), ctx.nbytes,
                                        ptr(d_params), ptr(d_x) if d_x is not None else None, stream()))
    grads, o = [], 0
    for shp in ctx.shapes:  # state_dict order == parameters() order == the blob's order (weights.py)
      n = 1
      for d in shp:
        n *= d
      grads.append(d_params[o:o + n].reshape(shp))
      o += n
    return (d_x.reshape(ctx.in_shape) if d_x is not None else None, None) + tuple(grads)


def motion_mlp(module, xyzt):
  """Differentiable MotionMLP.forward on [...,4] rows; honours `sf_mag_div` like render_ray.motion_mlp_forward."""
  from dynibar_b200 import weights
  m = weights.de_parallel(module)
  out = _MotionMLP.apply(xyzt, module, *m.parameters())
  div = float(getattr(m, "sf_mag_div", 1.0))
  return out / div if div != 1.0 else out
