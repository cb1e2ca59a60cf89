"""Backward kernels of the first f2 slice against torch autograd through the oracle's restatement of the same
reference functions (raw2outputs render_ray.py:214-330; compute_with_motions projection.py:103-176)."""

import pytest
import torch

import scenes
from dynibar_b200 import synthetic
from oracle import dynibar_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("R,S", [(37, 64), (5, 20), (16, 192)])
def test_composite_backward_matches_oracle_autograd(R, S):
  from dynibar_b200 import autograd as ag
  torch.manual_seed(R + S)
  raw_dy = torch.randn(R, S, 4)
  raw_st = torch.randn(R, S, 4)
  raw_dy[..., :3].sigmoid_(); raw_st[..., :3].sigmoid_()
  raw_dy[..., 3] -= 2.0; raw_st[..., 3] -= 2.5
  z = torch.sort(torch.rand(R, S) * 20 + 1, dim=1).values
  m_dy = (torch.rand(R, S, 8, 1) > 0.2).float()
  m_st = (torch.rand(R, S, 5, 1) > 0.2).float()
  # upstream gradients for every differentiable key
  keys = ("rgb", "rgb_static", "rgb_dy", "depth", "alpha_dy", "weights_dy", "weights_st", "alpha", "weights")
  a, b = raw_dy.clone().requires_grad_(True), raw_st.clone().requires_grad_(True)
  want = orc.composite(a, b, z, m_dy.sum(2)[..., 0] > 1, m_st.sum(2)[..., 0] > 1)  # per-sample validity
  gens = {k: torch.randn_like(want[k]) for k in keys}
  sum((want[k] * gens[k]).sum() for k in keys).backward()
  ad, bd = raw_dy.to(DEV).requires_grad_(True), raw_st.to(DEV).requires_grad_(True)
  got = ag.composite(ad, bd, z.to(DEV), m_dy.to(DEV), m_st.to(DEV))
  for k in keys:
    torch.testing.assert_close(got[k].detach().cpu(), want[k].detach(), rtol=1e-4, atol=1e-5)
  sum((got[k] * gens[k].to(DEV)).sum() for k in keys).backward()
  torch.testing.assert_close(ad.grad.cpu(), a.grad, rtol=2e-4, atol=2e-5)
  torch.testing.assert_close(bd.grad.cpu(), b.grad, rtol=2e-4, atol=2e-5)


@pytest.mark.parametrize("name", ["mv_linear", "mono_train"])
def test_project_gather_backward_matches_oracle_autograd(golden, name):
  from dynibar_b200 import autograd as ag
  fx = golden(name)
  cfg, st = fx["cfg"], fx["stages"]
  batch, feat_c, feat_f, frame, t, offs, model, args = scenes.build(cfg)
  pts, seq = st["pts"], st["seq"]
  fm = feat_c[0]
  # oracle (CPU, torch autograd through grid_sample)
  fo, so = fm.clone().requires_grad_(True), seq.clone().requires_grad_(True)
  want_f, _, want_m = orc.project_gather(pts, so, batch["camera"], batch["src_rgbs"], batch["src_cameras"], fo)
  g = torch.randn_like(want_f)
  (want_f * g).sum().backward()
  fd, sd = fm.to(DEV).requires_grad_(True), seq.to(DEV).requires_grad_(True)
  d = lambda x: synthetic.to_device(x, DEV)
  got_f, got_rd, got_m = ag.project_gather(d(pts), sd, d(batch["camera"]), d(batch["src_rgbs"]),
                                            d(batch["src_cameras"]), fd)
  assert not got_rd.requires_grad and not got_m.requires_grad
  (got_f * g.to(DEV)).sum().backward()
  torch.testing.assert_close(fd.grad.cpu(), fo.grad, rtol=2e-4, atol=2e-4)
  # d/d xyz: compare where the sample is inside the image (the in-bounds edge is a kink of the zero-padded bilinear)
  inb = (want_m[..., 0] > 0).permute(2, 0, 1)  # [V,R,S]
  gx, wx = sd.grad.cpu()[inb], so.grad[inb]
  bad = ((gx - wx).abs() > 1e-3 + 1e-3 * wx.abs()).float().mean().item()
  assert bad < 2e-3, bad
