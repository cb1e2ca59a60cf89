"""Per-parameter relative L2 error of the bf16 (tcgen05) training gradients against torch fp32 autograd through the
oracle: static net unit case + the whole mono_train step."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch, torch.nn.functional as F
import scenes, test_train_gpu as T
from dynibar_b200 import synthetic, autograd as ag, mlp_network as nets, render_ray as rr
from dynibar_b200.projection import Projector
from oracle import dynibar_oracle as orc
DEV = 'cuda:0'
def rel(a, b): return ((a.detach().cpu().double() - b.detach().double()).norm() / (b.detach().double().norm() + 1e-30)).item()
def static_case(prec):
  R, S, V, aa, mrgb = 40, 16, 8, 1, 0
  torch.manual_seed(R * S + V)
  mod = nets.DynibarStatic(synthetic.make_args(aa, mrgb), 32, S)
  g, pts, feat, mask, ray_dir = T._net_inputs(R, S, V, 11 + V)
  ref_rays = torch.randn(R, 6, generator=g); src_rays = torch.randn(R, S, V, 6, generator=g)
  ray_diff = torch.cat([F.normalize(torch.randn(R, S, V, 3, generator=g), dim=-1), torch.rand(R, S, V, 1, generator=g) * 0.3 + 0.7], -1)
  gen = torch.randn(R, S, 4, generator=g)
  w = T._leaves(mod); fo = feat.clone().requires_grad_(True)
  want = orc.net_static(w, pts, ref_rays, src_rays, fo, ray_diff, mask, anti_alias_pooling=True, mask_rgb=False)
  live = (mask.sum(2) >= 1).float(); scale = torch.cat([torch.ones(R, S, 3), live], -1)
  (want * gen * scale).sum().backward()
  mod = mod.to(DEV).requires_grad_(True); fd = feat.to(DEV).requires_grad_(True); d = lambda x: x.to(DEV)
  got = ag.net_static(mod, d(pts), d(ref_rays), d(src_rays), fd, d(ray_diff), d(mask), precision=prec)
  print('static', prec, 'fwd max err', (got.detach().cpu() - want.detach()).abs().max().item())
  (got * d(gen * scale)).sum().backward()
  for k, p in mod.named_parameters(): print('   %-34s %.4f  (|g| %.3g)' % (k, rel(p.grad, w[k].grad), w[k].grad.norm().item()))
  print('   rgb_feat %.4f' % rel(fd.grad, fo.grad))
def mono_case(prec):
  cfg = dict(scenes.GOLDEN_CONFIGS['mono_train']); cfg['rays'] = 96
  batch, feat_c, _, frame, t, offs, model, args = scenes.build(cfg)
  with torch.no_grad(): model.motion_mlp.coeff_linear.weight.normal_(0.0, 0.05)
  g = torch.Generator().manual_seed(99)
  om = type(model)(**vars(model))
  om.net_coarse_dy = T._leaves(model.net_coarse_dy, model.net_coarse_dy.shift); om.net_coarse_st = T._leaves(model.net_coarse_st); om.motion_mlp = T._leaves(model.motion_mlp)
  fo = tuple(f.clone().requires_grad_(True) for f in feat_c)
  want = orc.render_rays_mono(frame, t, offs, batch, om, fo, None, cfg['N_samples'], args, inv_uniform=True, det=True, is_train=True, num_vv=2)
  gens = {(o, k): torch.randn(want[o][k].shape, generator=g) for o, ks in T._TRAIN_KEYS.items() for k in ks}
  sum((want[o][k] * v).sum() for (o, k), v in gens.items()).backward()
  dev = torch.device(DEV); m_dev = synthetic.model_to(model, dev)
  for mod in (m_dev.net_coarse_dy, m_dev.net_coarse_st, m_dev.motion_mlp): mod.requires_grad_(True)
  fd = tuple(f.to(dev).requires_grad_(True) for f in feat_c)
  got = rr.render_rays_mono(frame, t, offs, synthetic.to_device(batch, dev), m_dev, fd, Projector(dev), cfg['N_samples'], args, inv_uniform=True, det=True, is_train=True, num_vv=2, precision=prec)
  for (o, k), v in gens.items(): print('   out %-28s %-16s max abs err %.3g' % (o, k, (got[o][k].detach().cpu() - want[o][k].detach()).abs().max().item()))
  sum((got[o][k] * v.to(dev)).sum() for (o, k), v in gens.items()).backward()
  for mname, w in (('net_coarse_dy', om.net_coarse_dy), ('net_coarse_st', om.net_coarse_st), ('motion_mlp', om.motion_mlp)):
    for k, p in getattr(m_dev, mname).named_parameters(): print('   %-14s %-34s %.4f (|g| %.3g)' % (mname, k, rel(p.grad, w[k].grad), w[k].grad.norm().item()))
  for i in range(3): print('   featmaps[%d] %.4f' % (i, rel(fd[i].grad, fo[i].grad)))
static_case(sys.argv[1] if len(sys.argv) > 1 else 'bf16')
mono_case(sys.argv[1] if len(sys.argv) > 1 else 'bf16')
