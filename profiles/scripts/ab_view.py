"""A/B of per-view kernel variants in one process: times net_static_fused / net_dynamic_fused (8192 rays x 128
samples x 8 views, bf16 fused path) for each variant id of dyn_debug_set_view_kernel and checks the outputs agree.
Usage: python profiles/scripts/ab_view.py 2 3"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from dynibar_b200 import synthetic, render_ray as rr
from dynibar_b200._lib import lib
DEV = 'cuda:0'
variants = [int(v) for v in sys.argv[1:]] or [2, 3]
R, S = 8192, 128
batch, feat_c, feat_f, frame, t, offs = synthetic.make_scene(rays=R, seed=0)
model, args = synthetic.make_model(64, 64)
d = lambda x: synthetic.to_device(x, DEV)
b, ff = d(batch), d(feat_f)
m = synthetic.model_to(model, DEV)
pts, z, s = rr.sample_along_camera_ray(b['ray_o'], b['ray_d'], b['depth_range'], S, True, True)
fst, fdy = rr.featmaps_channels_last(ff[2]), rr.featmaps_channels_last(ff[0])
coeff = rr.motion_coefficients(m.motion_mlp_fine, pts, 10 / 24)
seq = rr.displaced_points(pts, coeff, m.trajectory_basis_fine.cpu(), 10, offs[0], 0)
ray_dir = torch.nn.functional.normalize(b['ray_d'], dim=-1)
def run_st(): return rr.net_static_fused(m.net_fine_st, pts, b['ray_o'], b['ray_d'], b['camera'], b['static_src_rgbs'], b['static_src_cameras'], fst)
def run_dy(): return rr.net_dynamic_fused(m.net_fine_dy, pts, seq, ray_dir, b['camera'], b['src_rgbs'], b['src_cameras'], fdy, 10 / 24)
def timeit(fn, n=8):
  for _ in range(2): fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(n): out = fn()
  e1.record(); torch.cuda.synchronize()
  return e0.elapsed_time(e1) / n, out
ref = {}
for rep in range(2):
  for v in variants:
    lib.dyn_debug_set_view_kernel(v)
    ms_s, (raw_s, _) = timeit(run_st)
    ms_d, (raw_d, _) = timeit(run_dy)
    if v == variants[0]: ref = dict(s=raw_s.clone(), d=raw_d.clone())
    es = (raw_s - ref['s']).abs().max().item(); ed = (raw_d[..., :3] - ref['d'][..., :3]).abs().max().item()
    print('variant %d rep %d: static net %.3f ms  dynamic net %.3f ms   max|raw - variant %d| %.2e / %.2e' % (v, rep, ms_s, ms_d, variants[0], es, ed))
