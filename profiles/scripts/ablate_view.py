"""Ablation timing of the fused per-view kernels (results are wrong when DYN_ABLATE != 0;
only the kernel time matters).  python profiles/scripts/ablate_view.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import ctypes
from dynibar_b200 import _lib, synthetic, render_ray as rr

DEV = "cuda:0"
rr.set_precision("bf16")
R, S = 8192, 128
batch, feat_c, feat_f, frame, t, offs = synthetic.make_scene(rays=R, seed=0)
model, args = synthetic.make_model(64, 64)
d = lambda x: synthetic.to_device(x, DEV)
b, ff = d(batch), d(feat_f)
m = synthetic.model_to(model, DEV)
pts, z, s = rr.sample_along_camera_ray(b["ray_o"], b["ray_d"], b["depth_range"], S, True, True)
fcl = rr.featmaps_channels_last(ff[2])
run = lambda: rr.net_static_fused(m.net_fine_st, pts, b["ray_o"], b["ray_d"], b["camera"],
                                  b["static_src_rgbs"], b["static_src_cameras"], fcl)


def timed(n=5):
  run(); run()
  torch.cuda.synchronize()
  _lib.lib.dyn_profile_enable(1)
  for _ in range(n):
    run()
  torch.cuda.synchronize()
  tot, cnt = ctypes.c_float(), ctypes.c_int()
  _lib.check(_lib.lib.dyn_profile_read(0, ctypes.byref(tot), ctypes.byref(cnt)))  # class 0 = view_static
  _lib.lib.dyn_profile_enable(0)
  return tot.value / max(cnt.value, 1)


for tiles in ("1", "2"):
  os.environ["DYN_VIEW_TILES"] = tiles
  for ab, name in [(0, "full"), (1, "no gather loads"), (2, "no X/vis/mask stores"), (4, "no pooled stores"),
                   (8, "no second pooling"), (15, "none of them")]:
    os.environ["DYN_ABLATE"] = str(ab)
    print("tiles/CTA %s  %-22s %.3f ms per view_static launch" % (tiles, name, timed()))
os.environ["DYN_ABLATE"] = "0"
