"""Stand-alone projection+gather kernel (Projector.compute_with_motions) at the
BASELINE config-2 chunk shape, for the HBM-roofline figure."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from dynibar_b200 import _lib, synthetic, render_ray as rr
from dynibar_b200.projection import Projector
import ctypes
DEV = "cuda:0"
R, S = 8192, 128
batch, feat_c, feat_f, frame, t, offs = synthetic.make_scene(rays=R, seed=0)
d = lambda x: synthetic.to_device(x, DEV)
b, ff = d(batch), d(feat_f)
rr.set_precision("fp32")
pts, z, s = rr.sample_along_camera_ray(b["ray_o"], b["ray_d"], b["depth_range"], S, True, True)
P = Projector(DEV)
V = b["static_src_cameras"].shape[1]
for i in range(3):
  out = P.compute_with_motions(pts, pts[None].expand(V, -1, -1, -1).contiguous(), b["camera"],
                               b["static_src_rgbs"], b["static_src_cameras"], ff[2])
torch.cuda.synchronize()
_lib.lib.dyn_profile_enable(1)
for i in range(5):
  out = P.compute_with_motions(pts, pts[None].expand(V, -1, -1, -1).contiguous(), b["camera"],
                               b["static_src_rgbs"], b["static_src_cameras"], ff[2])
torch.cuda.synchronize()
ms, n = ctypes.c_float(), ctypes.c_int()
_lib.check(_lib.lib.dyn_profile_read(7, ctypes.byref(ms), ctypes.byref(n)))
pairs = R * S * V
alg = pairs * 172.0
print("gather: %d launches, %.3f ms avg, algorithmic %.3f GB/launch -> %.1f GB/s" %
      (n.value, ms.value / n.value, alg / 1e9, alg / (ms.value / n.value * 1e-3) / 1e9))
