"""clock64() phase timeline of block 0 of the quad-schedule static per-view kernel
(csrc/view_quad.cu), rows 0 of the four quads + the MMA issuer's per-chunk wait / issue times.

    python profiles/scripts/prof_phases_quad.py [static|dynamic] > gpurun_out/phases_quad.txt
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from dynibar_b200 import _lib, synthetic, render_ray as rr

NAMES_ST = ["geom0+F1op0+issue0", "geom1+F1op1", "consume0+issue1", "F1epi0 (wait+epi)", "consume1+F1epi1",
            "F2+pool1 0", "F2+pool1 1", "F3epi 0", "F3epi 1", "F4epi 0", "F4epi 1", "F5epi 0", "F5epi 1",
            "F6epi 0", "F6epi 1", "F7+pool2 0", "F7+pool2 1"]
NAMES_DY = ["geom+gather 0/1", "pool1 0", "pool1 1", "F3epi 0", "F3epi 1", "F4epi 0", "F4epi 1", "F5epi 0",
            "F5epi 1", "F6epi 0", "F6epi 1", "F7+pool2 0", "F7+pool2 1"]


def main():
  kind = sys.argv[1] if len(sys.argv) > 1 else "static"
  DEV = "cuda:0"
  rr.set_precision("bf16")
  _lib.lib.dyn_debug_set_view_kernel(1)  # quad kernel
  R, S = 8192, 128
  batch, feat_c, feat_f, frame, t, offs = synthetic.make_scene(rays=R, seed=0)
  model, args = synthetic.make_model(64, 64)
  d = lambda x: synthetic.to_device(x, DEV)
  b, ff = d(batch), d(feat_f)
  m = synthetic.model_to(model, DEV)
  pts, z, s = rr.sample_along_camera_ray(b["ray_o"], b["ray_d"], b["depth_range"], S, True, True)
  if kind == "static":
    fcl = rr.featmaps_channels_last(ff[2])
    run = lambda: rr.net_static_fused(m.net_fine_st, pts, b["ray_o"], b["ray_d"], b["camera"],
                                      b["static_src_rgbs"], b["static_src_cameras"], fcl)
    names = NAMES_ST
  else:
    fcl = rr.featmaps_channels_last(ff[0])
    seq = pts[None].repeat(8, 1, 1, 1).contiguous()
    ray_dir = torch.nn.functional.normalize(b["ray_d"], dim=-1)
    run = lambda: rr.net_dynamic_fused(m.net_fine_dy, pts, seq, ray_dir, b["camera"], b["src_rgbs"],
                                       b["src_cameras"], fcl, 0.4)
    names = NAMES_DY
  run(); run()
  torch.cuda.synchronize()
  buf = torch.zeros(1024, dtype=torch.int64, device=DEV)
  _lib.lib.dyn_debug_set_view_timestamps(buf.data_ptr())
  run()
  torch.cuda.synchronize()
  _lib.lib.dyn_debug_set_view_timestamps(None)
  allv = buf.cpu()
  iss = allv[256:259].tolist()
  print("issuer: lifetime %d cycles, waiting for A operands %.1f%%, waiting for weight chunks %.1f%%, "
        "issuing %.1f%%" % (iss[0], 100.0 * iss[1] / iss[0], 100.0 * iss[2] / iss[0],
                            100.0 * (iss[0] - iss[1] - iss[2]) / iss[0]))
  per = len(names) + 1
  base = int(allv[0])
  for q in range(4):
    ts = allv[64 * q:64 * q + 64]
    n = int((ts != 0).sum())
    print("quad", q, "timestamps", n)
    for it in range(min(n // per, 3)):
      seg = ts[it * per:(it + 1) * per]
      dd = (seg[1:] - seg[:-1]).tolist()
      tot = int(seg[-1] - seg[0])
      print(" iter", it, "total", tot, "(start %d)" % (int(seg[0]) - base))
      for nm, x in zip(names, dd):
        print("   %-22s %7d  %5.1f%%" % (nm, x, 100.0 * x / tot))
  ch = allv[264:264 + 480].view(120, 4) - base
  print("issuer per chunk event (cycles rel. to quad 0 iter 0 start): start, A ready, W ready, issued | waitA waitW issue")
  for i in range(120):
    a, b2, c, dd = ch[i].tolist()
    print("  ev %3d  %8d %8d %8d %8d | %6d %6d %6d" % (i, a, b2, c, dd, b2 - a, c - b2, dd - c))


if __name__ == "__main__":
  main()
