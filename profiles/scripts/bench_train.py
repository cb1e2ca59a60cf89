#!/usr/bin/env python
"""Times the training step of BASELINE config 3 (N_rand = 1024 rays, 64 coarse samples, render_rays_mono with the
cross-time branch, forward + backward + Adam step) through the differentiable fp32 path.  Secondary measurement,
not bench.py's metric.  Usage: python profiles/scripts/bench_train.py [steps] [rays] [bf16|fp32]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from dynibar_b200 import render_ray as rr, synthetic  # noqa: E402
from dynibar_b200.projection import Projector  # noqa: E402
from dynibar_b200._lib import lib  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
rays = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
prec = sys.argv[3] if len(sys.argv) > 3 else "bf16"
with_encoder = len(sys.argv) > 4 and sys.argv[4] == "encoder"  # feature maps = ResNet(source images), trained too
# multi-GPU (torchrun): the N_rand rays are block-partitioned over the ranks (weights replicated), every rank runs
# forward + backward on its share and ONE flat all-reduce sums the parameter gradients (distributed.allreduce_gradients)
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:
  import torch.distributed as dist
  from dynibar_b200 import distributed as dd
  dist.init_process_group("nccl", device_id=dev)
batch, feat_c, _, frame, t, offs = synthetic.make_scene(H=288, W=512, V_dy=8, V_st=8, num_vv=2, seed=3, rays=rays,
                                                        anchor_offset=2)
if world > 1:
  batch, (lo, hi) = dd.shard_ray_batch(batch, rank, world)
  rays_local = hi - lo
else:
  rays_local = rays
args = synthetic.make_args(1, 1, 0)
model, args = synthetic.make_model(64, 0, args=args, seed=3, mono=True)
model = synthetic.model_to(model, dev)
mods = [model.net_coarse_dy, model.net_coarse_st, model.motion_mlp]
params = []
for m in mods:
  m.requires_grad_(True)
  params += list(m.parameters())
b = synthetic.to_device(batch, dev)
if with_encoder:
  from dynibar_b200 import feature_network
  torch.manual_seed(5)
  enc = feature_network.ResNet().to(dev)
  enc.requires_grad_(True)
  imgs = [b[k][0].permute(0, 3, 1, 2).contiguous() for k in ("src_rgbs", "anchor_src_rgbs", "static_src_rgbs")]
  opt = torch.optim.Adam(params + list(enc.parameters()), lr=1e-4)
else:
  feat = tuple(f.to(dev).requires_grad_(True) for f in feat_c)
  opt = torch.optim.Adam(params + list(feat), lr=1e-4)
torch.manual_seed(11)
target = torch.rand(rays, 3, device=dev)
if world > 1:
  target = target[lo:hi]
proj = Projector(dev)


def step():
  opt.zero_grad(set_to_none=True)
  if with_encoder:  # train.py:264-281: coarse feature maps of the reference / anchor / static source views
    with rr.precision_scope(prec):
      fm = tuple(enc(im)[0] for im in imgs)
  else:
    fm = feat
  ret = rr.render_rays_mono(frame, t, offs, b, model, fm, proj, 64, args, inv_uniform=True, det=False, is_train=True,
                            num_vv=2, precision=prec)
  loss = ((ret["outputs_coarse_ref"]["rgb"] - target) ** 2).mean()
  loss = loss + ((ret["outputs_coarse_anchor"]["rgb"] - target) ** 2).mean()
  loss = loss + 1e-3 * ret["outputs_coarse_ref"]["render_flows"].abs().mean()
  loss = loss + 1e-2 * ret["outputs_coarse_anchor"]["sf_seq"].abs().mean()
  loss.backward()
  if world > 1:
    dd.allreduce_gradients([p for g in opt.param_groups for p in g["params"]])
  opt.step()
  return loss


for _ in range(2):
  l0 = step()
torch.cuda.synchronize()
if world > 1:
  dist.barrier()
n0 = lib.dyn_launch_count(0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
  l1 = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
if world > 1:
  tms = torch.tensor([ms], device=dev)
  dist.all_reduce(tms, op=dist.ReduceOp.MAX)
  ms = float(tms.item())
if rank == 0:
  print(json.dumps({"what": "training step, BASELINE config 3 shape (render_rays_mono is_train=True, fwd + bwd + Adam)",
                  "rays": rays, "n_gpus": world, "samples": 64, "views": "6+2 dynamic, 8 static, 6+2 anchor", "precision": prec, "encoder_in_step": with_encoder,
                  "ms_per_step": ms, "rays_per_s": rays / ms * 1e3, "loss_first": float(l0), "loss_last": float(l1),
                  "kernel_launches_per_step": (lib.dyn_launch_count(0) - n0) / steps,
                  "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}))
if world > 1:
  dist.destroy_process_group()
