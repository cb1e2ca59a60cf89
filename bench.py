#!/usr/bin/env python
"""Benchmark of the DynIBaR per-ray IBR hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W              # this repo (CUDA)
  python bench.py --impl reference --gpus N --steps K ...    # CPU reference arm

One "step" = every rank renders `--rays` rays (default 147456 = one full
512x288 frame, in the reference's 8192-ray chunks) of the synthetic scene with
64 coarse + 64 fine samples and 8 dynamic + 8 static source views through
`render_rays_mv`, then the rendered pixels (rgb, depth, mask) are gathered on
rank 0 with one NCCL gather.  Per-GPU work is fixed as N grows ("weak" scaling:
N GPUs render N frames' worth of rays -- e.g. N of the 11 held-out target views
of an eval time step).  Prints ONE JSON line on rank 0.
"""

import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

KERNEL_CLASSES = ["view_static", "view_dynamic", "motion", "point1", "point2", "rgbhead", "attention",
                  "gather"]
# algorithmic MACs per row of each fused kernel (dynibar_b200/flops.py, reference layer widths)
#   view_*: per (point, view); motion/point*: per point; rgbhead: per (point, view)
KERNEL_MAC = {"view_static": 35328 + 2310 + 86528 + 32896 + 16512, "view_dynamic": 123392,
              "motion": 530944, "point1": 98560 + 3 * 16384, "rgbhead": 41664}
# DRAM bytes per (point, view) row of the static per-view kernel from the committed ncu capture
# profiles/r01_view_static_ncu.md (dram__bytes_read.sum + dram__bytes_write.sum per launch / rows)
NCU_DRAM_BYTES_PER_ROW = {"view_static": 291.5}  # profiles/r01_view_twin_ncu.md (152.8 MB / 524 288 rows)

WORKLOAD = dict(H=288, W=512, V_dy=8, V_st=8, N_samples=64, N_importance=64, chunk=8192, seed=0)
METRIC = "rays/sec (64+64 samples x 8 src views)"


def parse():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=3)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
  ap.add_argument("--rays", type=int, default=147456, help="rays per GPU per step (147456 = one 512x288 frame)")
  ap.add_argument("--precision", default="bf16", choices=["fp32", "bf16"])
  ap.add_argument("--ref-rays", type=int, default=64, help="rays per step of the CPU reference arm")
  ap.add_argument("--cpu-rays", type=int, default=64, help="rays of the cpu_baseline sample")
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--view-kernel", default="quad", choices=["quad", "twin"],
                  help="per-view stage kernel: quad schedule (default) or the round-1 twin-warp kernel")
  return ap.parse_args()


def peaks():
  p = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(p):
    d = json.load(open(p))
    return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"],
                source="measured (MEASURED_PEAKS.json)")
  return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
  """Samples SM clock + throttle reasons of one GPU every 200 ms via NVML."""

  REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
             0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake", 0x2: "applications_clocks_setting"}

  def __init__(self, index):
    super().__init__(daemon=True)
    self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
    self._halt = threading.Event()
    self.ok = False
    try:
      import pynvml
      pynvml.nvmlInit()
      self.nv = pynvml
      self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
      self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
      self.ok = True
    except Exception as e:  # NVML missing: report that instead of clocks
      self.err = repr(e)

  def run(self):
    if not self.ok:
      return
    while not self._halt.is_set():
      try:
        self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
        r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for bit, name in self.REASONS.items():
          if r & bit:
            self.reasons.add(name)
      except Exception:
        pass
      self._halt.wait(0.2)

  def finish(self):
    self._halt.set()
    if self.is_alive():
      self.join()
    if not self.ok or not self.samples:
      return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
    s = sorted(self.samples)
    return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
            "samples": len(s)}


def build_scene(rays, seed_offset=0):
  from dynibar_b200 import synthetic
  w = WORKLOAD
  batch, feat_c, feat_f, frame, t, offs = synthetic.make_scene(
      H=w["H"], W=w["W"], V_dy=w["V_dy"], V_st=w["V_st"], seed=w["seed"] + seed_offset, rays=rays)
  model, args = synthetic.make_model(w["N_samples"], w["N_importance"], seed=w["seed"])
  return batch, feat_c, feat_f, frame, t, offs, model, args


def cpu_threads():
  """Threads for the CPU arm: torch's intra-op pool stops scaling (and on the
  128-core GPU hosts collapses to ~2 rays/s) well before all cores on these small
  GEMMs, so the arm uses min(cores, 32) and reports that number as `cores`."""
  return min(os.cpu_count() or 1, 32)


def run_oracle(rays, steps, warmup):
  """CPU arm: the oracle port of the reference's render_rays_mv on all host
  threads (the reference is Python and cannot travel to the GPU box;
  oracle/dynibar_oracle.py is pinned to it by tests/golden)."""
  from oracle import dynibar_oracle as orc
  torch.set_num_threads(cpu_threads())
  batch, feat_c, feat_f, frame, t, offs, model, args = build_scene(rays)
  w = WORKLOAD

  def once():
    with torch.no_grad():
      return orc.render_rays_mv(frame, t, offs, batch, model, None, feat_c, feat_f, w["N_samples"],
                                args, inv_uniform=True, N_importance=w["N_importance"], det=True,
                                is_train=False)
  for _ in range(warmup):
    once()
  t0 = time.perf_counter()
  for _ in range(steps):
    once()
  dt = time.perf_counter() - t0
  return rays * steps / dt, dt / steps


def main():
  a = parse()
  rank = int(os.environ.get("RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  w = WORKLOAD
  config = {"workload": "BASELINE configs[1] shape: synthetic 512x288 frame, 64 coarse + 64 fine "
                        "samples (fine pass evaluates 128), 8 dynamic + 8 static source views, "
                        "render_rays_mv, det=True, inv_uniform=True, chunk 8192",
            "rays_per_gpu_per_step": a.rays, "precision": a.precision,
            "l2": "per-step working set (source maps 66 MB + GBs of per-chunk intermediates) exceeds the "
                  "126 MB L2; plus an explicit 256 MB flush between steps"}

  if a.impl == "reference":
    if rank != 0:
      return 0
    val, sec = run_oracle(a.ref_rays, a.steps, a.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "rays/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(config, rays_per_gpu_per_step=a.ref_rays,
                           note="CPU oracle port of the reference path; bounded sample per step"),
            "cpu_baseline": {"value": val, "unit": "rays/s", "cores": cpu_threads(), "kind": "port",
                             "sample": "%d rays/step of the same workload" % a.ref_rays},
            "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0

  import torch.distributed as dist
  from dynibar_b200 import _lib, distributed as dd, flops, render_ray as rr, synthetic
  from dynibar_b200.projection import Projector
  dev = torch.device("cuda", local)
  torch.cuda.set_device(dev)
  if world > 1:
    dist.init_process_group("nccl", device_id=dev)
  rr.set_precision(a.precision)
  _lib.lib.dyn_debug_set_view_kernel(1 if a.view_kernel == "twin" else 0)

  # every rank renders its own bundle of rays of the same scene (ray shard = rank)
  # every rank renders `rays` rays of its own target view of the same scene
  full = WORKLOAD["H"] * WORKLOAD["W"]
  batch, feat_c, feat_f, frame, t, offs, model, args = build_scene(None if a.rays >= full else a.rays)
  host = dict(batch)
  for k in ("ray_o", "ray_d", "uv_grid"):
    reps = (a.rays + batch[k].shape[0] - 1) // batch[k].shape[0]
    host[k] = batch[k].repeat(reps, 1)[:a.rays].contiguous()
  pin = lambda x: x.pin_memory() if torch.is_tensor(x) else x
  host = {k: pin(v) for k, v in host.items()}
  host_fc = tuple(pin(x) if x is not None else None for x in feat_c)
  host_ff = tuple(pin(x) if x is not None else None for x in feat_f)
  model = synthetic.model_to(model, dev)
  P = Projector(dev)
  b_dev = synthetic.to_device(host, dev)
  fc_dev, ff_dev = synthetic.to_device(host_fc, dev), synthetic.to_device(host_ff, dev)
  flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
  out_host = torch.empty(a.rays, 5, pin_memory=True)

  def render(b, fc, ff):
    outs = []
    for i in range(0, a.rays, w["chunk"]):
      cb = dict(b)
      for k in ("ray_o", "ray_d", "uv_grid"):
        cb[k] = b[k][i:i + w["chunk"]]
      r = rr.render_rays_mv(frame, t, offs, cb, model, P, fc, ff, w["N_samples"], args,
                            inv_uniform=True, N_importance=w["N_importance"], det=True,
                            is_train=False)["outputs_fine_ref"]
      outs.append(torch.cat([r["rgb"], r["depth"][:, None], r["mask"][:, None].float()], 1))
    px = torch.cat(outs, 0)
    if world > 1:  # the path's one exchange step: rendered pixels -> rank 0 over NVLink
      dd.gather_pixels(px, a.rays * world)
    return px

  def step_resident():
    flush.zero_()
    return render(b_dev, fc_dev, ff_dev)

  def step_e2e():
    flush.zero_()
    b = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in host.items()}
    fc = tuple(x.to(dev, non_blocking=True) if x is not None else None for x in host_fc)
    ff = tuple(x.to(dev, non_blocking=True) if x is not None else None for x in host_ff)
    px = render(b, fc, ff)
    out_host.copy_(px, non_blocking=True)
    return px

  def timed(fn, steps):
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
      fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
      dist.barrier()
      dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item()

  for _ in range(max(a.warmup, 3)):
    step_resident()
  torch.cuda.synchronize()
  sampler = ClockSampler(local)
  sampler.start()
  _lib.lib.dyn_launch_count(1)
  _lib.lib.dyn_profile_enable(1)  # CUDA events around the big kernels, on their launching stream
  ms = timed(step_resident, a.steps)
  launches = int(_lib.lib.dyn_launch_count(0))
  clocks = sampler.finish()
  kernel_ms = {}
  for cls, name in enumerate(KERNEL_CLASSES):
    tot, n = ctypes.c_float(), ctypes.c_int()
    _lib.check(_lib.lib.dyn_profile_read(cls, ctypes.byref(tot), ctypes.byref(n)))
    if n.value:
      kernel_ms[name] = (tot.value, n.value)
  _lib.lib.dyn_profile_enable(0)
  step_e2e()
  ms_e2e = timed(step_e2e, a.steps)

  total_rays = a.rays * world * a.steps
  value = total_rays / (ms / 1e3)
  e2e_val = total_rays / (ms_e2e / 1e3)
  h2d = sum(v.numel() * v.element_size() for v in host.values() if torch.is_tensor(v))
  h2d += sum(x.numel() * x.element_size() for x in host_fc + host_ff if x is not None)
  d2h = out_host.numel() * out_host.element_size()

  if rank == 0:
    pk = peaks()
    fpr = flops.flop_per_ray(w["N_samples"], w["N_samples"] + w["N_importance"], w["V_dy"], w["V_st"])
    achieved = (a.rays * a.steps / (ms / 1e3)) * fpr / 1e12  # per GPU
    # dominant kernel (largest share of device time on rank 0), timed live with CUDA events
    S_tot = w["N_samples"] + (w["N_samples"] + w["N_importance"])
    rows_step = {"view_static": a.rays * S_tot * w["V_st"], "view_dynamic": a.rays * S_tot * w["V_dy"],
                 "motion": a.rays * S_tot, "point1": 2 * a.rays * S_tot, "rgbhead": a.rays * S_tot * w["V_st"]}
    kernels = {}
    for name, (tot_ms, n) in kernel_ms.items():
      k = {"ms_per_step": tot_ms / a.steps, "launches_per_step": n / a.steps,
           "share_of_step": tot_ms / ms}
      if name in KERNEL_MAC:
        k["tflops"] = rows_step[name] * a.steps * KERNEL_MAC[name] * 2 / (tot_ms / 1e3) / 1e12
      kernels[name] = k
    dom = max((n for n in kernels if n in KERNEL_MAC), key=lambda n: kernels[n]["ms_per_step"], default=None)
    line = {
        "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": a.steps,
        "warmup": max(a.warmup, 3), "ms_per_step": ms / a.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32" if a.precision == "fp32" else "bf16",
        "data": "synthetic", "config": config, "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": e2e_val, "unit": "rays/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / a.steps},
        "roofline_step": {"bound": "tensor", "achieved": achieved, "peak": pk["tf_sustained"],
                          "unit": "TFLOP/s", "frac": achieved / pk["tf_sustained"],
                          "flop_per_ray": fpr, "peak_source": pk["source"] + ", sustained bf16",
                          "scope": "whole step (all kernels of render_rays_mv), per GPU"},
        "kernels": kernels,
    }
    if dom is not None:
      kd = kernels[dom]
      rows_launch = rows_step[dom] / kd["launches_per_step"]
      line["roofline"] = {
          "kernel": dom, "bound": "tensor", "achieved": kd["tflops"], "peak": pk["tf_sustained"],
          "unit": "TFLOP/s", "frac": kd["tflops"] / pk["tf_sustained"],
          "traffic": (NCU_DRAM_BYTES_PER_ROW[dom] * rows_launch if dom in NCU_DRAM_BYTES_PER_ROW else None),
          "algorithmic_flop_per_launch": rows_launch * KERNEL_MAC[dom] * 2,
          "avg_launch_ms": kd["ms_per_step"] / kd["launches_per_step"],
          "peak_source": pk["source"] + ", sustained bf16 (kernel timed inside a long step)",
          "how": "CUDA events recorded by the library on the launching stream around every launch of "
                 "this kernel inside the timed region (dyn_profile_*)"}
    if world == 1 and not a.no_cpu_baseline:
      cv, csec = run_oracle(a.cpu_rays, 1, 1)
      line["cpu_baseline"] = {"value": cv, "unit": "rays/s", "cores": cpu_threads(), "kind": "port",
                              "sample": "%d rays of the same workload, 1 warm-up + 1 timed call (%.1f s)"
                                        % (a.cpu_rays, csec)}
    print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()
  return 0


if __name__ == "__main__":
  sys.exit(main())
