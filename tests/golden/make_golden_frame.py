"""Frame-level golden fixtures from the UNMODIFIED reference (run in the build container):

    python tests/golden/make_golden_frame.py

  sampler.pt     RaySamplerSingleImage.get_all / random_sample (ibrnet/sample_ray.py:165-331)
  frame_nvi.pt   render_single_image_nvi  (ibrnet/render_image.py:9-217)
  frame_mono.pt  render_single_image_mono (ibrnet/render_image.py:220-439), is_train=True
  occ_modes.pt   occlusion weights of the cross-time branch for occ_weights_mode 1 and 2
                 (ibrnet/render_ray.py:1243-1252)

Like make_golden.py the fixtures hold reference OUTPUTS only; inputs are regenerated from seeds
(tests/scenes.py) and guarded by a checksum.
"""

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import make_golden as mg  # noqa: E402


def tensors_only(d):
  return {k: v.detach().clone() for k, v in d.items() if torch.is_tensor(v)}


def main():
  import scenes
  ref = mg.import_reference()
  from ibrnet import render_image as ref_ri
  torch.set_grad_enabled(False)

  # ---- sampler ----
  cfg = dict(scenes.GOLDEN_CONFIGS["mono_train"], H=20, W=28, rays=None)
  batch, feat_c, feat_f, frame, t, offs, model, args = scenes.build(cfg)
  data = scenes.sampler_data(batch, cfg["H"], cfg["W"], cfg["seed"])
  fx = {"cfg": cfg}
  s = ref.sr.RaySamplerSingleImage(data, "cpu")
  fx["get_all"] = tensors_only(s.get_all())
  s2 = ref.sr.RaySamplerSingleImage(data, "cpu", render_stride=2)
  fx["get_all_stride2"] = {k: v for k, v in tensors_only(s2.get_all()).items() if k in ("ray_o", "ray_d", "uv_grid")}
  ref.sr.rng = np.random.RandomState(234)  # the module-level stream the reference draws pixels from
  fx["random_center"] = tensors_only(s.random_sample(40, "center", 0.8))
  fx["random_center_inds"] = torch.from_numpy(np.asarray(s.sample_random_pixel(40, "center", 0.6)))
  r = s.random_sample(33, "uniform")
  fx["random_uniform"] = tensors_only(r)
  fx["random_uniform_inds"] = torch.from_numpy(np.asarray(r["selected_inds"]))
  torch.save(fx, os.path.join(HERE, "sampler.pt"))
  print("sampler.pt", os.path.getsize(os.path.join(HERE, "sampler.pt")) // 1024, "KB")

  # ---- frame drivers ----
  for name, cfg in scenes.FRAME_CONFIGS.items():
    batch, feat_c, feat_f, frame, t, offs, model, args = scenes.build(cfg)
    mref = mg.reference_model(ref, model, args, cfg["mono"])
    P = ref.proj.Projector("cpu")
    data = scenes.sampler_data(batch, cfg["H"], cfg["W"], cfg["seed"])
    sampler = ref.sr.RaySamplerSingleImage(data, "cpu")
    ray_batch = sampler.get_all()
    fx = {"cfg": cfg, "checksum": mg.checksum(batch, [feat_c, feat_f])}
    if cfg["mono"]:
      ret = ref_ri.render_single_image_mono(frame, t, offs, sampler, ray_batch, mref, P, cfg["chunk"],
                                            cfg["N_samples"], args, inv_uniform=cfg["inv_uniform"], det=True,
                                            featmaps=feat_c, is_train=True, num_vv=cfg["num_vv"])
      keys = ("outputs_coarse_ref", "outputs_coarse_st", "outputs_coarse_anchor")
    else:
      ret = ref_ri.render_single_image_nvi(frame, t, offs, sampler, ray_batch, mref, P, cfg["chunk"],
                                           cfg["N_samples"], args, inv_uniform=cfg["inv_uniform"],
                                           N_importance=cfg["N_importance"], det=True, coarse_featmaps=feat_c,
                                           fine_featmaps=feat_f, is_train=False)
      keys = ("outputs_fine_ref", "outputs_coarse_ref")
    fx["top_keys"] = list(ret.keys())
    for k in keys:
      fx[k] = tensors_only(ret[k])
      fx[k + "/keys"] = list(ret[k].keys())
      fx[k + "/list_keys"] = [kk for kk, v in ret[k].items() if isinstance(v, list)]
    torch.save(fx, os.path.join(HERE, name + ".pt"))
    print(name + ".pt", os.path.getsize(os.path.join(HERE, name + ".pt")) // 1024, "KB")

  # ---- occlusion-weight modes ----
  fx = {}
  for name, cfg in scenes.OCC_MODE_CONFIGS.items():
    batch, feat_c, feat_f, frame, t, offs, model, args = scenes.build(cfg)
    mref = mg.reference_model(ref, model, args, True)
    ret = ref.rr.render_rays_mono(frame, t, offs, batch, mref, feat_c, ref.proj.Projector("cpu"),
                                  cfg["N_samples"], args, inv_uniform=cfg["inv_uniform"], det=True,
                                  is_train=True, num_vv=cfg["num_vv"])
    fx[name] = {"cfg": cfg, "checksum": mg.checksum(batch, [feat_c, feat_f])}
    for k in ("outputs_coarse_anchor", "outputs_coarse_anchor_dy"):
      fx[name][k] = {kk: ret[k][kk].detach().clone() for kk in ("occ_weights", "occ_weight_map", "weights", "rgb")}
  torch.save(fx, os.path.join(HERE, "occ_modes.pt"))
  print("occ_modes.pt", os.path.getsize(os.path.join(HERE, "occ_modes.pt")) // 1024, "KB")


if __name__ == "__main__":
  main()
