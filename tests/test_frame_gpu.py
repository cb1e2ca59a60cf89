"""Frame drivers and cross-time occlusion modes against outputs of the unmodified reference
(fixtures from tests/golden/make_golden_frame.py):

  render_single_image_nvi  ibrnet/render_image.py:9-217    (row a17)
  render_single_image_mono ibrnet/render_image.py:220-439  (row a17, is_train=True)
  occ_weights_mode 1 / 2   ibrnet/render_ray.py:1243-1252  (row a16)

fp32 mode, same tolerance as the per-call golden tests (rtol 5e-4 / atol 5e-5); the frame scenes
use the stress rigs, so a few samples sit on in-bounds discontinuities (max_bad_frac as there).
"""

import os

import pytest
import torch

import scenes
from dynibar_b200 import synthetic
from util import assert_close_frac

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
_dev = lambda x: synthetic.to_device(x, DEV)


@pytest.fixture(autouse=True)
def _fp32_parity_mode():
  from dynibar_b200 import render_ray as rr
  with rr.precision_scope("fp32"):
    yield


def _fixture(name):
  return torch.load(os.path.join(os.path.dirname(__file__), "golden", name + ".pt"), weights_only=False)


def _cmp_block(name, got, fx, stress):
  assert list(got.keys()) == fx[name + "/keys"], name
  for k in fx[name + "/list_keys"]:
    assert isinstance(got[k], list), (name, k)  # per-chunk lists are left unmerged (render_image.py:147-148)
  for k, want in fx[name].items():
    g = got[k]
    assert not g.is_cuda and g.shape == want.shape and g.dtype == want.dtype, (name, k, g.shape, want.shape)
    assert_close_frac("%s/%s" % (name, k), g, want, rtol=5e-4, atol=5e-5, max_bad_frac=0.03 if stress else 1e-3)


def test_render_single_image_nvi_against_reference():
  from dynibar_b200 import sample_ray as sr
  from dynibar_b200.projection import Projector
  from dynibar_b200.render_image import render_single_image_nvi
  fx = _fixture("frame_nvi")
  cfg = fx["cfg"]
  batch, feat_c, feat_f, frame, t, offs, model, args = scenes.build(cfg)
  data = scenes.sampler_data(batch, cfg["H"], cfg["W"], cfg["seed"])
  sampler = sr.RaySamplerSingleImage(data, DEV)
  ret = render_single_image_nvi(frame, t, offs, sampler, sampler.get_all(), synthetic.model_to(model, DEV),
                                Projector(DEV), cfg["chunk"], cfg["N_samples"], args,
                                inv_uniform=cfg["inv_uniform"], N_importance=cfg["N_importance"], det=True,
                                coarse_featmaps=_dev(feat_c), fine_featmaps=_dev(feat_f), is_train=False)
  assert list(ret.keys()) == fx["top_keys"] and ret["outputs_fine"] is None
  assert len(ret["outputs_fine_anchor"]) == 0
  for k in ("outputs_fine_ref", "outputs_coarse_ref"):
    _cmp_block(k, ret[k], fx, cfg["stress"])


def test_render_single_image_mono_against_reference():
  from dynibar_b200 import sample_ray as sr
  from dynibar_b200.projection import Projector
  from dynibar_b200.render_image import render_single_image_mono
  fx = _fixture("frame_mono")
  cfg = fx["cfg"]
  batch, feat_c, feat_f, frame, t, offs, model, args = scenes.build(cfg)
  data = scenes.sampler_data(batch, cfg["H"], cfg["W"], cfg["seed"])
  sampler = sr.RaySamplerSingleImage(data, DEV)
  ret = render_single_image_mono(frame, t, offs, sampler, sampler.get_all(), synthetic.model_to(model, DEV),
                                 Projector(DEV), cfg["chunk"], cfg["N_samples"], args,
                                 inv_uniform=cfg["inv_uniform"], det=True, featmaps=_dev(feat_c), is_train=True,
                                 num_vv=cfg["num_vv"])
  assert list(ret.keys()) == fx["top_keys"] and ret["outputs_fine"] is None
  for k in ("outputs_coarse_ref", "outputs_coarse_st", "outputs_coarse_anchor"):
    _cmp_block(k, ret[k], fx, cfg["stress"])


@pytest.mark.parametrize("name", list(scenes.OCC_MODE_CONFIGS))
def test_occlusion_weight_modes_against_reference(name):
  from dynibar_b200 import render_ray as rr
  from dynibar_b200.projection import Projector
  fx = _fixture("occ_modes")[name]
  cfg = fx["cfg"]
  batch, feat_c, feat_f, frame, t, offs, model, args = scenes.build(cfg)
  assert args.occ_weights_mode == cfg["occ_weights_mode"]
  got = rr.render_rays_mono(frame, t, offs, _dev(batch), synthetic.model_to(model, DEV), _dev(feat_c),
                            Projector(DEV), cfg["N_samples"], args, inv_uniform=cfg["inv_uniform"], det=True,
                            is_train=True, num_vv=cfg["num_vv"])
  for k in ("outputs_coarse_anchor", "outputs_coarse_anchor_dy"):
    for kk, want in fx[k].items():
      assert_close_frac("%s/%s" % (k, kk), got[k][kk], want, rtol=5e-4, atol=5e-5, max_bad_frac=1e-3)
