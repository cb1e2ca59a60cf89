"""Comparison helpers for parity tests."""

import torch


def assert_close_frac(name, got, want, rtol=2e-4, atol=2e-5, max_bad_frac=0.0):
  """|got-want| <= atol + rtol*|want| for all but `max_bad_frac` of elements.

  A non-zero `max_bad_frac` is only used where the path is discontinuous in
  its inputs (a sample whose projection lands within fp32 rounding of an image
  border flips its in-bounds mask), never to hide a systematic error: the
  remaining elements must meet the tolerance exactly.
  """
  got = got.detach().cpu()
  want = want.detach().cpu()
  assert got.shape == want.shape, "%s: shape %s vs %s" % (name, tuple(got.shape), tuple(want.shape))
  if want.dtype == torch.bool or not want.dtype.is_floating_point:
    bad = (got != want)
  else:
    assert torch.isfinite(got).all(), name + ": non-finite values"
    bad = (got - want).abs() > atol + rtol * want.abs()
  frac = bad.float().mean().item() if bad.numel() else 0.0
  if frac > max_bad_frac:
    if want.dtype.is_floating_point:
      err = (got - want).abs().max().item()
    else:
      err = float("nan")
    raise AssertionError("%s: %.4f%% elements out of tolerance (allowed %.4f%%), max abs err %.3e"
                         % (name, 100 * frac, 100 * max_bad_frac, err))


def psnr(a, b):
  mse = torch.mean((a.double() - b.double()) ** 2).item()
  return float("inf") if mse == 0 else -10.0 * __import__("math").log10(mse)
