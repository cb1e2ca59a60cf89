"""2-D feature encoder: drop-in for `ibrnet.feature_network.ResNet` (row f1).

The container reproduces the reference module tree (same constructor, same `state_dict` keys -- including
the layer2 / layer3 / decoder parameters the reference builds but never runs, feature_network.py:302-311 --
so published checkpoints load strictly).  `forward` runs the executed part on the CUDA library
(csrc/encoder.cu: conv7x7 s2 -> InstanceNorm -> ReLU -> 3 BasicBlocks -> 1x1 conv) and returns
(coarse [N,32,H/4,W/4], fine [N,32,H/4,W/4]) like the reference.  No PyTorch math here.  With gradients enabled and
parameters that require grad the forward keeps its activations and `backward()` runs the library's encoder backward
(`torch.autograd.Function` below): the gradients of the executed parameters, as train.py's optimiser expects.
"""

import torch
import torch.nn as nn

from dynibar_b200 import _lib
from dynibar_b200._lib import lib, ptr, f32c, check, stream, dev_of


def _conv3x3(i, o, stride=1):
  return nn.Conv2d(i, o, 3, stride=stride, padding=1, bias=False, padding_mode="reflect")


def _in(c):
  return nn.InstanceNorm2d(c, track_running_stats=False, affine=True)


class _BasicBlock(nn.Module):
  """feature_network.py:42-84 (parameters only)."""

  def __init__(self, inplanes, planes, stride=1, downsample=None):
    super().__init__()
    self.conv1, self.bn1 = _conv3x3(inplanes, planes, stride), _in(planes)
    self.conv2, self.bn2 = _conv3x3(planes, planes), _in(planes)
    self.downsample = downsample


class _Conv(nn.Module):
  """feature_network.py:141-160 (parameters only)."""

  def __init__(self, i, o, k, stride):
    super().__init__()
    self.conv = nn.Conv2d(i, o, kernel_size=k, stride=stride, padding=(k - 1) // 2, padding_mode="reflect")
    self.bn = _in(o)


class _UpConv(nn.Module):
  def __init__(self, i, o, k, scale):
    super().__init__()
    self.conv = _Conv(i, o, k, 1)


def _layer(inplanes, planes, blocks, stride):
  down = None
  if stride != 1 or inplanes != planes:
    down = nn.Sequential(nn.Conv2d(inplanes, planes, 1, stride=stride, bias=False, padding_mode="reflect"),
                         _in(planes))
  return nn.Sequential(_BasicBlock(inplanes, planes, stride, down),
                       *[_BasicBlock(planes, planes) for _ in range(1, blocks)])


# executed parameters in the order csrc/encoder.cu expects them (enc_layout)
_EXECUTED = (["conv1.weight", "bn1.weight", "bn1.bias"]
             + [k for b in range(3) for k in
                (["layer1.%d.conv1.weight" % b, "layer1.%d.bn1.weight" % b, "layer1.%d.bn1.bias" % b,
                  "layer1.%d.conv2.weight" % b, "layer1.%d.bn2.weight" % b, "layer1.%d.bn2.bias" % b]
                 + (["layer1.0.downsample.0.weight", "layer1.0.downsample.1.weight",
                     "layer1.0.downsample.1.bias"] if b == 0 else []))]
             + ["out_conv.weight", "out_conv.bias"])


class ResNet(nn.Module):
  """Same constructor and state_dict as ibrnet.feature_network.ResNet (:179-300)."""

  def __init__(self, encoder="resnet34", coarse_out_ch=32, fine_out_ch=32, norm_layer=None, coarse_only=False):
    super().__init__()
    assert encoder in ("resnet18", "resnet34"), "only the BasicBlock encoders are mirrored"
    self.coarse_only = coarse_only
    if coarse_only:
      fine_out_ch = 0
    self.coarse_out_ch, self.fine_out_ch = coarse_out_ch, fine_out_ch
    out_ch = coarse_out_ch + fine_out_ch
    self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False, padding_mode="reflect")
    self.bn1 = _in(64)
    self.layer1 = _layer(64, 64, 3, 2)
    self.layer2 = _layer(64, 128, 4, 2)    # built, never run (feature_network.py:302-311)
    self.layer3 = _layer(128, 256, 6, 2)
    self.upconv3 = _UpConv(256, 128, 3, 2)
    self.iconv3 = _Conv(128 + 128, 128, 3, 1)
    self.upconv2 = _UpConv(128, 64, 3, 2)
    self.iconv2 = _Conv(64 + 64, out_ch, 3, 1)
    self.out_conv = nn.Conv2d(out_ch, out_ch, 1, 1)

  def _blob(self, device):
    key = tuple((p.data_ptr(), p._version) for p in self.parameters())
    cache = self.__dict__.get("_dyn_blob")
    if cache is None or cache[0] != key or cache[1].device != device:
      sd = self.state_dict()
      blob = torch.cat([sd[k].detach().reshape(-1).float() for k in _EXECUTED]).to(device).contiguous()
      cache = (key, blob)
      self.__dict__["_dyn_blob"] = cache
    return cache[1]

  def forward(self, x):
    """x [N,3,H,W] in [0,1] -> (coarse [N,32,h,w], fine [N,32,h,w]), h = H/4, w = W/4."""
    if self.coarse_out_ch != 32 or self.fine_out_ch != 32:
      raise NotImplementedError("the CUDA encoder is built for coarse_out_ch = fine_out_ch = 32")
    if torch.is_grad_enabled() and any(self.state_dict(keep_vars=True)[k].requires_grad for k in _EXECUTED):
      if x.requires_grad:
        raise NotImplementedError("the encoder backward does not produce a gradient of the input images")
      from dynibar_b200.autograd import _prec_code
      sd = self.state_dict(keep_vars=True)
      return _EncoderFn.apply(x, _prec_code(None), *[sd[k] for k in _EXECUTED])
    dev = dev_of(x)
    xi = f32c(x)
    N, _, H, W = xi.shape
    H2, W2 = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
    h, w = (H2 + 2 - 3) // 2 + 1, (W2 + 2 - 3) // 2 + 1
    coarse = torch.empty(N, 32, h, w, device=dev)
    fine = torch.empty(N, 32, h, w, device=dev)
    blob = self._blob(dev)
    nbytes = lib.dyn_encoder_workspace_bytes(N, H, W)
    ws = _lib.workspace.get(nbytes, dev, slot=2)
    with torch.cuda.device(dev):
      check(lib.dyn_encoder_forward(ptr(blob), blob.numel(), ptr(xi), N, H, W, ptr(coarse), ptr(fine),
                                    ws.data_ptr(), nbytes, stream()))
    return coarse, fine


class _EncoderFn(torch.autograd.Function):
  """ResNet.forward (executed part) with its backward kernels (csrc/encoder.cu, training section)."""

  @staticmethod
  def forward(ctx, x, prec, *params):
    dev = dev_of(x)
    xi = f32c(x)
    N, _, H, W = xi.shape
    H2, W2 = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
    h, w = (H2 + 2 - 3) // 2 + 1, (W2 + 2 - 3) // 2 + 1
    blob = torch.cat([p.detach().reshape(-1).float() for p in params]).to(dev).contiguous()
    coarse = torch.empty(N, 32, h, w, device=dev)
    fine = torch.empty(N, 32, h, w, device=dev)
    nbytes = int(lib.dyn_encoder_train_workspace_bytes(N, H, W))
    saved = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
      check(lib.dyn_encoder_train_forward(ptr(blob), blob.numel(), ptr(xi), N, H, W, ptr(coarse), ptr(fine),
                                          saved.data_ptr(), nbytes, stream()))
    ctx.blob, ctx.x, ctx.saved, ctx.nbytes, ctx.prec = blob, xi, saved, nbytes, prec
    ctx.shapes = [p.shape for p in params]
    return coarse, fine

  @staticmethod
  def backward(ctx, g_coarse, g_fine):
    xi = ctx.x
    N, _, H, W = xi.shape
    dev = xi.device
    gc = f32c(g_coarse) if g_coarse is not None else None
    gf = f32c(g_fine) if g_fine is not None else None
    d_params = torch.zeros_like(ctx.blob)
    sbytes = int(lib.dyn_encoder_backward_scratch_bytes(N, H, W))
    scratch = _lib.workspace.get(sbytes, dev, slot=3)
    with torch.cuda.device(dev):
      check(lib.dyn_encoder_backward(ptr(ctx.blob), ctx.blob.numel(), ptr(xi), N, H, W,
                                     ptr(gc) if gc is not None else None, ptr(gf) if gf is not None else None,
                                     ctx.saved.data_ptr(), ctx.nbytes, scratch.data_ptr(), sbytes, ptr(d_params),
                                     ctx.prec, stream()))
    ctx.saved = None
    grads, o = [], 0
    for shp in ctx.shapes:
      n = 1
      for d in shp:
        n *= d
      grads.append(d_params[o:o + n].reshape(shp))
      o += n
    return (None, None) + tuple(grads)
