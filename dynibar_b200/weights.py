"""Weight ingest: nn.Module parameters -> flat fp32 blob + library handle.

Canonical order = `state_dict()` order of the containers in
`dynibar_b200.mlp_network`, which is the reference's own order
(mlp_network.py:159-214 dynamic, :331-403 static incl. the scalar `s` first,
:591-603 motion).  The C side (`csrc/nets_f32.cu: *_layout`) indexes the blob
with the same order; `dyn_net_create` checks the element count.
"""

import ctypes as C

import torch

from dynibar_b200 import _lib

# keyed by class name so the reference's own nn.Modules (built by
# ibrnet/model.py DynibarFF / DynibarMono) are accepted as well as our containers
_KIND = {"DynibarDynamic": _lib.NET_DYNAMIC, "DynibarStatic": _lib.NET_STATIC,
         "MotionMLP": _lib.NET_MOTION}


def de_parallel(m):
  """Unwrap nn.DataParallel like ibrnet/model.py:13-15."""
  return m.module if hasattr(m, "module") else m


def flatten(module):
  sd = de_parallel(module).state_dict()
  return torch.cat([v.detach().reshape(-1).to(torch.float32) for v in sd.values()])


class PackedNet(object):
  """Owns the device blob(s) and the dyn_net_t handle."""

  def __init__(self, module, device, level=2):
    """`level`: 2 = every tensor-core operand image (inference: fused kernels), 1 = per-layer images only (bf16
    training: rebuilt by device kernels after every optimizer step), 0 = fp32 parameters only (fp32 training)."""
    m = de_parallel(module)
    self.kind = _KIND[type(m).__name__]
    self.level = level
    self.blob = flatten(m).to(device).contiguous()
    nbytes = (_lib.lib.dyn_net_packed_bytes(self.kind) if level == 2
              else _lib.lib.dyn_net_layer_images_bytes(self.kind))
    self.packed = None if level == 0 else torch.empty(max(nbytes, 16), dtype=torch.uint8, device=device)
    h = C.c_void_p()
    with torch.cuda.device(device):
      _lib.check(_lib.lib.dyn_net_create_ex(
          self.kind, self.blob.data_ptr(), self.blob.numel(),
          None if level == 0 else self.packed.data_ptr(), level,
          int(getattr(m, "n_samples", 0)), float(getattr(m, "shift", 0.0)),
          int(bool(getattr(m, "anti_alias_pooling", 0))), int(bool(getattr(m, "mask_rgb", 0))),
          _lib.stream(), C.byref(h)))
    self.handle = h
    self.num_basis = (m.coeff_linear.out_features // 3
                      if self.kind == _lib.NET_MOTION else None)

  def __del__(self):
    h = getattr(self, "handle", None)
    if h is not None and h.value and _lib is not None and getattr(_lib, "lib", None) is not None:
      _lib.lib.dyn_net_destroy(h)  # (module globals may already be gone at interpreter exit)
      self.handle = None


def pack(module, device):
  return PackedNet(module, device)


def packed_of(module, device, level=2):
  """Cached PackedNet of a network module; re-packed when any parameter's
  storage or `_version` changes (optimizer steps, load_state_dict).  `level`: see PackedNet."""
  m = de_parallel(module)
  key = (tuple((p.data_ptr(), p._version) for p in m.parameters()), str(device))
  slot = "_dyn_pack_cache_%d" % level if level != 2 else "_dyn_pack_cache"
  cache = m.__dict__.get(slot)
  if cache is None or cache[0] != key:
    cache = (key, PackedNet(m, device, level))
    m.__dict__[slot] = cache
  return cache[1]
