"""Checkpoint ingest (row f4): the reference's save format -> the duck-typed model namespace the
renderers read (`net_coarse_st/dy`, `net_fine_st/dy`, `motion_mlp(_fine)`, `trajectory_basis(_fine)`).

Formats (torch.save'd dicts; tensors de-parallelised state_dicts):
  * DynibarFF coarse stage   ibrnet/model.py:190-210 (`load_coarse_model`):
      net_coarse_st, net_coarse_dy, feature_net, motion_mlp, traj_basis, global_step
  * DynibarFF fine stage     ibrnet/model.py:177-190, :211-232 (`save_model` / `load_fine_model`):
      net_fine_st, net_fine_dy, feature_net_fine, motion_mlp_fine, traj_basis_fine, global_step
  * DynibarMono              ibrnet/model.py:424-468:
      net_coarse_st, net_coarse_dy, feature_net, feature_net_st, motion_mlp, traj_basis[, net_fine]
`optimizer` / `scheduler` entries are ignored (inference path).  State dicts are loaded strictly into
the parameter containers of `dynibar_b200.mlp_network`, whose key names are the reference's; the CUDA
library packs them on first use (`weights.packed_of`).  The 2-D encoder state dicts (`feature_net*`)
are kept verbatim on the returned namespace for the encoder stage (row f1).
"""

from types import SimpleNamespace

import numpy as np
import torch

from dynibar_b200 import mlp_network as nets


def init_dct_basis(num_basis, num_frames):
  """DCT-II trajectory basis [T,K] (ibrnet/model.py:18-30)."""
  t = torch.arange(num_frames, dtype=torch.float64)[:, None]
  k = torch.arange(1, num_basis + 1, dtype=torch.float64)[None, :]
  return (np.sqrt(2.0 / num_frames) * torch.cos(np.pi / (2.0 * num_frames) * (2 * t + 1) * k)).float()


def _read(ckpt):
  if ckpt is None or isinstance(ckpt, dict):
    return ckpt
  return torch.load(ckpt, map_location="cpu", weights_only=False)


def _strip(sd):
  """state_dict saved from an nn.DataParallel wrapper without de_parallel carries a 'module.' prefix"""
  return {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}


def _net(cls, sd, *ctor, **kw):
  m = cls(*ctor, **kw)
  m.load_state_dict(_strip(sd), strict=True)
  return m.requires_grad_(False).eval()


def _basis(t):
  return t.detach().float().clone()


def model_from_checkpoints(args, coarse=None, fine=None, mono=False, device=None):
  """Build the model namespace from reference-format checkpoints (paths or already-loaded dicts).

  args needs: N_samples, N_importance, coarse_feat_dim, fine_feat_dim, anti_alias_pooling, mask_rgb,
  input_dir (the reference's config.py names).  `mono=True` reads a DynibarMono checkpoint from
  `coarse` (shift=5 on the dynamic net, ibrnet/model.py:307).  Returns (model, info) where info holds
  the global steps and the untouched encoder state dicts."""
  coarse, fine = _read(coarse), _read(fine)
  m, info = SimpleNamespace(), {"encoders": {}}
  if coarse is not None:
    shift = 5.0 if mono else 0.0
    m.net_coarse_st = _net(nets.DynibarStatic, coarse["net_coarse_st"], args, args.coarse_feat_dim, args.N_samples)
    m.net_coarse_dy = _net(nets.DynibarDynamic, coarse["net_coarse_dy"], args, args.coarse_feat_dim,
                           args.N_samples, shift=shift)
    nb = coarse["motion_mlp"]["coeff_linear.weight"].shape[0] // 3 if "coeff_linear.weight" in coarse["motion_mlp"] \
        else _strip(coarse["motion_mlp"])["coeff_linear.weight"].shape[0] // 3
    m.motion_mlp = _net(nets.MotionMLP, coarse["motion_mlp"], num_basis=nb)
    m.trajectory_basis = _basis(coarse["traj_basis"])
    info["coarse_step"] = int(coarse.get("global_step", 0))
    for k in ("feature_net", "feature_net_st"):
      if k in coarse:
        info["encoders"][k] = coarse[k]
  if fine is not None:
    S = args.N_samples + args.N_importance
    m.net_fine_st = _net(nets.DynibarStatic, fine["net_fine_st"], args, args.fine_feat_dim, S)
    m.net_fine_dy = _net(nets.DynibarDynamic, fine["net_fine_dy"], args, args.fine_feat_dim, S)
    nb = _strip(fine["motion_mlp_fine"])["coeff_linear.weight"].shape[0] // 3
    m.motion_mlp_fine = _net(nets.MotionMLP, fine["motion_mlp_fine"], num_basis=nb)
    m.trajectory_basis_fine = _basis(fine["traj_basis_fine"])
    info["fine_step"] = int(fine.get("global_step", 0))
    if "feature_net_fine" in fine:
      info["encoders"]["feature_net_fine"] = fine["feature_net_fine"]
  if device is not None:
    for k, v in list(vars(m).items()):
      setattr(m, k, v.to(device))
  return m, info


def checkpoint_dicts(model, global_step=0, encoders=None):
  """The inverse, in the reference's key layout (used by tests and for handing weights back to the
  reference): returns (coarse_dict, fine_dict or None)."""
  enc = encoders or {}
  sd = lambda mod: {k: v.detach().cpu().clone() for k, v in mod.state_dict().items()}
  coarse = {"net_coarse_st": sd(model.net_coarse_st), "net_coarse_dy": sd(model.net_coarse_dy),
            "motion_mlp": sd(model.motion_mlp), "traj_basis": model.trajectory_basis.detach().cpu().clone(),
            "global_step": int(global_step)}
  for k in ("feature_net", "feature_net_st"):
    if k in enc:
      coarse[k] = enc[k]
  fine = None
  if hasattr(model, "net_fine_st"):
    fine = {"net_fine_st": sd(model.net_fine_st), "net_fine_dy": sd(model.net_fine_dy),
            "motion_mlp_fine": sd(model.motion_mlp_fine),
            "traj_basis_fine": model.trajectory_basis_fine.detach().cpu().clone(), "global_step": int(global_step)}
    if "feature_net_fine" in enc:
      fine["feature_net_fine"] = enc["feature_net_fine"]
  return coarse, fine
