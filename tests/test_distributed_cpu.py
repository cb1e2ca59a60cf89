"""N>1 host logic on CPU: world_size-2 gloo processes shard a frame's rays,
"render" their block with a deterministic per-ray function (the CUDA kernels
need a GPU; the oracle plays the renderer here) and gather the pixels on rank 0.
The assembled frame must equal the single-process result exactly."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dynibar_b200 import distributed as dd


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _render_stub(batch):
  # any per-ray function: rays are independent
  o, d = batch["ray_o"], batch["ray_d"]
  return torch.cat([torch.sin(o + 2 * d), (o * d).sum(-1, keepdim=True), (d[:, :1] > 0).float()], 1)


def _worker(rank, world, port, n_rays, q):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  g = torch.Generator().manual_seed(3)
  batch = {"ray_o": torch.randn(n_rays, 3, generator=g), "ray_d": torch.randn(n_rays, 3, generator=g),
           "uv_grid": torch.randn(n_rays, 2, generator=g), "camera": torch.zeros(1, 34),
           "flows": torch.randn(6, n_rays, 2, generator=g)}
  local, (lo, hi) = dd.shard_ray_batch(batch, rank, world)
  assert local["ray_o"].shape[0] == hi - lo and local["flows"].shape[1] == hi - lo
  assert local["camera"] is batch["camera"]
  frame = dd.render_frame_sharded(_render_stub, batch)
  t = torch.ones(4) * (rank + 1)
  dd.broadcast_frame_inputs([t, None], src=0)
  assert torch.equal(t, torch.ones(4))
  if rank == 0:
    q.put(torch.equal(frame, _render_stub(batch)))
  else:
    assert frame is None
  dist.barrier()
  dist.destroy_process_group()


@pytest.mark.parametrize("n_rays", [1000, 1001, 7])
def test_two_rank_ray_sharding_matches_single_process(n_rays):
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, 2, port, n_rays, q)) for r in range(2)]
  for p in procs:
    p.start()
  ok = q.get(timeout=120)
  for p in procs:
    p.join(timeout=120)
    assert p.exitcode == 0
  assert ok


def test_shard_bounds_cover_all_rays():
  for n in (0, 1, 7, 147456, 518400):
    for w in (1, 2, 3, 4, 8):
      prev = 0
      for r in range(w):
        lo, hi = dd.shard_bounds(n, r, w)
        assert lo == prev and hi >= lo
        prev = hi
      assert prev == n
      sizes = [dd.shard_bounds(n, r, w)[1] - dd.shard_bounds(n, r, w)[0] for r in range(w)]
      assert max(sizes) - min(sizes) <= 1


def _grad_worker(rank, world, port, q):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  torch.manual_seed(0)
  net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ELU(), torch.nn.Linear(7, 3))
  unused = torch.nn.Parameter(torch.ones(4))  # receives no gradient on any rank
  g = torch.Generator().manual_seed(1)
  x, y = torch.randn(10, 5, generator=g), torch.randn(10, 3, generator=g)
  lo, hi = dd.shard_bounds(10, rank, world)
  # sum-of-squares loss over this rank's rows; averaging the ranks' gradients = the full-batch gradient / world
  ((net(x[lo:hi]) - y[lo:hi]) ** 2).sum().backward()
  n = dd.allreduce_gradients(list(net.parameters()) + [unused])
  assert n == sum(p.numel() for p in net.parameters()) + 4
  if rank == 0:
    ref = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ELU(), torch.nn.Linear(7, 3))
    ref.load_state_dict(net.state_dict())
    ((ref(x) - y) ** 2).sum().backward()
    ok = all(torch.allclose(p.grad * world, q_.grad, rtol=1e-5, atol=1e-6)
             for p, q_ in zip(net.parameters(), ref.parameters()))
    q.put(ok and bool((unused.grad == 0).all()))
  dist.barrier()
  dist.destroy_process_group()


def test_two_rank_gradient_allreduce_matches_full_batch():
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  ok = q.get(timeout=120)
  for p in procs:
    p.join(timeout=120)
    assert p.exitcode == 0
  assert ok
