"""World-size-2 gloo test (CPU) of the multi-GPU reduction contract: every rank
computes [sum nll, sum grad, count] over its round-robin task shard, ONE
all-reduce(sum) combines them, and the replicated Adam update is identical on
all ranks.  The per-shard arithmetic comes from the oracle here;
on GPUs it is hb_nll_grad_batched."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hyperbo_b200.gp_utils import gp
from oracle import hyperbo_oracle as O
from tests import helpers as H


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _shard_sums(model, items, d):
  tot, gsum = 0.0, np.zeros(3 + d)
  for _, x, y in items:
    v, g = O.nll_and_grad_sub_dataset("constant", "squared_exponential", model, x,
                                      y, O.DEFAULT_WARP_FUNC)
    tot += v
    gsum += H.grad_vec(g, d)
  return np.concatenate([[tot], gsum, [float(len(items))]])


def _worker(rank, world, port, out):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  d = 2
  ds = O.make_dataset(5, 12, d)
  items = [(k, v[0], v[1]) for k, v in ds.items()]
  model = O.init_raw_params(d)
  opt = O.Adam(1e-2)
  losses = []
  for _ in range(3):
    mine = gp.shard_tasks(items, rank, world)
    sums = torch.from_numpy(_shard_sums(model, mine, d))
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    s = sums.numpy()
    losses.append(s[0] / s[-1])
    g = s[1:-1] / s[-1]
    grads = {"constant": g[0], "signal_variance": g[1], "noise_variance": g[2],
             "lengthscale": g[3:]}
    model = opt.update(model, grads)
  out[rank] = (losses, H.raw_vec(model, d))
  dist.destroy_process_group()


def test_two_rank_sharded_training_matches_single_process():
  world = 2
  mgr = mp.Manager()
  out = mgr.dict()
  mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
  l0, p0 = out[0]
  l1, p1 = out[1]
  assert l0 == l1 and np.array_equal(p0, p1)  # replicas stay bit-identical
  d = 2
  ds = O.make_dataset(5, 12, d)
  ref_model, ref_losses = O.infer_parameters_adam(
      "constant", "squared_exponential", O.init_raw_params(d), ds,
      O.DEFAULT_WARP_FUNC, 1e-2, 3, 1000)
  assert H.rel(l0, ref_losses) < 1e-12
  assert H.rel(p0, H.raw_vec(ref_model, d)) < 1e-12
