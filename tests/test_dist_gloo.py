"""World-size-2 gloo test (CPU) of the PRODUCT's multi-rank training path:
gp.infer_parameters on every rank shards the tasks round-robin (gp.shard_tasks),
each step combines the ranks' [sum nll, sum grad, count] with ONE all-reduce
(AdamTrainer, the torch.distributed fallback of the engine's peer-memory
all-reduce) and applies the replicated Adam update.  The engine's
arithmetic comes from tests/fake_engine.py (the oracle on CPU tensors); the
result must equal the oracle's single-process Adam loop, and the replicas must
stay bit-identical.  On GPUs the same host code runs over hb_nll_grad_batched
and hb_allreduce_adam_step (tests/test_gpu_multirank.py)."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from hyperbo_b200.basics import definitions as defs
from hyperbo_b200.gp_utils import gp, kernel, mean, objectives, utils
from oracle import hyperbo_oracle as O
from tests import fake_engine
from tests import helpers as H

D, TASKS, N, STEPS, LR = 2, 5, 12, 3, 1e-2


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _worker(rank, world, port, out):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  from hyperbo_b200 import engine as _engine
  eng = fake_engine.FakeEngine()
  _engine.Engine.get = staticmethod(lambda *a, **k: eng)
  ds = O.make_dataset(TASKS, N, D)
  dataset = {k: defs.SubDataset(*v) for k, v in ds.items()}
  params = defs.GPParams(
      model=dict(O.init_raw_params(D)),
      config={"method": "adam", "learning_rate": LR, "max_training_step": STEPS,
              "batch_size": 1000, "objective": objectives.nll})
  losses = []
  res = gp.infer_parameters(
      mean.constant, kernel.squared_exponential, params, dataset,
      warp_func=utils.DEFAULT_WARP_FUNC, objective=objectives.nll, key=0,
      callback=lambda i, model, loss: losses.append(float(loss)))
  out[rank] = (losses, H.raw_vec({k: np.asarray(v, dtype=np.float64)
                                  for k, v in res.model.items()}, D))
  dist.destroy_process_group()


def test_two_rank_sharded_training_matches_single_process():
  world = 2
  mgr = mp.Manager()
  out = mgr.dict()
  mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
  l0, p0 = out[0]
  l1, p1 = out[1]
  assert l0 == l1 and np.array_equal(p0, p1)  # replicas stay bit-identical
  ds = O.make_dataset(TASKS, N, D)
  ref_model, ref_losses = O.infer_parameters_adam(
      "constant", "squared_exponential", O.init_raw_params(D), ds,
      O.DEFAULT_WARP_FUNC, LR, STEPS, 1000)
  assert len(l0) == STEPS
  assert H.rel(l0, ref_losses[:STEPS]) < 1e-12
  assert H.rel(p0, H.raw_vec(ref_model, D)) < 1e-12


def test_shard_tasks_is_a_balanced_partition():
  items = list(range(11))
  shards = [gp.shard_tasks(items, r, 4) for r in range(4)]
  assert sorted(sum(shards, [])) == items
  assert max(map(len, shards)) - min(map(len, shards)) <= 1
