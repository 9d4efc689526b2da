"""Multi-GPU parity of the task-sharded training step, through
the product path: every rank runs gp.AdamTrainer on its round-robin task shard;
the P+2 partial sums are exchanged by hb_allreduce_adam_step (NVLink peer
memory, one kernel per rank, inside the step's CUDA graph) and every rank
applies the same Adam update.  Checked against the same trainer run on ONE GPU
over all tasks: parameters and losses agree to 1e-12 relative, the replicas are
bit-identical, and hb_allreduce sums in rank order.

Needs >= 2 GPUs (skipped otherwise): run with `gpurun --gpus 2 -- python -m
pytest tests/test_gpu_multirank.py -m gpu`."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

D, TASKS, N, STEPS, LR = 4, 12, 200, 6, 1e-2


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _batch():
  rng = np.random.default_rng(5)
  x = rng.random((TASKS, N, D))
  y = 5.0 + np.sin(3.0 * x.sum(-1)) + 0.1 * rng.standard_normal((TASKS, N))
  return x, y


def _train(eng, tasks, allreduce, use_graph):
  from hyperbo_b200.engine import PackedDataset
  from hyperbo_b200.gp_utils.gp import AdamTrainer
  x, y = _batch()
  xs = torch.as_tensor(x[tasks].reshape(-1, D), device=eng.device)
  ys = torch.as_tensor(y[tasks].reshape(-1), device=eng.device)
  ds = PackedDataset(tasks, xs, ys, [N * t for t in range(len(tasks) + 1)])
  raw0 = np.concatenate([[5.1, 0.0, -4.0], np.zeros(D)])
  mask = 0b110 | (((1 << D) - 1) << 3)
  tr = AdamTrainer(eng, 2, 1, raw0, mask, D, LR, allreduce=allreduce)
  losses = []
  for _ in range(STEPS):
    tr.step(ds, use_graph=use_graph)
    losses.append(tr.loss())
  return np.array(losses), tr.raw.cpu().numpy()


def _worker(rank, world, port, out):
  import torch.distributed as dist
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  torch.cuda.set_device(rank)
  dist.init_process_group("nccl", rank=rank, world_size=world,
                          device_id=torch.device("cuda", rank))
  from hyperbo_b200.engine import Engine
  from hyperbo_b200.gp_utils.gp import shard_tasks
  eng = Engine.get(rank)
  peer = eng.comm_init()
  # hb_allreduce: rank-ordered sum, identical on every rank
  buf = torch.arange(1, 6, device=eng.device, dtype=torch.float64) * (rank + 1) * 0.1
  if peer:
    eng.allreduce(buf)
  mine = shard_tasks(list(range(TASKS)), rank, world)
  res = {}
  for graph in (False, True):
    res[graph] = _train(eng, mine, True, graph)
  torch.cuda.synchronize()
  out[rank] = (peer, buf.cpu().numpy(), res[False], res[True])
  dist.barrier()
  dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_training_matches_single_gpu():
  import torch.multiprocessing as mp
  world = min(torch.cuda.device_count(), 4)
  mgr = mp.Manager()
  out = mgr.dict()
  mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
  from hyperbo_b200.engine import Engine
  ref_losses, ref_raw = _train(Engine.get(0), list(range(TASKS)), False, False)
  peer0 = out[0][0]
  assert peer0, "peer-memory all-reduce not available on this box"
  want = sum((r + 1) * 0.1 for r in range(world)) * np.arange(1, 6)
  for r in range(world):
    peer, buf, eager, graphed = out[r]
    assert peer
    assert np.allclose(buf, want, rtol=1e-15)
    assert np.array_equal(buf, out[0][1])             # bit-identical replicas
    for losses, raw in (eager, graphed):
      assert np.array_equal(raw, out[0][2][1])         # replicas, eager == graph
      assert np.max(np.abs(losses - ref_losses)) <= 1e-12 * np.max(np.abs(ref_losses))
      assert np.max(np.abs(raw - ref_raw)) <= 1e-12 * np.max(np.abs(ref_raw))
