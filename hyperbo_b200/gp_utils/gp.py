"""Inference and prediction for a (multi-task) GP -- mirrors
hyperbo/gp_utils/gp.py (infer_parameters :53-195, predict :242-305, GP :308-620)
with the per-task hot path running in the B200 engine.

Differences forced by the platform (documented in DESIGN.md):
  * arrays are torch CUDA float64 tensors; `key` arguments are int seeds or
    torch.Generators (jax.random does not exist here);
  * the objective must be built from this package's objectives (nll, kl / ekl /
    regkl and their add / mul combinations, or their names as strings): the
    engine differentiates them in closed form and cannot trace arbitrary
    Python callables.
"""
from __future__ import annotations

import logging
import math
import os
from typing import Any, Dict, List, Union

import numpy as np
import torch

from hyperbo_b200 import engine as _engine
from hyperbo_b200.basics import data_utils
from hyperbo_b200.basics import definitions as defs
from hyperbo_b200.basics import linalg
from hyperbo_b200.basics import params_utils
from hyperbo_b200.gp_utils import kernel as _kernel
from hyperbo_b200.gp_utils import mean as _mean
from hyperbo_b200.gp_utils import objectives as obj

retrieve_params = params_utils.retrieve_params

GPCache = defs.GPCache
SubDataset = defs.SubDataset
GPParams = defs.GPParams


def _is_nll(objective) -> bool:
  """True for the plain NLL (the fast path: hb_nll_grad_batched on one packed
  batch); other objectives run as an objectives.ObjectiveProgram."""
  return obj.objective_terms(objective) == [(1.0, "nll", {})]


def _dist_world():
  import torch.distributed as dist
  if dist.is_available() and dist.is_initialized():
    return dist.get_rank(), dist.get_world_size()
  return 0, 1


class AdamTrainer:
  """Device-resident Adam loop state for gp.infer_parameters (gp.py:114-157).

  One `step()` = batched NLL + gradient over this rank's task shard
  (hb_nll_grad_batched), one all-reduce(sum) of the P+2 partial sums when
  several ranks share the dataset, and one optax.adam update (hb_adam_step).
  Parameters and optimiser state never leave the device; `loss()` is the one
  host read the reference's isfinite check needs.
  """

  def __init__(self, eng, kid, mid, raw_init, mask, d, lr, b1=0.9, b2=0.999,
               eps=1e-8, allreduce=False, tie_lengthscale=False):
    self.eng, self.kid, self.mid, self.mask, self.d = eng, kid, mid, mask, d
    self.P = 3 + d
    self.lr, self.b1, self.b2, self.eps = lr, b1, b2, eps
    dev, dt = eng.device, eng.dtype
    self.raw = eng.tensor(raw_init).clone()
    self.m = torch.zeros(self.P, device=dev, dtype=dt)
    self.v = torch.zeros(self.P, device=dev, dtype=dt)
    self.accepted = self.raw.clone()
    self.sums = torch.zeros(self.P + 2, device=dev, dtype=dt)
    self.scal = torch.zeros(4, device=dev, dtype=dt)
    self.allreduce = allreduce
    self.tie_lengthscale = tie_lengthscale
    self._graphs = {}
    self._graph_gen = -1
    self._graph_failed = False
    self._copy_stream = None
    self._host_bufs = None
    self._host_step = 0
    self._loss_pin = None
    self._loss_evt = None
    self._nsteps = 0

  def _peer_comm(self) -> bool:
    """True when the all-reduce runs in the engine (NVLink peer memory, fused
    with the Adam update: hb_allreduce_adam_step) instead of in NCCL."""
    if not self.allreduce or os.environ.get("HB_PEER_ALLREDUCE", "1") == "0":
      return False
    init = getattr(self.eng, "comm_init", None)
    return bool(init and init())

  def _enqueue(self, ds):
    if isinstance(ds, obj.ObjectiveProgram):
      # general objective (nll + c * kl, ...): its launches, scaling and
      # all-reduce are the program's; sums = [value, gradient, 1]
      ds.sums(self.raw, self.mask, out=self.sums)
    else:
      sampler = getattr(ds, "sampler", None)
      if sampler is not None:
        # per-step sub-sampling on the device, keyed by the step counter the
        # Adam kernel maintains in scal[1] (data_utils.py:72-100)
        sampler.sample(scal=self.scal)
      self.eng.nll_grad(self.kid, self.mid, ds, self.raw, self.mask,
                        sums_out=self.sums)
      if self.allreduce:
        if self._peer_comm():
          # ONE kernel: exchange of the P+2 partial sums over peer memory, the
          # rank-ordered sum and the Adam update
          self.eng.allreduce_adam_step(self.P, self.raw, self.m, self.v,
                                       self.accepted, self.sums, self.scal,
                                       self.lr, self.b1, self.b2, self.eps,
                                       self.tie_lengthscale)
          return
        import torch.distributed as dist
        dist.all_reduce(self.sums, op=dist.ReduceOp.SUM)
    self.eng.adam_step(self.P, self.raw, self.m, self.v, self.accepted,
                       self.sums, self.scal, self.lr, self.b1, self.b2,
                       self.eps, self.tie_lengthscale)

  def _graphed(self, key, ds, fn):
    """Capture fn() once into a CUDA graph (after an eager warm-up that sizes
    the engine workspace) and replay it afterwards.  A graph holds raw engine
    workspace pointers and the batch's device buffers: it is dropped when the
    handle's generation changes (a buffer was re-allocated / a plan evicted,
    e.g. by a callback that ran predict on a larger task) and it is only
    replayed for the very same batch object."""
    gen = self.eng.generation() if hasattr(self.eng, "generation") else 0
    if gen != self._graph_gen:
      self._graphs = {}
      self._graph_gen = gen
    ent = self._graphs.get(key)
    if ent is None or ent[1] is not ds:
      tensors = (self.raw, self.m, self.v, self.accepted, self.scal)
      state = [t.clone() for t in tensors]
      s = torch.cuda.Stream(device=self.eng.device)
      s.wait_stream(torch.cuda.current_stream(self.eng.device))
      with torch.cuda.stream(s):
        fn()
      torch.cuda.current_stream(self.eng.device).wait_stream(s)
      torch.cuda.synchronize(self.eng.device)
      for t, c in zip(tensors, state):
        t.copy_(c)
      if hasattr(self.eng, "generation"):
        self._graph_gen = self.eng.generation()  # (the warm-up may have grown it)
      # capture by hand: the torch.cuda.graph() context adds a gc.collect() and an
      # empty_cache() (~7 ms of the fixed cost of a GP.train() call)
      g = torch.cuda.CUDAGraph()
      cur = torch.cuda.current_stream(self.eng.device)
      s.wait_stream(cur)
      with torch.cuda.stream(s):
        g.capture_begin()
        try:
          fn()
        finally:
          g.capture_end()
      cur.wait_stream(s)
      for t, c in zip(tensors, state):
        t.copy_(c)
      if len(self._graphs) >= 4:  # bounded cache (a new batch object per step)
        self._graphs.pop(next(iter(self._graphs)))
      ent = (g, ds)  # (keeps the batch alive as long as its graph)
      self._graphs[key] = ent
    ent[0].replay()

  def _graph_ok(self, use_graph):
    if not use_graph or self._graph_failed or self.eng.device.type != "cuda":
      return False
    # several ranks: the whole step (incl. the peer-memory all-reduce + Adam
    # kernel) is one graph; with the NCCL fallback the collective stays eager
    # (capturing it hung on the round-1 test box; HB_GRAPH_NCCL=1 opts in)
    return (not self.allreduce or self._peer_comm() or
            os.environ.get("HB_GRAPH_NCCL", "0") == "1")

  def _run(self, key, ds, fn, use_graph, force=False):
    if not (force or self._graph_ok(use_graph)):
      fn()
      return
    try:
      self._graphed(key, ds, fn)
    except RuntimeError as e:  # capture not possible here: stay eager
      logging.warning("CUDA-graph capture failed (%s); using eager launches", e)
      self._graph_failed = True
      self._graphs = {}
      torch.cuda.synchronize(self.eng.device)
      fn()

  def step(self, ds, use_graph=False):
    """Enqueue one optimiser step on the current stream."""
    if (use_graph and self.allreduce and not self._graph_failed and
        self.eng.device.type == "cuda" and not self._peer_comm() and
        not isinstance(ds, obj.ObjectiveProgram) and
        os.environ.get("HB_GRAPH_NCCL", "0") != "1"):
      # NCCL fallback: the kernel sequence of hb_nll_grad_batched replays from a
      # CUDA graph; the all-reduce and the Adam kernel stay eager.
      self._run(("nll", id(ds)), ds,
                lambda: self.eng.nll_grad(self.kid, self.mid, ds, self.raw,
                                          self.mask, sums_out=self.sums),
                True, force=True)
      import torch.distributed as dist
      dist.all_reduce(self.sums, op=dist.ReduceOp.SUM)
      self.eng.adam_step(self.P, self.raw, self.m, self.v, self.accepted,
                         self.sums, self.scal, self.lr, self.b1, self.b2,
                         self.eps, self.tie_lengthscale)
      return
    self._run(("dev", id(ds)), ds, lambda: self._enqueue(ds), use_graph)

  def step_from_host(self, ds, x_host: torch.Tensor, y_host: torch.Tensor,
                     use_graph=False):
    """One optimiser step whose batch arrives in (pinned) HOST memory: copies
    it into the packed device batch `ds` on the current stream, then steps.
    This is the shape of the reference's loop, where every step receives a
    fresh (sub-sampled) batch from the host iterator (gp.py:133)."""

    # Double-buffered upload on a copy stream: the H2D copy of THIS step's batch
    # overlaps the previous step's kernels (it lands in the buffer that step
    # k-2 used); the compute stream waits for it, then runs the step.
    dev = self.eng.device
    if self._host_bufs is None or self._host_bufs[0] is not ds:
      other = _engine.PackedDataset(ds.keys, torch.empty_like(ds.x),
                                    torch.empty_like(ds.y), ds.offs)
      self._host_bufs = (ds, other)
      self._copy_stream = torch.cuda.Stream(device=dev)
      self._copied = [torch.cuda.Event() for _ in range(2)]
      self._read_done = [None, None]
      self._host_step = 0
    b = self._host_step & 1
    buf = self._host_bufs[b]
    cur = torch.cuda.current_stream(dev)
    cs = self._copy_stream
    if self._read_done[b] is not None:
      cs.wait_event(self._read_done[b])
    else:
      cs.wait_stream(cur)
    with torch.cuda.stream(cs):
      buf.x.copy_(x_host, non_blocking=True)
      buf.y.copy_(y_host, non_blocking=True)
      self._copied[b].record(cs)
    cur.wait_event(self._copied[b])
    self.step(buf, use_graph=use_graph)
    if self._read_done[b] is None:
      self._read_done[b] = torch.cuda.Event()
    self._read_done[b].record(cur)
    self._host_step += 1

  def loss(self) -> float:
    return float(self.scal[0])  # device -> host sync (gp.py:135-138)

  # ---- pipelined loss read-back -------------------------------------------
  # The reference reads the loss on the host after every step (gp.py:135-142).
  # The Adam kernel carries the same accept / stop logic on the device
  # (scal[2]: once a loss is non-finite nothing is updated any more), so the
  # host may run ONE step ahead: step k+1 is enqueued before the loss of step k
  # is read from pinned memory.  Still one read-back per step, but the GPU never
  # idles while the host marshals the next step.
  def step_pipelined(self, ds, x_host=None, y_host=None, use_graph=False):
    """Enqueue a step (optionally from host buffers) and return the loss of
    the PREVIOUS step (None on the first call)."""
    if self._loss_pin is None:
      self._loss_pin = [torch.zeros(1, dtype=self.eng.dtype).pin_memory()
                        for _ in range(2)]
      self._loss_evt = [torch.cuda.Event() for _ in range(2)]
    if x_host is not None:
      self.step_from_host(ds, x_host, y_host, use_graph=use_graph)
    else:
      self.step(ds, use_graph=use_graph)
    slot = self._nsteps & 1
    self._loss_pin[slot].copy_(self.scal[0:1], non_blocking=True)
    self._loss_evt[slot].record(torch.cuda.current_stream(self.eng.device))
    self._nsteps += 1
    if self._nsteps == 1:
      return None
    prev = (self._nsteps - 2) & 1
    self._loss_evt[prev].synchronize()
    return float(self._loss_pin[prev][0])

  def flush(self):
    """Loss of the last enqueued step (waits for it)."""
    if self._nsteps == 0:
      return None
    last = (self._nsteps - 1) & 1
    self._loss_evt[last].synchronize()
    return float(self._loss_pin[last][0])

  @property
  def stopped(self) -> bool:
    return bool(self.scal[2] != 0)


def shard_tasks(items: List, rank: int, world: int) -> List:
  """Round-robin task shard  t = rank (mod world)  (SURVEY.md 8e)."""
  return [it for t, it in enumerate(items) if t % world == rank]


def _infer_parameters_quasi_newton(eng, kid, mid, params, dataset, warp_func,
                                   method, key, batch_size, pack, callback,
                                   world, get_params_path=lambda x=0: None):
  """L-BFGS / BFGS branches of infer_parameters (gp.py:158-191): ONE
  sub-sampled batch (gp.py:102-107), then a host-side quasi-Newton driver whose
  objective is the engine's batched value-and-gradient."""
  from hyperbo_b200.basics import bfgs as _bfgs
  from hyperbo_b200.basics import lbfgs as _lbfgs
  # only lbfgs sub-samples (gp.py:102-107, "to handle very large sub
  # datasets"); bfgs optimises the full dataset
  batch = (next(data_utils.sub_sample_dataset_iterator(key, dataset, batch_size))
           if method == "lbfgs" else dataset)
  ds = pack(batch)
  any_x = next(iter(dataset.values())).x
  d = int(torch.as_tensor(any_x).shape[1])
  need_mean = mid == 1
  template = dict(params.model)
  raw0, mask, scalar_ls = params_utils.pack_raw(template, d, need_mean,
                                                warp_func)
  # optimisation variables = the model's own entries (a scalar lengthscale is
  # ONE variable: broadcast in, summed gradient out)
  def to_raw(v):
    raw = np.empty(3 + d)
    raw[0] = v[0] if need_mean else 0.0
    raw[1], raw[2] = v[1], v[2]
    raw[3:] = v[3] if scalar_ls else v[3:]
    return raw

  def from_raw_grad(g):
    ls = [g[3:].sum()] if scalar_ls else list(g[3:])
    return np.array([g[0] if need_mean else 0.0, g[1], g[2]] + ls)

  v0 = np.array([raw0[0], raw0[1], raw0[2]] +
                ([raw0[3]] if scalar_ls else list(raw0[3:])))

  def val_and_grad(v):
    if isinstance(ds, obj.ObjectiveProgram):
      sums = ds.sums(to_raw(v), mask)
    else:
      sums = eng.nll_grad(kid, mid, ds, to_raw(v), mask)
      if world > 1:
        import torch.distributed as dist
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    s = sums.cpu().numpy()
    cnt = max(s[-1], 1.0)
    return float(s[0] / cnt), from_raw_grad(s[1:-1] / cnt)

  def cb(step, model_params, loss):
    if callback:
      callback(step, params_utils.unpack_like(template, to_raw(model_params), d,
                                              need_mean), loss)

  def val_and_grad_multi(vs):
    """Several points in ONE engine call (hb_nll_grad_multi, the second batch
    axis): the speculative line search of basics/lbfgs.py."""
    raws = np.stack([to_raw(v) for v in vs])
    s = eng.nll_grad_multi(kid, mid, ds, raws, mask).cpu().numpy()
    out = []
    for row in s:
      cnt = max(row[-1], 1.0)
      out.append((float(row[0] / cnt), from_raw_grad(row[1:-1] / cnt)))
    return out

  # speculation pays while three parameter sets still fit the GPU in one wave of
  # the persistent kernel (small / few tasks: the notebook-scale case, where a
  # call is latency-bound); bigger batches keep the plain sequential search
  multi = None
  if (method == "lbfgs" and world == 1 and
      not isinstance(ds, obj.ObjectiveProgram) and
      getattr(eng, "h", None) is not None and
      os.environ.get("HB_LBFGS_SPECULATE", "1") != "0"):
    tiles = sum(((ds.offs[t + 1] - ds.offs[t] + 63) // 64) *
                ((ds.offs[t + 1] - ds.offs[t] + 63) // 64 + 1) // 2
                for t in range(ds.num_tasks))
    if 0 < 3 * tiles <= 444:
      multi = val_and_grad_multi

  if method == "lbfgs":
    ls_stats = {}
    final_loss, v, _ = _lbfgs.lbfgs(
        val_and_grad, v0, steps=params.config["max_training_step"],
        alpha=params.config.get("alpha", 1.0), callback=cb, multi_fn=multi,
        stats=ls_stats)
    logging.info("lbfgs: %s engine calls for %s points", ls_stats.get("calls"),
                 ls_stats.get("points"))
  else:
    v, _ = _bfgs.bfgs(val_and_grad, v0, tol=params.config["tol"],
                      max_training_step=params.config["max_training_step"])
  params.model = params_utils.unpack_like(template, to_raw(v), d, need_mean)
  if method == "lbfgs":  # gp.py:186-191 (the bfgs branch does not log / save)
    params_utils.log_params_loss(step=params.config["max_training_step"],
                                 params=params, loss=final_loss,
                                 warp_func=warp_func,
                                 params_save_file=get_params_path())
  params.cache = {}
  return params


def infer_parameters(mean_func,
                     cov_func,
                     init_params,
                     dataset,
                     warp_func=None,
                     objective=obj.neg_log_marginal_likelihood,
                     key=None,
                     get_params_path=None,
                     callback=None):
  """Posterior inference for a meta GP (gp.py:53-195, Adam branch :114-157).

  When torch.distributed is initialised with world_size > 1, every rank passes
  the SAME dataset; tasks are sharded round-robin across ranks and the partial
  sums are combined with one all-reduce per step, so all ranks return identical
  parameters.
  """
  if not get_params_path:
    get_params_path = lambda x=0: None  # gp.py:90-91
  if key is None:
    key = 0
    logging.info("Using default random state in infer_parameters.")
  if not dataset:
    logging.info("No dataset present to train GP.")
    return init_params
  params = init_params
  method = params.config["method"]
  batch_size = params.config["batch_size"]
  max_training_step = init_params.config["max_training_step"]
  if max_training_step <= 0 and method != "slice_sample":
    return init_params
  if method not in ("adam", "lbfgs", "bfgs"):
    raise ValueError(f"Optimization method {method} is not supported.")
  plain_nll = _is_nll(objective)  # raises for objectives the engine cannot
  if "priors" in params.config:    # differentiate
    raise NotImplementedError("log-prior terms (objectives.py:197-207)")

  kid = _kernel.kernel_id_of(cov_func)
  mid = _mean.mean_id_of(mean_func)
  eng = _engine.Engine.get()
  rank, world = _dist_world()

  def pack(batch):
    if not plain_nll:
      prog = obj.compile_objective(objective, mean_func, cov_func, batch, rank,
                                   world)
      if not prog.has_exact_grad:
        raise NotImplementedError(
            "kl_multivariate_normal with eps > 0 is value-only on the engine")
      return prog
    items = obj._select(batch, exclude_aligned=True)  # pylint: disable=protected-access
    return eng.pack(shard_tasks(items, rank, world))

  # data_utils.sub_sample_dataset_iterator only changes tasks with
  # n >= batch_size; when none qualifies the batch is the dataset every step.
  dataset = {k: SubDataset(*v) for k, v in dataset.items()}
  if method != "adam":
    return _infer_parameters_quasi_newton(
        eng, kid, mid, params, dataset, warp_func, method, key, batch_size,
        pack, callback, world, get_params_path)
  needs_subsample = any(
      torch.as_tensor(s.x).shape[0] >= batch_size for s in dataset.values())
  dataset_iter = data_utils.sub_sample_dataset_iterator(key, dataset,
                                                        batch_size)
  static_ds = None if needs_subsample else pack(dataset)
  if (needs_subsample and plain_nll and getattr(eng, "h", None) is not None and
      os.environ.get("HB_DEVICE_SUBSAMPLE", "1") != "0"):
    # sub-sample on the device into a fixed-shape packed batch: the step keeps
    # its CUDA graph and no batch is re-packed on the host (the host iterator
    # below remains for objective programs and the CPU test double)
    items = obj._select(dataset, exclude_aligned=True)  # pylint: disable=protected-access
    ids = shard_tasks(list(range(len(items))), rank, world)
    src = eng.pack(shard_tasks(items, rank, world))
    seed = (int(key.initial_seed()) if isinstance(key, torch.Generator) else int(key))
    static_ds = _engine.DeviceSampler(eng, src, batch_size, seed, task_ids=ids).dst
    needs_subsample = False
  any_x = next(iter(dataset.values())).x
  d = int(torch.as_tensor(any_x).shape[1])
  raw0, mask, scalar_ls = params_utils.pack_raw(params.model, d, mid == 1,
                                                warp_func)
  trainer = AdamTrainer(eng, kid, mid, raw0, mask, d,
                        params.config["learning_rate"], allreduce=world > 1,
                        tie_lengthscale=scalar_ls and d > 1)
  model_template = dict(params.model)

  def to_model(vec):
    return params_utils.unpack_like(model_template, vec, d, mid == 1)

  ds = static_ds
  ran = False

  def check(i, current_loss):
    """gp.py:135-142; returns False when the loop must stop."""
    if math.isnan(current_loss) and i == 0:
      raise ValueError("Encountered NaN in loss function. current_loss = "
                       f"{current_loss}.")
    if not math.isfinite(current_loss):
      logging.info(msg=f"{method} stopped at step {i} due to instability.")
      return False
    return True

  if callback is None and not needs_subsample and eng.device.type == "cuda":
    # one step ahead of the loss read-back (see AdamTrainer.step_pipelined)
    for i in range(max_training_step):
      prev = trainer.step_pipelined(ds, use_graph=True)
      ran = True
      if prev is not None and not check(i - 1, prev):
        break
    else:
      check(max_training_step - 1, trainer.flush())
  else:
    for i in range(max_training_step):
      if needs_subsample:
        ds = pack(next(dataset_iter))
      trainer.step(ds, use_graph=not needs_subsample)
      ran = True
      current_loss = trainer.loss()
      if not check(i, current_loss):
        break
      if callback:
        callback(i, to_model(trainer.accepted), current_loss)
  if ran:
    final = trainer.accepted
    if not trainer.stopped:
      # gp.py:147-150: evaluate the last update once more, accept iff finite
      if plain_nll:
        sums = eng.nll_grad(kid, mid, ds, trainer.raw, mask)
        if world > 1:
          import torch.distributed as dist
          dist.all_reduce(sums, op=dist.ReduceOp.SUM)
      else:
        sums = ds.sums(trainer.raw, mask)
      final_loss = float(sums[0] / sums[-1])
      if math.isfinite(final_loss):
        final = trainer.raw
    else:
      final_loss = float("nan")
    params.model = to_model(final)
    # gp.py:151-157: log (and checkpoint, if a path is given) the final state
    params_utils.log_params_loss(step=max_training_step, params=params,
                                 loss=final_loss, warp_func=warp_func,
                                 params_save_file=get_params_path())
  params.cache = {}
  return params


def sample_from_gp(key, mean_func, cov_func, params, x, warp_func=None,
                   num_samples=1, method="cholesky", eps=1e-6):
  """Sample functions from a GP at x (gp.py:198-239): mean + chol(K+noise+eps) z.

  The factor comes from the engine (hb_factorize_batched on one pseudo task at
  the reference's jitter eps = 1e-6, linalg.py:42; another eps falls back to an
  explicit-matrix Cholesky)."""
  del method
  eng = _engine.Engine.get()
  xt = eng.tensor(x)
  n, d = xt.shape
  gen = key if isinstance(key, torch.Generator) else torch.Generator().manual_seed(
      int(key) if key is not None else 0)
  z = torch.randn((n, num_samples), generator=gen, dtype=torch.float64).to(
      device=eng.device, dtype=eng.dtype)
  mean = eng.tensor(mean_func(params, x, warp_func=warp_func))
  if eps == _engine.JITTER and getattr(eng, "h", None) is not None and n > 0:
    kid = _kernel.kernel_id_of(cov_func)
    raw, mask, _ = params_utils.pack_raw(params.model, d, False, warp_func)
    ds = eng.pack([(0, xt, torch.zeros((n,), device=eng.device, dtype=eng.dtype))])
    chols, _, _, _ = eng.factorize(kid, 0, ds, raw, mask, want_chol=True,
                                   want_alpha=False)
    return mean + chols[0] @ z
  _, cov = linalg.compute_delta_y_and_cov(
      mean_func, cov_func, params, x, torch.zeros((n, 1)), warp_func, eps)
  return mean + torch.linalg.cholesky(cov) @ z


def predict(mean_func,
            cov_func,
            params,
            x_observed,
            y_observed,
            x_query,
            warp_func=None,
            full_cov=False,
            cache=None,
            _noise_flag=0.0,
            _var_scale=1.0):
  """Predict the GP at x_query conditioned on observations (gp.py:242-305).

  Returns (mu (nq,1), var (nq,1)) or (mu, cov (nq,nq)) if full_cov.
  """
  kid = _kernel.kernel_id_of(cov_func)
  mid = _mean.mean_id_of(mean_func)
  eng = _engine.Engine.get()
  x_query = eng.tensor(x_query)
  d = x_query.shape[1]
  raw, mask, _ = params_utils.pack_raw(params.model, d, mid == 1, warp_func)
  if x_observed is None or torch.as_tensor(x_observed).shape[0] == 0:
    # prior (gp.py:275-282)
    if full_cov:
      return eng.predict_cov(kid, mid, None, None, raw, mask, x_query,
                             noise_flag=_noise_flag, var_scale=_var_scale)
    mu, var, _ = eng.predict(kid, mid, None, None, raw, mask, x_query,
                             noise_flag=_noise_flag, var_scale=_var_scale)
    return mu, var
  x_observed = eng.tensor(x_observed)
  packed = getattr(cache, "packed", None) if cache is not None else None
  if packed is None:
    _, _, _, packed = linalg.solve_gp_linear_system(
        mean_func, cov_func, params, x_observed, y_observed, warp_func,
        return_cache=True)
  if not full_cov:
    mu, var, _ = eng.predict(kid, mid, x_observed, packed, raw, mask, x_query,
                             noise_flag=_noise_flag, var_scale=_var_scale)
    return mu, var
  # full covariance (gp.py:295-300): V = L^-1 K* and K** - V'V on the tensor pipe
  # (hb_predict_cov), noise / N/(N-1) of GP.predict fused into the epilogue
  return eng.predict_cov(kid, mid, x_observed, packed, raw, mask, x_query,
                         noise_flag=_noise_flag, var_scale=_var_scale)


class GP:
  """A Gaussian process that supports learning with historical data
  (gp.py:308-620).  Same attributes and method semantics as the reference."""
  dataset: Dict[Union[int, str], SubDataset]

  def __init__(self, dataset, mean_func, cov_func, params, warp_func=None):
    self.mean_func = mean_func
    self.cov_func = cov_func
    self.params = params if params is not None else GPParams()
    self.warp_func = warp_func
    self.set_dataset(dataset)
    if "objective" not in self.params.config:
      self.params.config["objective"] = obj.neg_log_marginal_likelihood
    self.rng = None

  # ---- dataset / cache bookkeeping (gp.py:328-346,403-452,535-538) --------
  @staticmethod
  def _arr(a):
    if not isinstance(a, torch.Tensor):
      a = np.asarray(a, dtype=np.float64)
    t = torch.as_tensor(a)
    if t.dtype != torch.float64:
      t = t.to(torch.float64)
    if torch.cuda.is_available() and not t.is_cuda:
      t = t.cuda()
    return t

  def initialize_params(self, key):
    """Initialize params (gp.py:347-401): a float lengthscale is broadcast to
    ones(d) * l -- that is how ARD is switched on."""
    if not self.dataset:
      raise ValueError("Cannot initialize GPParams without dataset.")
    if logging.getLogger().isEnabledFor(logging.INFO):  # gp.py:352
      data_utils.log_dataset(self.dataset)
    if isinstance(self.params.config["objective"], str):
      self.params.config["objective"] = getattr(
          obj, self.params.config["objective"])
    name = getattr(self.mean_func, "__name__", "") + getattr(
        self.cov_func, "__name__", "")
    if "mlp" in name or "linear" in name:
      raise NotImplementedError("MLP / linear bases are outside the hot path")
    ls = self.params.model.get("lengthscale", None)
    if isinstance(ls, float):
      self.params.model["lengthscale"] = np.ones(self.input_dim) * ls
    self.rng = key

  def set_dataset(self, dataset):
    """Reset GP dataset (gp.py:403-419)."""
    self.dataset = {}
    self.params.cache = {}
    if isinstance(dataset, list):
      dataset = {i: dataset[i] for i in range(len(dataset))}
    items = [(key, SubDataset(*val)) for key, val in dataset.items()]
    # host arrays travel in ONE upload (a 256-task dataset is 512 arrays); every
    # x / y is then a view into that device buffer
    host = [(i, f, np.asarray(a, dtype=np.float64))
            for i, (_, val) in enumerate(items) for f, a in ((0, val.x), (1, val.y))
            if not isinstance(a, torch.Tensor)]
    dev = {}
    if len(host) > 2 and torch.cuda.is_available():
      flat = torch.from_numpy(
          np.concatenate([a.reshape(-1) for _, _, a in host])).cuda()
      off = 0
      for i, f, a in host:
        dev[(i, f)] = flat[off:off + a.size].view(a.shape)
        off += a.size
    for i, (key, val) in enumerate(items):
      x = dev[(i, 0)] if (i, 0) in dev else self._arr(val.x)
      y = dev[(i, 1)] if (i, 1) in dev else self._arr(val.y)
      self.dataset[key] = SubDataset(x, y, val.aligned)

  @property
  def input_dim(self) -> int:
    key = list(self.dataset.keys())[0]
    return self.dataset[key].x.shape[1]

  def update_sub_dataset(self, sub_dataset, sub_dataset_key=0, is_append=False):
    """Update a sub-dataset (gp.py:426-452)."""
    sub_dataset = SubDataset(*sub_dataset)
    x = self._arr(sub_dataset.x)
    y = self._arr(sub_dataset.y)
    if is_append:
      if sub_dataset_key not in self.dataset:
        assert self.dataset, "dataset cannot be empty."
        any_x = next(iter(self.dataset.values())).x
        self.dataset[sub_dataset_key] = SubDataset(
            x=torch.empty((0, self.input_dim), dtype=torch.float64,
                          device=any_x.device),
            y=torch.empty((0, 1), dtype=torch.float64, device=any_x.device))
      cur = self.dataset[sub_dataset_key]
      new_x = torch.vstack((cur.x, x.reshape(-1, cur.x.shape[1]).to(cur.x.device)))
      new_y = torch.vstack((cur.y, y.reshape(-1, cur.y.shape[1]).to(cur.y.device)))
      self.dataset[sub_dataset_key] = SubDataset(x=new_x, y=new_y)
    else:
      self.dataset[sub_dataset_key] = SubDataset(x, y, sub_dataset.aligned)
    if sub_dataset_key in self.params.cache:
      self.params.cache[sub_dataset_key].needs_update = True

  def update_model_params(self, model_params: Dict[str, Any]):
    """Update params.model (must clean up params.cache) (gp.py:535-538)."""
    self.params.model = model_params
    self.params.cache = {}

  # ---- training (gp.py:454-485) ------------------------------------------
  def train(self, key=None, get_params_path=None, callback=None) -> GPParams:
    if key is None:
      if self.rng is None:
        self.rng = 0
        logging.info("Using default random state in GP.train.")
      if isinstance(self.rng, torch.Generator):
        subkey = self.rng
      else:
        subkey = int(self.rng)
        self.rng = int(self.rng) + 1
    else:
      subkey = key
    self.params = infer_parameters(
        mean_func=self.mean_func,
        cov_func=self.cov_func,
        init_params=self.params,
        dataset=self.dataset,
        warp_func=self.warp_func,
        objective=self.params.config["objective"],
        key=subkey,
        get_params_path=get_params_path,
        callback=callback)
    logging.info(msg=f"params = {self.params}")
    return self.params

  def neg_log_marginal_likelihood(self, use_cholesky=True):
    """Total nll and dict key -> nll (gp.py:487-497).  The reference evaluates
    this with its SVD branch (use_cholesky=False: Gram matrix from the engine,
    SVD from cuSOLVER); the default here is the engine's Cholesky branch, which
    the reference's own test pins to agree (objectives_test.py:298-301)."""
    return obj.neg_log_marginal_likelihood(
        mean_func=self.mean_func,
        cov_func=self.cov_func,
        params=self.params,
        dataset=self.dataset,
        warp_func=self.warp_func,
        return_key2nll=True,
        use_cholesky=use_cholesky)

  def empirical_divergence(self, distance=obj._utils.kl_multivariate_normal):
    """Empirical divergence from sample mean / covariance (gp.py:499-509)."""
    return obj.multivariate_normal_divergence(
        mean_func=self.mean_func,
        cov_func=self.cov_func,
        params=self.params,
        dataset=self.dataset,
        warp_func=self.warp_func,
        distance=distance)

  def stats(self, verbose=True):
    """Objective stats of the current model (gp.py:511-533):
    (nll, ekl, ekl_partial, euc, key2nll)."""
    import functools
    nll, key2nll = self.neg_log_marginal_likelihood(use_cholesky=False)
    kl = obj._utils.kl_multivariate_normal
    ekl = self.empirical_divergence(
        distance=functools.partial(kl, eps=1e-6, partial=False))
    ekl_partial = self.empirical_divergence(
        distance=functools.partial(kl, eps=1e-6, partial=True))
    euc = self.empirical_divergence(
        distance=obj._utils.euclidean_multivariate_normal)
    nll, ekl, ekl_partial, euc = (float(v) for v in (nll, ekl, ekl_partial, euc))
    msg = f"nll = {nll}, ekl = {ekl}, ekl_partial = {ekl_partial}, euc = {euc}"
    if verbose:
      print(msg)
    logging.info(msg=msg)
    return nll, ekl, ekl_partial, euc, {k: float(v) for k, v in key2nll.items()}

  # ---- prediction (gp.py:540-620) ----------------------------------------
  def setup_predictor(self, sub_dataset_key=0):
    """Compute and cache (chol, kinvy) for a sub-dataset (gp.py:540-560)."""
    if sub_dataset_key in self.params.cache and not self.params.cache[
        sub_dataset_key].needs_update:
      return
    chol, kinvy, _, packed = linalg.solve_gp_linear_system(
        mean_func=self.mean_func,
        cov_func=self.cov_func,
        params=self.params,
        x=self.dataset[sub_dataset_key].x,
        y=self.dataset[sub_dataset_key].y,
        warp_func=self.warp_func,
        return_cache=True)
    self.params.cache[sub_dataset_key] = GPCache(
        chol=chol, kinvy=kinvy, needs_update=False, packed=packed)

  def _noise_and_scale(self, with_noise, unbiased):
    noise_flag = 1.0 if with_noise else 0.0
    scale = 1.0
    if unbiased:
      len_dataset = len(
          [k for k, v in self.dataset.items() if v.aligned is None])
      if len_dataset > 1:
        scale = len_dataset / (len_dataset - 1.0)
    return noise_flag, scale

  def predict(self, queried_inputs, sub_dataset_key=0, full_cov=False,
              with_noise=True, unbiased=True):
    """Predict mean and (co)variance at queried_inputs (gp.py:562-620): noise
    is added WITHOUT the 1e-6 jitter, then the N/(N-1) inflation."""
    noise_flag, scale = self._noise_and_scale(with_noise, unbiased)
    has_obs = (sub_dataset_key in self.dataset and
               self.dataset[sub_dataset_key].x.shape[0] > 0)
    if sub_dataset_key in self.dataset:
      if has_obs:
        self.setup_predictor(sub_dataset_key)
      x_obs = self.dataset[sub_dataset_key].x
      y_obs = self.dataset[sub_dataset_key].y
      cache = self.params.cache.get(sub_dataset_key) if has_obs else None
    else:
      x_obs = y_obs = cache = None
    return predict(self.mean_func, self.cov_func, self.params, x_obs, y_obs,
                   queried_inputs, self.warp_func, full_cov, cache,
                   _noise_flag=noise_flag, _var_scale=scale)

  def engine_ids(self, d: int):
    """(kernel_id, mean_id, raw, warp_mask) of this model for the C ABI."""
    kid = _kernel.kernel_id_of(self.cov_func)
    mid = _mean.mean_id_of(self.mean_func)
    raw, mask, _ = params_utils.pack_raw(self.params.model, d, mid == 1,
                                         self.warp_func)
    return kid, mid, raw, mask

  def acquisition(self, queried_inputs, sub_dataset_key, acq_id, acq_param):
    """Fused predict + acfun_sub epilogue (acfun.py:84-88 + :96-142) with the
    GP.predict conventions full_cov=False, with_noise=True, unbiased=True."""
    noise_flag, scale = self._noise_and_scale(True, True)
    kid = _kernel.kernel_id_of(self.cov_func)
    mid = _mean.mean_id_of(self.mean_func)
    eng = _engine.Engine.get()
    xq = eng.tensor(queried_inputs)
    raw, mask, _ = params_utils.pack_raw(self.params.model, xq.shape[1],
                                         mid == 1, self.warp_func)
    has_obs = (sub_dataset_key in self.dataset and
               self.dataset[sub_dataset_key].x.shape[0] > 0)
    if has_obs:
      self.setup_predictor(sub_dataset_key)
      x_obs = self.dataset[sub_dataset_key].x
      packed = self.params.cache[sub_dataset_key].packed
    else:
      x_obs, packed = None, None
    _, _, acq = eng.predict(kid, mid, x_obs, packed, raw, mask, xq,
                            noise_flag=noise_flag, var_scale=scale,
                            acq_id=acq_id, acq_param=acq_param, want_mu=False,
                            want_var=False)
    return acq


class HGP(GP):
  """Hierarchical GP over hyperparameter samples (gp.py:623-682)."""

  def get_model_params_samples(self):
    return self.params.samples if self.params.samples else [self.params.model]

  def stats(self, verbose=True):
    """Objective stats averaged over the hyper-parameter samples (gp.py:634-664)."""
    import collections
    samples = self.get_model_params_samples()
    all_stats, all_key2nll, key2nll = [], collections.defaultdict(float), {}
    for model_params in samples:
      self.update_model_params(model_params)
      nll, ekl, ekl_partial, euc, key2nll = super().stats(verbose=verbose)
      all_stats.append((nll, ekl, ekl_partial, euc))
      for k in key2nll:
        all_key2nll[k] += key2nll[k]
    for k in key2nll:
      all_key2nll[k] /= len(samples)
    nll, ekl, ekl_partial, euc = (float(v) for v in np.mean(
        np.asarray(all_stats, dtype=np.float64), axis=0))
    msg = (f"HGP nll = {nll}, ekl = {ekl}, ekl_partial = {ekl_partial}, euc ="
           f" {euc}")
    if verbose:
      print(msg)
    logging.info(msg=msg)
    return nll, ekl, ekl_partial, euc, all_key2nll

  def predict(self, queried_inputs, sub_dataset_key=0, full_cov=False,
              with_noise=True):
    samples = self.get_model_params_samples()
    eng = _engine.Engine.get()
    has_obs = (sub_dataset_key in self.dataset and
               self.dataset[sub_dataset_key].x.shape[0] > 0)
    if (not full_cov and has_obs and len(samples) > 1 and
        getattr(eng, "h", None) is not None):
      # second batch axis: the S samples' factorisations are ONE launch sequence
      # (hb_build_predictors_multi); the per-sample predict sweeps reuse it
      xq = eng.tensor(queried_inputs)
      kid = _kernel.kernel_id_of(self.cov_func)
      mid = _mean.mean_id_of(self.mean_func)
      packs = [params_utils.pack_raw(m, xq.shape[1], mid == 1, self.warp_func)
               for m in samples]
      if len({p[1] for p in packs}) == 1:
        mask = packs[0][1]
        raws = np.stack([p[0] for p in packs])
        sd = self.dataset[sub_dataset_key]
        caches, _, _ = eng.build_predictors_multi(kid, mid, sd.x, sd.y, raws, mask)
        noise_flag, scale = self._noise_and_scale(with_noise, True)
        x_obs = eng.tensor(sd.x)
        out = []
        for s_ in range(len(samples)):
          mu, var, _ = eng.predict(kid, mid, x_obs, caches[s_], raws[s_], mask, xq,
                                   noise_flag=noise_flag, var_scale=scale)
          out.append((mu, var))
        return out
    results = []
    for model_params in samples:
      self.update_model_params(model_params)
      results.append(super().predict(
          queried_inputs=queried_inputs, sub_dataset_key=sub_dataset_key,
          full_cov=full_cov, with_noise=with_noise))
    return results
