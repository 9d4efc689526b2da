"""Host-side driver of the C-ABI engine: device memory and streams come from
PyTorch (plumbing); all arithmetic happens in libhyperbo_b200.so.

There is NO CPU fallback: constructing an Engine without the CUDA extension or
without a CUDA device raises.
"""
from __future__ import annotations

import threading
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

KERNEL_IDS = {"squared_exponential": 0, "matern32": 1, "matern52": 2}
MEAN_IDS = {"zero": 0, "constant": 1}
ACQ_IDS = {"none": 0, "ei": 1, "pi": 2, "ucb": 3}
DTYPES = {torch.float64: 0, torch.float32: 1}
JITTER = 1e-6  # basics/linalg.py:42 of the reference

_lock = threading.Lock()
_engines: Dict[Tuple[int, torch.dtype], "Engine"] = {}
_default_dtype = torch.float64


def set_default_dtype(dtype) -> None:
  """Engine precision used by the API mirror (gp_utils / bo_utils / basics):
  torch.float64 (default; fp64 DMMA tile products) or torch.float32 (the
  reference's JAX default; 3xTF32 tile products)."""
  global _default_dtype
  if dtype not in DTYPES:
    raise ValueError("engine dtype must be torch.float64 or torch.float32")
  _default_dtype = dtype


def get_default_dtype():
  return _default_dtype


def _load_ext():
  try:
    from hyperbo_b200 import _C  # type: ignore
  except ImportError as e:  # fail loudly: no silent eager fallback
    raise RuntimeError(
        "hyperbo_b200._C (the CUDA extension) is not built; run "
        "`python -c 'import __graft_entry__ as g; g.build()'` or "
        "`python hyperbo_b200/_build.py`") from e
  return _C


class PackedDataset:
  """Ragged task batch in the C-ABI layout: rows of all non-empty, non-aligned
  tasks concatenated; offs = prefix sums (host list)."""

  def __init__(self, keys, x: torch.Tensor, y: torch.Tensor, offs: List[int]):
    self.keys = list(keys)
    self.x = x
    self.y = y
    self.offs = [int(o) for o in offs]

  @property
  def num_tasks(self) -> int:
    return len(self.offs) - 1

  @property
  def d(self) -> int:
    return int(self.x.shape[1])


class Engine:
  """One C-ABI handle per (device, dtype)."""

  def __init__(self, device: Optional[int] = None, dtype=None):
    dtype = dtype or _default_dtype
    if not torch.cuda.is_available():
      raise RuntimeError("hyperbo_b200 needs a CUDA device (no CPU fallback)")
    self._C = _load_ext()
    if device is None:
      device = torch.cuda.current_device()
    self.device_index = int(device)
    self.device = torch.device("cuda", self.device_index)
    self.dtype = dtype
    with torch.cuda.device(self.device):
      self.h = self._C.Handle(self.device_index, DTYPES[dtype])
    self.max_dim = self._C.MAX_DIM

  # ------------------------------------------------------------- helpers --
  @staticmethod
  def get(device: Optional[int] = None, dtype=None) -> "Engine":
    dtype = dtype or _default_dtype
    if device is None:
      if not torch.cuda.is_available():
        raise RuntimeError("hyperbo_b200 needs a CUDA device (no CPU fallback)")
      device = torch.cuda.current_device()
    key = (int(device), dtype)
    with _lock:
      if key not in _engines:
        _engines[key] = Engine(device, dtype)
      return _engines[key]

  def _stream(self) -> int:
    return torch.cuda.current_stream(self.device).cuda_stream

  def tensor(self, a, shape=None) -> torch.Tensor:
    if not isinstance(a, torch.Tensor):  # python floats/lists must not round
      a = np.asarray(a, dtype=np.float64)  # through torch's float32 default
    t = torch.as_tensor(a)
    t = t.to(device=self.device, dtype=self.dtype)
    if shape is not None:
      t = t.reshape(shape)
    return t.contiguous()

  def launch_count(self) -> int:
    return int(self.h.launch_count())

  def workspace_bytes(self) -> int:
    return int(self.h.workspace_bytes())

  def _check_dim(self, d):
    if d > self.max_dim:
      raise NotImplementedError(
          f"input dimension {d} > {self.max_dim} is not supported by the engine")

  # --------------------------------------------------------------- packing --
  def pack(self, tasks: Sequence[Tuple[object, torch.Tensor, torch.Tensor]]
           ) -> PackedDataset:
    """tasks: iterable of (key, x (n,d), y (n,1) or (n,)).  Empty tasks are
    dropped (objectives.py:184 of the reference)."""
    keys, xs, ys, offs = [], [], [], [0]
    tasks = list(tasks)
    if tasks and all(not isinstance(x, torch.Tensor) and not isinstance(y, torch.Tensor)
                     for _, x, y in tasks):
      # host arrays: concatenate on the host, ONE upload per array (a batch of
      # 256 tasks would otherwise be 512 small host->device copies)
      hx, hy = [], []
      for k, x, y in tasks:
        x = np.asarray(x, dtype=np.float64)
        if x.shape[0] == 0:
          continue
        y = np.asarray(y, dtype=np.float64).reshape(-1)
        if y.shape[0] != x.shape[0]:
          raise ValueError(f"dataset[{k}].x has shape {tuple(x.shape)} but y has "
                           f"{y.shape[0]} rows")
        keys.append(k)
        hx.append(x)
        hy.append(y)
        offs.append(offs[-1] + x.shape[0])
      if hx:
        x = self.tensor(np.concatenate(hx, 0))
        self._check_dim(x.shape[1])
        return PackedDataset(keys, x, self.tensor(np.concatenate(hy, 0)), offs)
      tasks = []
    for k, x, y in tasks:
      x = self.tensor(x)
      if x.shape[0] == 0:
        continue
      y = self.tensor(y).reshape(-1)
      if y.shape[0] != x.shape[0]:
        raise ValueError(f"dataset[{k}].x has shape {tuple(x.shape)} but y has "
                         f"{y.shape[0]} rows")
      keys.append(k)
      xs.append(x)
      ys.append(y)
      offs.append(offs[-1] + x.shape[0])
    if not xs:
      return PackedDataset([], torch.zeros((0, 1), device=self.device,
                                           dtype=self.dtype),
                           torch.zeros((0,), device=self.device,
                                       dtype=self.dtype), [0])
    x = torch.cat(xs, 0).contiguous()
    self._check_dim(x.shape[1])
    return PackedDataset(keys, x, torch.cat(ys, 0).contiguous(), offs)

  # ------------------------------------------------------------ operations --
  def kernel_matrix(self, kernel_id: int, x1, x2, raw, mask: int, diag=False,
                    add_noise=False, jitter=JITTER) -> torch.Tensor:
    x1 = self.tensor(x1)
    n1, d = x1.shape
    self._check_dim(d)
    raw = self.tensor(raw)
    if x2 is None:
      if diag:
        out = torch.empty((n1,), device=self.device, dtype=self.dtype)
      else:
        out = torch.empty((n1, n1), device=self.device, dtype=self.dtype)
      self.h.kernel_matrix(kernel_id, x1.data_ptr(), n1, 0, 0, d,
                           raw.data_ptr(), mask, int(diag), int(add_noise),
                           float(jitter), out.data_ptr(), self._stream())
      return out
    x2 = self.tensor(x2)
    n2 = x2.shape[0]
    out = torch.empty((n1, n2), device=self.device, dtype=self.dtype)
    self.h.kernel_matrix(kernel_id, x1.data_ptr(), n1, x2.data_ptr(), n2, d,
                         raw.data_ptr(), mask, 0, 0, float(jitter),
                         out.data_ptr(), self._stream())
    return out

  def factorize(self, kernel_id: int, mean_id: int, ds: PackedDataset, raw,
                mask: int, want_chol=True, want_alpha=True):
    """-> (chol list or None, alpha (sum n,) or None, nll (T,), info (T,))."""
    raw = self.tensor(raw)
    T = ds.num_tasks
    ns = [ds.offs[t + 1] - ds.offs[t] for t in range(T)]
    chol = torch.empty((sum(n * n for n in ns),), device=self.device,
                       dtype=self.dtype) if want_chol else None
    alpha = torch.empty((ds.offs[-1],), device=self.device,
                        dtype=self.dtype) if want_alpha else None
    nll = torch.zeros((max(T, 1),), device=self.device, dtype=self.dtype)
    info = torch.zeros((max(T, 1),), device=self.device, dtype=torch.int32)
    self.h.factorize_batched(
        kernel_id, mean_id, ds.offs, ds.d, ds.x.data_ptr(), ds.y.data_ptr(),
        raw.data_ptr(), mask, chol.data_ptr() if want_chol else 0,
        alpha.data_ptr() if want_alpha else 0, nll.data_ptr(), info.data_ptr(),
        self._stream())
    chols = None
    if want_chol:
      chols, o = [], 0
      for n in ns:
        chols.append(chol[o:o + n * n].view(n, n))
        o += n * n
    return chols, alpha, nll[:T], info[:T]

  def nll_grad(self, kernel_id: int, mean_id: int, ds: PackedDataset, raw,
               mask: int, sums_out: Optional[torch.Tensor] = None,
               want_task_nll=False, weights: Optional[torch.Tensor] = None,
               jitter: Optional[float] = None):
    """-> sums (P+2,) = [sum w nll, sum w d/d raw_p ..., count] (+ per-task
    nll).  `weights` (T,) and `jitter` select hb_nll_grad_weighted (the building
    block of the KL objectives); without them it is hb_nll_grad_batched."""
    raw = self.tensor(raw)
    P = 3 + ds.d
    if sums_out is None:
      sums_out = torch.empty((P + 2,), device=self.device, dtype=self.dtype)
    T = ds.num_tasks
    nll_task = torch.zeros((max(T, 1),), device=self.device,
                           dtype=self.dtype) if want_task_nll else None
    if weights is not None or jitter is not None:
      if weights is not None:
        weights = self.tensor(weights).reshape(-1)
        if weights.shape[0] != T:
          raise ValueError(f"weights has {weights.shape[0]} entries for {T} tasks")
      self.h.nll_grad_weighted(
          kernel_id, mean_id, ds.offs, ds.d, ds.x.data_ptr(), ds.y.data_ptr(),
          raw.data_ptr(), mask, weights.data_ptr() if weights is not None else 0,
          JITTER if jitter is None else float(jitter), sums_out.data_ptr(),
          nll_task.data_ptr() if want_task_nll else 0, 0, self._stream())
      if want_task_nll:
        return sums_out, nll_task[:T]
      return sums_out
    self.h.nll_grad_batched(
        kernel_id, mean_id, ds.offs, ds.d, ds.x.data_ptr(), ds.y.data_ptr(),
        raw.data_ptr(), mask, sums_out.data_ptr(),
        nll_task.data_ptr() if want_task_nll else 0, 0, self._stream())
    if want_task_nll:
      return sums_out, nll_task[:T]
    return sums_out

  def nll_grad_mrhs(self, kernel_id: int, mean_id: int, ds: PackedDataset, R: int,
                    B: torch.Tensor, col_weight: torch.Tensor,
                    col_mean: Optional[torch.Tensor], raw, mask: int,
                    weights: Optional[torch.Tensor] = None,
                    jitter: Optional[float] = None,
                    sums_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """hb_nll_grad_mrhs: R right-hand-side columns per task on ONE factorisation
    of each task's inputs (ds.y is not read).  B: flat, task t owns the (R, n_t)
    block at offs[t] * R with every column contiguous; col_weight (T, R);
    col_mean (R,) int32 (1: subtract the model mean from that column)."""
    raw = self.tensor(raw)
    T = ds.num_tasks
    if sums_out is None:
      sums_out = torch.empty((3 + ds.d + 2,), device=self.device, dtype=self.dtype)
    B = self.tensor(B).reshape(-1)
    col_weight = self.tensor(col_weight).reshape(-1)
    if B.shape[0] != ds.offs[-1] * R or col_weight.shape[0] != T * R:
      raise ValueError("B / col_weight do not match the (tasks, R) layout")
    if col_mean is not None:
      col_mean = torch.as_tensor(col_mean, device=self.device).to(torch.int32)
      if col_mean.shape[0] != R:
        raise ValueError("col_mean needs R entries")
    if weights is not None:
      weights = self.tensor(weights).reshape(-1)
      if weights.shape[0] != T:
        raise ValueError(f"weights has {weights.shape[0]} entries for {T} tasks")
    self.h.nll_grad_mrhs(
        kernel_id, mean_id, ds.offs, ds.d, ds.x.data_ptr(), R, B.data_ptr(),
        col_weight.data_ptr(), col_mean.data_ptr() if col_mean is not None else 0,
        raw.data_ptr(), mask, weights.data_ptr() if weights is not None else 0,
        JITTER if jitter is None else float(jitter), sums_out.data_ptr(), 0,
        self._stream())
    return sums_out

  def euclid_grad(self, kernel_id: int, mean_id: int, ds: PackedDataset, R: int,
                  Yc: torch.Tensor, mu0: torch.Tensor, raw, mask: int,
                  mean_weight: float = 1.0, cov_weight: float = 1.0,
                  weights: Optional[torch.Tensor] = None,
                  sums_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """hb_euclid_grad: sum_t w_t (mean_weight ||mu0 - m(x)|| + cov_weight
    ||Yc Yc' - (K + nv I)||_F) and its raw-parameter gradient.  Yc: flat, task t
    owns the (R, n_t) block at offs[t] * R (columns contiguous, already divided by
    sqrt(m)); mu0: (sum n,)."""
    raw = self.tensor(raw)
    T = ds.num_tasks
    if sums_out is None:
      sums_out = torch.empty((3 + ds.d + 2,), device=self.device, dtype=self.dtype)
    Yc = self.tensor(Yc).reshape(-1)
    mu0 = self.tensor(mu0).reshape(-1)
    if Yc.shape[0] != ds.offs[-1] * R or mu0.shape[0] != ds.offs[-1]:
      raise ValueError("Yc / mu0 do not match the (tasks, R) layout")
    if weights is not None:
      weights = self.tensor(weights).reshape(-1)
      if weights.shape[0] != T:
        raise ValueError(f"weights has {weights.shape[0]} entries for {T} tasks")
    self.h.euclid_grad(kernel_id, mean_id, ds.offs, ds.d, ds.x.data_ptr(), R,
                       Yc.data_ptr(), mu0.data_ptr(), raw.data_ptr(), mask,
                       float(mean_weight), float(cov_weight),
                       weights.data_ptr() if weights is not None else 0,
                       sums_out.data_ptr(), self._stream())
    return sums_out

  def generation(self) -> int:
    """Bumped whenever a workspace buffer / cached plan of the handle moves:
    CUDA graphs that captured engine calls must be re-captured then."""
    return int(self.h.generation())

  # ---- peer-memory all-reduce (hb_comm_*) -----------------------
  def comm_init(self) -> bool:
    """Set up the NVLink peer-memory all-reduce between the ranks of the default
    torch.distributed group (one process per GPU of ONE node).  Returns False
    when it cannot be used (single rank, no CUDA IPC / peer access); the caller
    then falls back to torch.distributed.all_reduce."""
    import torch.distributed as dist
    if getattr(self, "_comm_ready", None) is not None:
      return self._comm_ready
    self._comm_ready = False
    if getattr(self, "h", None) is None or self.device.type != "cuda":
      return False  # (an engine without a C-ABI handle: the CPU test double)
    if not (dist.is_available() and dist.is_initialized()):
      return False
    world, rank = dist.get_world_size(), dist.get_rank()
    if world < 2 or world > 64:
      return False
    ok = 1
    try:
      handle = self.h.comm_export()
    except RuntimeError:
      handle, ok = b"\0" * 64, 0
    gathered = [None] * world
    dist.all_gather_object(gathered, (ok, handle))
    if all(g[0] for g in gathered):
      try:
        self.h.comm_import(rank, world, b"".join(g[1] for g in gathered))
      except RuntimeError:
        ok = 0
    else:
      ok = 0
    flags = [None] * world
    dist.all_gather_object(flags, ok)  # (also the barrier hb_comm_import needs)
    self._comm_ready = all(flags)
    return self._comm_ready

  def allreduce(self, buf: torch.Tensor):
    """In-place deterministic all-reduce(sum) of <= 64 scalars (hb_allreduce)."""
    self.h.allreduce(buf.data_ptr(), int(buf.numel()), self._stream())

  def allreduce_adam_step(self, P: int, raw, m, v, accepted, sums, scal, lr,
                          b1=0.9, b2=0.999, eps=1e-8, tie_lengthscale=False):
    self.h.allreduce_adam_step(P, raw.data_ptr(), m.data_ptr(), v.data_ptr(),
                               accepted.data_ptr(), sums.data_ptr(),
                               scal.data_ptr(), float(lr), float(b1), float(b2),
                               float(eps), int(bool(tie_lengthscale)),
                               self._stream())

  # ---- second batch axis: S hyper-parameter sets x the same data -----------
  def nll_grad_multi(self, kernel_id: int, mean_id: int, ds: PackedDataset, raws,
                     mask: int, want_task_nll=False):
    """raws (S, P) -> sums (S, P+2) [+ per-task nll (S, T)]: hb_nll_grad_multi,
    ONE launch sequence for all S parameter sets."""
    raws = self.tensor(raws)
    S, P = int(raws.shape[0]), 3 + ds.d
    if raws.shape[1] != P:
      raise ValueError(f"raws must have {P} columns")
    sums = torch.empty((S, P + 2), device=self.device, dtype=self.dtype)
    T = ds.num_tasks
    nll_task = torch.zeros((S, max(T, 1)), device=self.device,
                           dtype=self.dtype) if want_task_nll else None
    self.h.nll_grad_multi(kernel_id, mean_id, S, ds.offs, ds.d, ds.x.data_ptr(),
                          ds.y.data_ptr(), raws.data_ptr(), mask, sums.data_ptr(),
                          nll_task.data_ptr() if want_task_nll else 0, self._stream())
    return (sums, nll_task[:, :T]) if want_task_nll else sums

  def build_predictors_multi(self, kernel_id: int, mean_id: int, x, y, raws, mask: int):
    """S predictor caches of one task from ONE factorisation launch
    (hb_build_predictors_multi) -> (caches (S, stride) uint8, nll (S,), info (S,))."""
    x = self.tensor(x)
    n, d = x.shape
    self._check_dim(d)
    y = self.tensor(y).reshape(-1)
    raws = self.tensor(raws)
    S = int(raws.shape[0])
    stride = (int(self.h.predictor_bytes(n)) + 255) // 256 * 256
    caches = torch.empty((S, stride), device=self.device, dtype=torch.uint8)
    nll = torch.zeros((S,), device=self.device, dtype=self.dtype)
    info = torch.zeros((S,), device=self.device, dtype=torch.int32)
    self.h.build_predictors_multi(kernel_id, mean_id, S, n, d, x.data_ptr(), y.data_ptr(),
                                  raws.data_ptr(), mask, caches.data_ptr(), stride,
                                  nll.data_ptr(), info.data_ptr(), self._stream())
    return caches, nll, info

  def adam_step(self, P: int, raw, m, v, accepted, sums, scal, lr, b1=0.9,
                b2=0.999, eps=1e-8, tie_lengthscale=False):
    self.h.adam_step(P, raw.data_ptr(), m.data_ptr(), v.data_ptr(),
                     accepted.data_ptr(), sums.data_ptr(), scal.data_ptr(),
                     float(lr), float(b1), float(b2), float(eps),
                     int(bool(tie_lengthscale)), self._stream())

  def build_predictor(self, kernel_id: int, mean_id: int, x, y, raw, mask: int):
    """-> (cache bytes tensor, chol (n,n), kinvy (n,1), nll scalar, info)."""
    x = self.tensor(x)
    n, d = x.shape
    self._check_dim(d)
    y = self.tensor(y).reshape(-1)
    raw = self.tensor(raw)
    nbytes = int(self.h.predictor_bytes(n))
    cache = torch.empty((nbytes,), device=self.device, dtype=torch.uint8)
    chol = torch.empty((n, n), device=self.device, dtype=self.dtype)
    kinvy = torch.empty((n, 1), device=self.device, dtype=self.dtype)
    nll = torch.zeros((1,), device=self.device, dtype=self.dtype)
    info = torch.zeros((1,), device=self.device, dtype=torch.int32)
    self.h.build_predictor(kernel_id, mean_id, n, d, x.data_ptr(), y.data_ptr(),
                           raw.data_ptr(), mask, cache.data_ptr(),
                           chol.data_ptr(), kinvy.data_ptr(), nll.data_ptr(),
                           info.data_ptr(), self._stream())
    return cache, chol, kinvy, nll, info

  def predict(self, kernel_id: int, mean_id: int, x, cache, raw, mask: int, xq,
              noise_flag=0.0, var_scale=1.0, acq_id=0, acq_param=0.0,
              want_mu=True, want_var=True):
    xq = self.tensor(xq)
    nq, d = xq.shape
    self._check_dim(d)
    raw = self.tensor(raw)
    if x is None or x.shape[0] == 0:
      n, xp, cp = 0, 0, 0
    else:
      x = self.tensor(x)
      n, xp, cp = x.shape[0], x.data_ptr(), cache.data_ptr()
    mu = torch.empty((nq, 1), device=self.device,
                     dtype=self.dtype) if want_mu else None
    var = torch.empty((nq, 1), device=self.device,
                      dtype=self.dtype) if want_var else None
    acq = torch.empty((nq, 1), device=self.device,
                      dtype=self.dtype) if acq_id else None
    self.h.predict(kernel_id, mean_id, n, d, xp, cp, raw.data_ptr(), mask, nq,
                   xq.data_ptr(), float(noise_flag), float(var_scale), acq_id,
                   float(acq_param), mu.data_ptr() if want_mu else 0,
                   var.data_ptr() if want_var else 0,
                   acq.data_ptr() if acq_id else 0, self._stream())
    return mu, var, acq

  def predict_cov(self, kernel_id: int, mean_id: int, x, cache, raw, mask: int, xq,
                  noise_flag=0.0, var_scale=1.0):
    """hb_predict_cov: (mu (nq,1), cov (nq,nq)) of gp.predict(full_cov=True),
    cov = (k(xq,xq) - V'V + noise_flag * noise_variance * I) * var_scale."""
    xq = self.tensor(xq)
    nq, d = xq.shape
    self._check_dim(d)
    raw = self.tensor(raw)
    if x is None or x.shape[0] == 0:
      n, xp, cp = 0, 0, 0
    else:
      x = self.tensor(x)
      n, xp, cp = x.shape[0], x.data_ptr(), cache.data_ptr()
    mu = torch.empty((nq, 1), device=self.device, dtype=self.dtype)
    cov = torch.empty((nq, nq), device=self.device, dtype=self.dtype)
    self.h.predict_cov(kernel_id, mean_id, n, d, xp, cp, raw.data_ptr(), mask, nq,
                       xq.data_ptr(), float(noise_flag), float(var_scale),
                       mu.data_ptr(), cov.data_ptr(), self._stream())
    return mu, cov

  def acquisition(self, acq_id: int, param: float, mu, var) -> torch.Tensor:
    mu = self.tensor(mu).reshape(-1)
    var = self.tensor(var).reshape(-1)
    out = torch.empty_like(mu)
    self.h.acquisition(acq_id, float(param), mu.shape[0], mu.data_ptr(),
                       var.data_ptr(), out.data_ptr(), self._stream())
    return out.reshape(-1, 1)


class DeviceSampler:
  """Per-step sub-sampling on the device (hb_subsample; data_utils.py:72-100):
  `dst` is a packed batch of FIXED shape -- task t keeps min(n_t, batch_size)
  rows if n_t >= batch_size, else all n_t -- that `sample()` refills from the
  full packed batch `src` with a uniform random subset per (seed, step, global
  task id).  With `scal` (the trainer's scalars, [1] = step counter) the call
  is CUDA-graph capturable: every replay draws a new sample."""

  def __init__(self, eng: "Engine", src: PackedDataset, batch_size: int, seed: int,
               task_ids: Optional[Sequence[int]] = None):
    self.eng, self.src, self.seed = eng, src, int(seed) & ((1 << 63) - 1)
    T = src.num_tasks
    ns = [src.offs[t + 1] - src.offs[t] for t in range(T)]
    nd = [batch_size if n >= batch_size else n for n in ns]
    offs_dst = [0]
    for n in nd:
      offs_dst.append(offs_dst[-1] + n)
    self.max_rows = max(nd) if nd else 0
    dev = eng.device
    self.offs_src_d = torch.tensor(src.offs, dtype=torch.int64, device=dev)
    self.offs_dst_d = torch.tensor(offs_dst, dtype=torch.int64, device=dev)
    ids = list(range(T)) if task_ids is None else [int(i) for i in task_ids]
    self.ids_d = torch.tensor(ids, dtype=torch.int64, device=dev)
    self.dst = PackedDataset(src.keys,
                             torch.empty((offs_dst[-1], src.d), device=dev, dtype=eng.dtype),
                             torch.empty((offs_dst[-1],), device=dev, dtype=eng.dtype),
                             offs_dst)
    self.dst.sampler = self

  def sample(self, step: int = 0, scal: Optional[torch.Tensor] = None):
    s, d_ = self.src, self.dst
    self.eng.h.subsample(s.num_tasks, s.d, self.offs_src_d.data_ptr(),
                         self.offs_dst_d.data_ptr(), self.ids_d.data_ptr(),
                         self.max_rows, s.x.data_ptr(), s.y.data_ptr(),
                         d_.x.data_ptr(), d_.y.data_ptr(), self.seed,
                         scal.data_ptr() if scal is not None else 0, int(step),
                         self.eng._stream())
    return d_


class BoSession:
  """Device-resident simulated-BO state of one queried task (hb_bo_init /
  hb_bo_step): observations, candidates and the packed inverse factor stay on
  the GPU; one `step` = acquisition over all candidates + arg-max + append of
  the chosen (x, y) + O(n^2) rank-1 update, with no host synchronisation."""

  def __init__(self, eng: Engine, kernel_id: int, mean_id: int, x0, y0, raw,
               mask: int, capacity: int, d: Optional[int] = None):
    self.eng, self.kid, self.mid, self.mask = eng, kernel_id, mean_id, mask
    x0 = eng.tensor(x0) if x0 is not None and len(x0) else None
    self.n = 0 if x0 is None else int(x0.shape[0])
    self.d = int(d if x0 is None else x0.shape[1])
    eng._check_dim(self.d)
    self.cap = int(capacity)
    if self.cap < self.n + 1:
      raise ValueError("capacity must exceed the number of initial observations")
    self.x = torch.zeros((self.cap, self.d), device=eng.device, dtype=eng.dtype)
    self.y = torch.zeros((self.cap,), device=eng.device, dtype=eng.dtype)
    if self.n:
      self.x[:self.n] = x0
      self.y[:self.n] = eng.tensor(y0).reshape(-1)
    self.raw = eng.tensor(raw)
    self.cache = torch.empty((int(eng.h.bo_cache_bytes(self.cap)),),
                             device=eng.device, dtype=torch.uint8)
    self.sel = torch.full((self.cap,), -1, device=eng.device, dtype=torch.int32)
    self.n0 = self.n
    eng.h.bo_init(kernel_id, mean_id, self.n, self.cap, self.d, self.x.data_ptr(),
                  self.y.data_ptr(), self.raw.data_ptr(), mask,
                  self.cache.data_ptr(), eng._stream())

  def step(self, xq: torch.Tensor, yq: torch.Tensor, acq_id: int, acq_param: float,
           target_is_ymax: bool, noise_flag=1.0, var_scale=1.0):
    """xq (nq, d), yq (nq,) device tensors of the engine's dtype."""
    if self.n + 1 > self.cap:
      raise ValueError("BoSession capacity exhausted")
    self.eng.h.bo_step(self.kid, self.mid, self.n, self.cap, self.d,
                       self.x.data_ptr(), self.y.data_ptr(), self.raw.data_ptr(),
                       self.mask, self.cache.data_ptr(), int(xq.shape[0]),
                       xq.data_ptr(), yq.data_ptr(), float(noise_flag),
                       float(var_scale), int(acq_id), float(acq_param),
                       int(bool(target_is_ymax)),
                       self.sel[self.n - self.n0:].data_ptr(), self.eng._stream())
    self.n += 1

  def selected(self) -> torch.Tensor:
    """Indices of the candidates chosen so far (one device->host read)."""
    return self.sel[:self.n - self.n0].cpu()

  def observations(self):
    return self.x[:self.n], self.y[:self.n].reshape(-1, 1)

  def predict(self, xq, noise_flag=0.0, var_scale=1.0):
    """Posterior at xq from the session's (appended) factor -- for tests."""
    eng = self.eng
    xq = eng.tensor(xq)
    # the growable layout keeps alpha at the capacity's offset: predict through a
    # compact copy (M tiles of the current n, then alpha)
    nblk = (self.n + 63) // 64
    es = 8 if eng.dtype == torch.float64 else 4
    cap_blk = (self.cap + 63) // 64
    mt = nblk * (nblk + 1) // 2 * 4096 * es
    al_off = cap_blk * (cap_blk + 1) // 2 * 4096 * es
    compact = torch.empty((mt + nblk * 64 * es + 256,), device=eng.device,
                          dtype=torch.uint8)
    compact[:mt] = self.cache[:mt]
    compact[mt:mt + nblk * 64 * es] = self.cache[al_off:al_off + nblk * 64 * es]
    return eng.predict(self.kid, self.mid, self.x[:self.n], compact, self.raw,
                       self.mask, xq, noise_flag=noise_flag, var_scale=var_scale)
