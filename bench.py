#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json):

  metric   : NLL pre-train steps/sec at 256 tasks x n=512 x d=8
  workload : configs[1] -- batched Cholesky + NLL(+grad), 256 x 512 x 8, fp64
  step     : mean-NLL and its gradient w.r.t. all P raw parameters over all T
             tasks + one Adam update (+ the all-reduce when task-sharded over
             N GPUs) + the loss read-back.

  python bench.py --gpus N --steps K --warmup W          (our arm)
  python bench.py --impl reference ...                   (CPU reference arm)

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_TASKS, N_PTS, DIM = 256, 512, 8
METRIC = "nll_pretrain_steps_per_sec_256tasks_n512_d8"
LR = 1e-3


def algorithmic_flops(T, n, d):
  """per-task n^3 [potrf n^3/3 + inverse 2n^3/3] + 4n^2 +
  n^2(3d+8) [kernel build] + n^2(2d+6) [gradient contraction]."""
  potrf = n**3 / 3.0
  trtri = n**3 / 3.0
  lauum = n**3 / 3.0
  build = n * n * (3 * d + 8)
  solves = 4.0 * n * n
  contract = n * n * (2 * d + 6)
  return {
      "step": T * (potrf + trtri + lauum + build + solves + contract),
      "factor_launches": T * (potrf + trtri + build + solves),
      "lauum_grad": T * (lauum + contract),
  }


def synthetic_batch(T, n, d, seed=0, tasks=None):
  """SURVEY.md 8(d) recipe (bo_utils/data.py:720-775 + gp.py:230-239): per task
  t, X_t ~ U[0,1]^{n x d} (PCG64 seed 1000+t) and y_t = c* + chol(K*(X_t) +
  (sigma_n*^2 + 1e-6) I) z, z ~ N(0, I) (seed 2000+t): ONE draw from the
  ground-truth GP (SE kernel, c* = 5, l* = 1, sigma_f*^2 = 1, sigma_n*^2 = 0.01,
  used un-warped as in gp_test.py:64-70)."""
  import numpy as np
  tasks = range(T) if tasks is None else tasks
  xs, ys = [], []
  for t in tasks:
    x = np.random.Generator(np.random.PCG64(1000 + t + seed)).random((n, d))
    z = np.random.Generator(np.random.PCG64(2000 + t + seed)).standard_normal(n)
    sq = np.sum(x * x, axis=1)
    r2 = np.maximum(sq[:, None] + sq[None, :] - 2.0 * (x @ x.T), 0.0)  # l* = 1
    k = np.exp(-0.5 * r2)
    k[np.diag_indices(n)] = 1.0 + 0.01 + 1e-6
    ys.append(5.0 + np.linalg.cholesky(k) @ z)
    xs.append(x)
  return np.stack(xs), np.stack(ys)[..., None]


def init_raw(d):
  import numpy as np
  # gp_test.py:102-108: constant 5.1, lengthscale 0, signal 0, noise -4 (raw)
  return np.concatenate([[5.1, 0.0, -4.0], np.zeros(d)])


# ------------------------------------------------------------------ clocks ---
class ClockSampler(threading.Thread):
  """Samples SM clock / throttle reasons of one GPU while the timed region
  runs (pynvml; nvidia-smi fallback)."""

  def __init__(self, index):
    super().__init__(daemon=True)
    self.index, self.samples, self.reasons = index, [], set()
    self.max_mhz = None
    self._stop_evt = threading.Event()
    self._nvml = None
    try:
      import pynvml
      pynvml.nvmlInit()
      self._nvml = pynvml
      self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
      self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h,
                                                       pynvml.NVML_CLOCK_SM)
    except Exception:
      self._nvml = None

  def _sample_nvml(self):
    n = self._nvml
    self.samples.append(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM))
    try:
      r = n.nvmlDeviceGetCurrentClocksEventReasons(self._h)
    except Exception:
      r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
    names = {
        "sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20,
        "hw_thermal_slowdown": 0x40, "hw_power_brake_slowdown": 0x80,
    }
    for k, bit in names.items():
      if r & bit:
        self.reasons.add(k)

  def _sample_smi(self):
    import subprocess
    q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    out = subprocess.run(
        ["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
         "--format=csv,noheader,nounits"], capture_output=True, text=True,
        timeout=5).stdout.strip().split(",")
    self.samples.append(int(out[0]))
    self.max_mhz = int(out[1])
    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown",
                        "sw_thermal_slowdown", "sw_power_cap"), out[2:]):
      if "Active" in v and "Not" not in v:
        self.reasons.add(name)

  def run(self):
    while not self._stop_evt.is_set():
      try:
        if self._nvml is not None:
          self._sample_nvml()
        else:
          self._sample_smi()
      except Exception:
        pass
      self._stop_evt.wait(0.02)

  def stop(self):
    self._stop_evt.set()
    self.join(timeout=2)
    s = sorted(self.samples)
    return {
        "sm_mhz": s[len(s) // 2] if s else None,
        "sm_max_mhz": self.max_mhz,
        "reasons": sorted(self.reasons),
        "samples": len(s),
    }


# ---------------------------------------------------------- reference arm ---
def _host_cores():
  try:
    return len(os.sched_getaffinity(0))
  except AttributeError:
    return os.cpu_count() or 1


def cpu_port_trainer(tasks, dtype_name="f64", batched=True):
  """The CPU restatement of the reference step (torch-CPU op-by-op port with
  autograd + Adam, oracle/hyperbo_oracle_torch.py) on `tasks` of the workload's
  256 tasks, with all host threads (torchrun exports OMP_NUM_THREADS=1 to its
  workers: a launcher default, not a property of the baseline)."""
  import torch
  from oracle import hyperbo_oracle_torch as OT
  ncores = _host_cores()
  if torch.get_num_threads() < ncores:
    torch.set_num_threads(ncores)
  x, y = synthetic_batch(T_TASKS, N_PTS, DIM, tasks=tasks)
  model = {"constant": 5.1, "lengthscale": [0.0] * DIM, "signal_variance": 0.0,
           "noise_variance": -4.0}
  dt = torch.float64 if dtype_name == "f64" else torch.float32
  return OT.AdamTrainer("constant", "squared_exponential", model, x, y, lr=LR,
                        dtype=dt, batched=batched), torch.get_num_threads()


def time_steps(tr, steps, warmup):
  for _ in range(warmup):
    tr.step()
  t0 = time.perf_counter()
  loss = None
  for _ in range(steps):
    loss = tr.step()
  return (time.perf_counter() - t0) / steps, loss


def run_reference(args):
  """The reference's own CPU path for the SAME config, steps and warm-up: all
  256 tasks per step (nothing sampled, nothing scaled).  JAX is not installable
  in this image, so the arm is the op-by-op torch-CPU port of the reference step
  (`kind: "port"`), with the reference's Python loop over tasks
  (objectives.py:181) -- on this host that is also the FASTER variant (the
  task-batched program materialises the (T, n, n, d) difference tensor); the
  batched variant is timed next to it on one step."""
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  tr, cores = cpu_port_trainer(range(T_TASKS), "f64", batched=False)
  dt_s, loss = time_steps(tr, args.steps, args.warmup)
  v = 1.0 / dt_s
  batched = None
  if not args.no_looped:
    trb, _ = cpu_port_trainer(range(T_TASKS), "f64", batched=True)
    db, _ = time_steps(trb, 1, 1)
    batched = {"value": 1.0 / db, "unit": "steps/s", "steps": 1,
               "what": "same step as ONE task-batched torch program instead of "
                       "the reference's loop over tasks"}
  line = {
      "impl": "reference", "metric": METRIC, "value": v, "unit": "steps/s",
      "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
      "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "strong",
      "vs_baseline": None, "dtype": "f64", "data": "synthetic",
      "config": {"workload": "configs[1]: 256 tasks x n=512 x d=8 SE-ARD + "
                             "constant mean, fp64 NLL+grad+Adam step",
                 "tasks": T_TASKS, "n": N_PTS, "d": DIM, "lr": LR},
      "final_loss": loss,
      "cpu_baseline": {
          "value": v, "unit": "steps/s", "cores": cores, "kind": "port",
          "sample": f"all {T_TASKS} tasks per step, {args.steps} timed steps after "
                    f"{args.warmup} warm-up ({dt_s:.3f} s each); torch-CPU op-by-op "
                    "port of the reference step (JAX is not installable in this "
                    "image), looping over the tasks as objectives.py:181 does, all "
                    "host threads",
          "task_batched": batched},
      "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0,
              "d2h_bytes_per_step": 0},
  }
  print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ our arm ---
def run_ours(args):
  import numpy as np
  import torch
  import torch.distributed as dist
  from hyperbo_b200.engine import Engine, PackedDataset
  from hyperbo_b200.gp_utils.gp import AdamTrainer, shard_tasks

  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  torch.cuda.set_device(local)
  if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
  T, n, d = T_TASKS, N_PTS, DIM
  mine = shard_tasks(list(range(T)), rank, world)  # strong scaling
  Tl = len(mine)
  x_np, y_np = synthetic_batch(T, n, d, tasks=mine)
  mask = 0b110 | (((1 << d) - 1) << 3)
  NWIN = max(1, args.windows)

  def barrier(dev):
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize(dev)

  def measure(dtype_name, with_sections):
    """-> dict with the device-resident and the host-buffer (e2e) timings of
    the step for one engine precision."""
    f32 = dtype_name == "f32"
    tdt = torch.float32 if f32 else torch.float64
    eng = Engine.get(local, dtype=tdt)
    dev = eng.device
    x_host = torch.from_numpy(np.ascontiguousarray(x_np.reshape(Tl * n, d))).to(tdt).pin_memory()
    y_host = torch.from_numpy(np.ascontiguousarray(y_np.reshape(Tl * n))).to(tdt).pin_memory()
    ds = PackedDataset(mine, x_host.to(dev), y_host.to(dev),
                       [n * t for t in range(Tl + 1)])

    def make_trainer():
      return AdamTrainer(eng, 0, 1, init_raw(d), mask, d, LR, allreduce=world > 1)

    def timed(trainer, from_host, steps, graph=True):
      """K steps bracketed by barrier+sync, device-timed with CUDA events; max
      over ranks.  One loss read-back per step (the isfinite host check of
      gp.py:135-138), pipelined one step behind the enqueue."""
      e0 = torch.cuda.Event(enable_timing=True)
      e1 = torch.cuda.Event(enable_timing=True)
      barrier(dev)
      e0.record()
      for _ in range(steps):
        prev = trainer.step_pipelined(ds, x_host if from_host else None,
                                      y_host if from_host else None,
                                      use_graph=graph)
        if prev is not None and not math.isfinite(prev):
          raise FloatingPointError("non-finite loss in bench")
      last = trainer.flush()
      if not math.isfinite(last):
        raise FloatingPointError("non-finite loss in bench")
      e1.record()
      barrier(dev)
      ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
      if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
      return float(ms) / steps, last

    def windows(trainer, from_host):
      """NWIN windows of exactly K steps; the median window is the result."""
      w = [timed(trainer, from_host, args.steps) for _ in range(NWIN)]
      ms = sorted(v[0] for v in w)
      return ms[len(ms) // 2], ms, w[-1][1]

    tr = make_trainer()
    l0 = eng.launch_count()
    tr.step(ds)
    launches_per_step = eng.launch_count() - l0
    tr.loss()
    for _ in range(max(args.warmup - 1, 2)):
      tr.step(ds, use_graph=True)
      tr.loss()
    res = {"eng": eng, "launches_per_step": launches_per_step, "esz": 4 if f32 else 8}
    res["ms"], res["ms_windows"], res["loss"] = windows(tr, False)
    if with_sections:
      # the same K steps once more, eagerly, with the library's per-kernel CUDA
      # events (events cannot be queried inside a graph): roofline numerators
      eng.h.profile_enable(True)
      res["ms_eager"], _ = timed(tr, False, args.steps, graph=False)
      res["prof_ms"], res["prof_cnt"] = eng.h.profile_read()
      eng.h.profile_enable(False)
    # end to end through the public trainer with HOST buffers
    tr2 = make_trainer()
    for _ in range(3):
      tr2.step_from_host(ds, x_host, y_host, use_graph=True)
      tr2.loss()
    res["ms_e2e"], res["ms_e2e_windows"], _ = windows(tr2, True)
    res["h2d"] = int((x_host.numel() + y_host.numel()) * res["esz"])
    res["ds"], res["x_host"], res["y_host"] = ds, x_host, y_host
    return res

  main_dt = args.dtype
  sampler = ClockSampler(local)
  sampler.start()
  R = measure(main_dt, True)
  clocks = sampler.stop()
  eng = R["eng"]
  dev = eng.device
  f32 = main_dt == "f32"
  # the other precision, same steps (one line for the driver: configs[1] AND [2])
  other = None
  if not args.single_dtype:
    other_dt = "f64" if f32 else "f32"
    O = measure(other_dt, False)
    other = {"dtype": other_dt, "value": 1e3 / O["ms"], "unit": "steps/s",
             "ms_per_step": O["ms"], "ms_per_step_windows": O["ms_windows"],
             "e2e_value": 1e3 / O["ms_e2e"], "final_loss": O["loss"],
             "workload": "configs[2] (fp32 engine, same shapes)" if other_dt == "f32"
                         else "configs[1] (fp64 engine, same shapes)"}

  # ---- e2e through the reference's entry point: GP(...).train() (gp.py:454-485)
  e2e_train = None
  if world == 1 and not args.no_train_e2e:
    from hyperbo_b200.basics import definitions as defs
    from hyperbo_b200.gp_utils import gp as gpm, kernel, mean, objectives, utils
    from hyperbo_b200 import engine as engmod
    engmod.set_default_dtype(torch.float32 if f32 else torch.float64)
    dataset = {t: defs.SubDataset(x_np[i], y_np[i]) for i, t in enumerate(mine)}
    K = args.steps

    def train_once(steps):
      params = defs.GPParams(
          model={"constant": 5.1, "lengthscale": np.zeros(d), "signal_variance": 0.0,
                 "noise_variance": -4.0},
          config={"method": "adam", "learning_rate": LR, "max_training_step": steps,
                  "batch_size": n + 1, "objective": objectives.nll})
      model = gpm.GP(dataset, mean.constant, kernel.squared_exponential, params,
                     utils.DEFAULT_WARP_FUNC)
      torch.cuda.synchronize(dev)
      t0 = time.perf_counter()
      model.train(key=0)
      torch.cuda.synchronize(dev)
      return time.perf_counter() - t0

    train_once(3)  # warm-up: workspace, plan, graph
    wall = sorted(train_once(K) for _ in range(3))[1]
    e2e_train = {"value": K / wall, "unit": "steps/s", "steps": K, "wall_s": wall,
                 "path": "GP(dataset, mean.constant, kernel.squared_exponential, "
                         "params).train(): host numpy dataset in, packing + upload "
                         "+ K Adam steps + per-step loss read-back + the final "
                         "acceptance evaluation, wall clock (median of 3)"}

  # ---- per-step sub-sampling on the device (data_utils.py:72-100), batch 256
  subsampled = None
  if world == 1 and not args.no_train_e2e:
    from hyperbo_b200.engine import DeviceSampler
    smp = DeviceSampler(eng, R["ds"], 256, 0)
    trs = AdamTrainer(eng, 0, 1, init_raw(d), mask, d, LR)
    for _ in range(3):
      trs.step(smp.dst, use_graph=True)
      trs.loss()
    torch.cuda.synchronize(dev)
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for _ in range(args.steps):
      trs.step_pipelined(smp.dst, use_graph=True)
    trs.flush()
    a1.record()
    torch.cuda.synchronize(dev)
    ms_ss = a0.elapsed_time(a1) / args.steps
    subsampled = {"batch_size": 256, "value": 1e3 / ms_ss, "unit": "steps/s",
                  "ms_per_step": ms_ss,
                  "what": "same 256 x 512 x 8 dataset, every step draws 256 of each "
                          "task's 512 points on the device (hb_subsample in the "
                          "step's CUDA graph), then NLL+grad+Adam on 256 x 256 x 8"}

  # ---- the reference's second objective (8f rank 3): empirical KL on aligned data,
  # value + gradient per call: one factorisation per sub-dataset (hb_nll_grad_mrhs)
  # against the m + 2 weighted-task form.  Extra evidence only: never fails the line.
  kl_objective = None
  if world == 1 and not args.no_train_e2e and not f32:
    try:
      from hyperbo_b200.basics import definitions as defs, params_utils
      from hyperbo_b200.gp_utils import kernel, mean, objectives, utils
      rng = np.random.default_rng(5)
      nk, mk = 512, 20
      xk = rng.uniform(size=(nk, d))
      yk = np.sin(xk.sum(1))[:, None] + 0.3 * rng.standard_normal((nk, mk))
      dsk = {"a": defs.SubDataset(xk, yk, aligned=1)}
      rawk, maskk, _ = params_utils.pack_raw(
          {"constant": 0.1, "signal_variance": 0.0, "noise_variance": -3.0,
           "lengthscale": np.zeros(d)}, d, True, utils.DEFAULT_WARP_FUNC)
      rawk = eng.tensor(rawk)
      kl_objective = {"n": nk, "m": mk, "d": d, "unit": "ms per value+gradient"}
      vals = {}
      for flag, key in ((True, "multi_rhs_ms"), (False, "weighted_tasks_ms")):
        objectives.KL_MULTI_RHS = flag
        prog = objectives.compile_objective(objectives.kl, mean.constant,
                                            kernel.matern52, dsk)
        for _ in range(3):
          sk = prog.sums(rawk, maskk)
        torch.cuda.synchronize(dev)
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        for _ in range(20):
          sk = prog.sums(rawk, maskk)
        k1.record()
        torch.cuda.synchronize(dev)
        kl_objective[key] = k0.elapsed_time(k1) / 20
        vals[key] = float(sk[0])
      objectives.KL_MULTI_RHS = True
      kl_objective["speedup"] = (kl_objective["weighted_tasks_ms"] /
                                 kl_objective["multi_rhs_ms"])
      kl_objective["value_rel_diff"] = abs(
          vals["multi_rhs_ms"] - vals["weighted_tasks_ms"]) / abs(vals["weighted_tasks_ms"])
    except Exception as e:  # pylint: disable=broad-except
      kl_objective = {"error": repr(e)}

  # ---- factorise-only timings: the second half of BASELINE's metric
  def chol_block(tag, T_, n_, d_, kid, reps):
    rng = np.random.default_rng(11)
    xs = torch.as_tensor(rng.random((T_ * n_, d_)), device=dev, dtype=eng.dtype)
    ys = torch.as_tensor(5.0 + rng.standard_normal(T_ * n_), device=dev, dtype=eng.dtype)
    dsc = PackedDataset(list(range(T_)), xs, ys, [n_ * t for t in range(T_ + 1)])
    raw = init_raw(d_)
    msk = 0b110 | (((1 << d_) - 1) << 3)
    for _ in range(2):
      eng.factorize(kid, 1, dsc, raw, msk, want_chol=False, want_alpha=False)
    torch.cuda.synchronize(dev)
    best = []
    for _ in range(reps):
      a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      a.record()
      eng.factorize(kid, 1, dsc, raw, msk, want_chol=False, want_alpha=False)
      b.record()
      torch.cuda.synchronize(dev)
      best.append(a.elapsed_time(b))
    ms = sorted(best)[len(best) // 2]
    fl = T_ * n_**3 / 3.0
    return {"config": tag, "tasks": T_, "n": n_, "d": d_, "ms": ms,
            "flops_n3_over_3": fl, "achieved": fl / ms / 1e9, "unit": "TFLOP/s"}

  if rank != 0:
    if world > 1:
      dist.destroy_process_group()
    return

  # ---- roofline denominators: measured tensor peak of the engine's precision:
  # cuBLAS DGEMM for fp64; cuBLAS TF32 GEMM / 3 for the fp32 engine (3xTF32)
  tdt = eng.dtype
  if f32:
    torch.backends.cuda.matmul.allow_tf32 = True
  a = torch.randn(8192, 8192, device=dev, dtype=tdt)
  b = torch.randn(8192, 8192, device=dev, dtype=tdt)
  for _ in range(2):
    a @ b
  best = 1e9
  for _ in range(5):
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(
        enable_timing=True)
    s0.record()
    a @ b
    s1.record()
    torch.cuda.synchronize(dev)
    best = min(best, s0.elapsed_time(s1))
  peak_tf = 2 * 8192**3 / best / 1e9
  if f32:
    peak_tf /= 3.0
  del a, b
  peak_source = ("in-run cuBLAS TF32 GEMM 8192^3 best of 5, divided by 3 (the fp32 "
                 "engine spends 3 TF32 MMAs per product)" if f32 else
                 "in-run cuBLAS DGEMM 8192^3 best of 5 (fp64 tensor pipe; "
                 "MEASURED_PEAKS.json holds no fp64 figure)")

  cholesky = None
  if world == 1 and not args.no_cholesky:
    cholesky = [chol_block("configs[1]/[2]: 256 x 512 x 8 SE", T, n, d, 0, 5),
                chol_block("configs[4]: 32 x 4096 x 16 Matern-5/2", 32, 4096, 16, 2, 3)]
    for c in cholesky:
      c["peak"] = peak_tf
      c["frac"] = c["achieved"] / peak_tf

  fl = algorithmic_flops(Tl, n, d)
  prof_ms, prof_cnt = R["prof_ms"], R["prof_cnt"]
  ms_step = R["ms"]
  # DRAM bytes of the dominant kernel from the committed ncu capture, valid for
  # the task count it was taken at only (null otherwise)
  traffic = None
  tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
  if os.path.exists(tpath) and not f32:
    try:
      tj = json.load(open(tpath))
      if tj.get("tasks_per_gpu") == Tl:
        traffic = tj.get("k_fused_dram_bytes_per_launch")
    except Exception:
      traffic = None

  # the dominant kernel IS the step: one persistent launch (k_fused) computes
  # kernel tiles, Cholesky, inverse, alpha and the gradient contraction.  (With
  # few tasks per GPU the launch-per-column path runs instead: the dominant
  # group is then its k_step launches.)
  ms_fused = prof_ms[2] / max(prof_cnt[2], 1)
  persistent = R["launches_per_step"] <= 6
  if persistent:
    dom_ms, dom_fl = ms_fused, fl["step"]
    dom_name = ("k_fused (persistent, dependency-driven: kernel-matrix tiles + blocked "
                "Cholesky + triangular inverse + alpha + K~^-1 tiles with the gradient "
                "contraction) -- the dominant kernel, ~98% of the step")
  else:
    dom_ms, dom_fl = prof_ms[0] / max(prof_cnt[0], 1), fl["factor_launches"]
    dom_name = ("k_step x (nblk+1) launches (few tasks per GPU: launch-per-column path): "
                "kernel build + blocked Cholesky + triangular inverse")
  ach = dom_fl / dom_ms / 1e9
  roofline = {
      "bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s",
      "frac": ach / peak_tf, "ms_per_launch": dom_ms, "traffic": traffic,
      "kernel": dom_name,
      "peak_source": peak_source, "flops_per_launch": dom_fl,
      "algorithmic_flops": "SURVEY 8(d): T (n^3 + 4 n^2 + n^2 (3d+8) + n^2 (2d+6))"}
  roofline_step = {
      "bound": "tensor", "achieved": fl["step"] / ms_step / 1e9,
      "peak": peak_tf, "unit": "TFLOP/s",
      "frac": fl["step"] / ms_step / 1e9 / peak_tf,
      "flops_per_step": fl["step"]}

  # ---- CPU baseline on this box's host cores (bounded sample, rank 0, N=1)
  cpu = None
  if not args.no_cpu_baseline and world == 1:
    sample = 64
    trc, cores = cpu_port_trainer(range(sample), main_dt, batched=False)
    dt_s, _ = time_steps(trc, 3, 1)
    cpu = {"value": 1.0 / (dt_s * T / sample), "unit": "steps/s", "cores": cores,
           "kind": "port",
           "sample": f"{sample} of {T} tasks x 3 steps after 1 warm-up "
                     f"({dt_s:.3f} s/step), scaled x{T // sample} (the full-size "
                     f"run is `--impl reference`); torch-CPU port of the reference "
                     f"step, {main_dt}, looping over the tasks (objectives.py:181)"}

  esz = R["esz"]
  line = {
      "metric": METRIC, "value": 1e3 / ms_step, "unit": "steps/s",
      "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
      "ms_per_step": ms_step, "windows": NWIN,
      "ms_per_step_windows": R["ms_windows"],
      "ms_per_step_eager_launches": R["ms_eager"],
      "higher_is_better": True, "scaling": "strong",
      "vs_baseline": None, "dtype": main_dt, "data": "synthetic",
      "config": {
          "workload": ("configs[2]" if f32 else "configs[1]") +
                      ": 256 tasks x n=512 x d=8 SE-ARD + constant mean, " +
                      ("fp32" if f32 else "fp64") +
                      " NLL+grad+Adam step (tasks sharded t%N over N GPUs, one "
                      "peer-memory all-reduce of P+2 scalars fused with Adam per step)",
          "tasks": T, "n": n, "d": d, "tasks_per_gpu": Tl, "lr": LR,
          "inputs": "SURVEY 8(d): one draw per task from the ground-truth GP",
          "timing": f"median of {NWIN} windows of {args.steps} steps each",
          "l2_policy": "no explicit flush: each step streams its packed L, L^-1 "
                       "and W tiles (%.0f MB per GPU) through the 126 MB L2, so no "
                       "step starts with its working set resident"
                       % (3 * Tl * 36 * 4096 * esz / 1e6)},
      "final_loss": R["loss"],
      "clocks": clocks,
      "e2e": {"value": 1e3 / R["ms_e2e"], "unit": "steps/s",
              "ms_per_step": R["ms_e2e"],
              "ms_per_step_windows": R["ms_e2e_windows"],
              "h2d_bytes_per_step": R["h2d"], "d2h_bytes_per_step": esz,
              "path": "gp.AdamTrainer.step_pipelined(ds, x_host, y_host): every "
                      "step uploads its batch from pinned host memory on a copy "
                      "stream into one of two device buffers (overlapping the "
                      "previous step's kernels) and reads its loss back"},
      "e2e_train": e2e_train,
      "gpu_launches": int(R["launches_per_step"] * args.steps * NWIN),
      "gpu_launches_per_step": int(R["launches_per_step"]),
      "roofline": roofline,
      "roofline_step": roofline_step,
      "cholesky": cholesky,
      "section_ms_per_step": (
          {"path": "persistent kernel", "prep": prof_ms[0] / max(prof_cnt[0], 1),
           "k_fused": ms_fused, "task_final": prof_ms[1] / max(prof_cnt[1], 1),
           "reduce": prof_ms[3] / max(prof_cnt[3], 1)}
          if R["launches_per_step"] <= 6 else
          {"path": "launch per block column (few tasks per GPU)",
           "prep+k_step": prof_ms[0] / max(prof_cnt[0], 1),
           "k_alpha": prof_ms[1] / max(prof_cnt[1], 1),
           "k_lauum_grad": ms_fused,
           "reduce": 2 * prof_ms[3] / max(prof_cnt[3], 1)}),
      "subsampled_training": subsampled,
      "kl_objective": kl_objective,
      "other_precision": other,
      "cpu_baseline": cpu,
  }
  print(json.dumps(line), flush=True)
  if world > 1:
    dist.destroy_process_group()


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=50)
  ap.add_argument("--warmup", type=int, default=5)
  ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--windows", type=int, default=5,
                  help="timed windows of --steps steps each (median reported)")
  ap.add_argument("--single-dtype", action="store_true",
                  help="skip the measurement of the other engine precision")
  ap.add_argument("--no-train-e2e", action="store_true")
  ap.add_argument("--no-cholesky", action="store_true")
  ap.add_argument("--no-looped", action="store_true",
                  help="reference arm: skip the task-looped variant")
  ap.add_argument("--tasks", type=int, default=None,
                  help="experiments only: total task count instead of 256")
  ap.add_argument("--dtype", default="f64", choices=["f64", "f32"],
                  help="engine precision: f64 = BASELINE configs[1] (default), "
                       "f32 = configs[2]")
  args = ap.parse_args()
  args.warmup = max(args.warmup, 3)
  if args.tasks:
    global T_TASKS
    T_TASKS = args.tasks
  if args.impl == "reference":
    run_reference(args)
  else:
    run_ours(args)


if __name__ == "__main__":
  main()
