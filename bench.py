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


def synthetic_batch(T, n, d, seed=0):
  """Synthetic (T x n x d) batch: X ~ U[0,1], y = c* + smooth signal + noise
  (cheap surrogate of the GP draw of SURVEY 8d, same shapes / conditioning)."""
  import numpy as np
  rng = np.random.Generator(np.random.PCG64(seed))
  x = rng.random((T, n, d))
  y = 5.0 + np.sum(np.sin(2 * np.pi * x[..., :2]), axis=-1, keepdims=True) \
      + 0.1 * rng.standard_normal((T, n, 1))
  return x, y


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
def cpu_port_steps_per_sec(sample_tasks, steps, warmup, dtype_name="f64"):
  """Times the CPU restatement of the reference step (torch-CPU op-by-op port
  with autograd + Adam, all host threads) on a bounded sample of the workload
  and scales to the full 256-task step."""
  import torch
  from oracle import hyperbo_oracle_torch as OT
  # all the host cores this process may use (torchrun exports OMP_NUM_THREADS=1
  # to its workers: a launcher default, not a property of the baseline)
  try:
    ncores = len(os.sched_getaffinity(0))
  except AttributeError:
    ncores = os.cpu_count() or 1
  if torch.get_num_threads() < ncores:
    torch.set_num_threads(ncores)
  x, y = synthetic_batch(sample_tasks, N_PTS, DIM)
  model = {"constant": 5.1, "lengthscale": [0.0] * DIM, "signal_variance": 0.0,
           "noise_variance": -4.0}
  dt = torch.float64 if dtype_name == "f64" else torch.float32
  tr = OT.AdamTrainer("constant", "squared_exponential", model, x, y, lr=LR,
                      dtype=dt, batched=True)
  for _ in range(warmup):
    tr.step()
  t0 = time.perf_counter()
  for _ in range(steps):
    tr.step()
  dt_s = (time.perf_counter() - t0) / steps
  full_step_s = dt_s * (T_TASKS / sample_tasks)
  return 1.0 / full_step_s, dt_s, torch.get_num_threads()


def run_reference(args):
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  sample = 32
  steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
  v, dt_s, cores = cpu_port_steps_per_sec(sample, steps, warmup)
  line = {
      "impl": "reference", "metric": METRIC, "value": v, "unit": "steps/s",
      "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
      "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "strong",
      "vs_baseline": None, "dtype": "f64", "data": "synthetic",
      "config": {"workload": "configs[1]: 256 tasks x n=512 x d=8 SE-ARD + "
                             "constant mean, fp64 NLL+grad+Adam step",
                 "tasks": T_TASKS, "n": N_PTS, "d": DIM},
      "cpu_baseline": {
          "value": v, "unit": "steps/s", "cores": cores, "kind": "port",
          "sample": f"{sample} of {T_TASKS} tasks per timed step "
                    f"({dt_s:.3f} s each), scaled x{T_TASKS // sample}; "
                    "torch-CPU op-by-op port of the reference step (JAX is "
                    "not installable in this image), task-batched, all host "
                    "threads"},
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
  f32 = args.dtype == "f32"
  tdt = torch.float32 if f32 else torch.float64
  esz = 4 if f32 else 8
  eng = Engine.get(local, dtype=tdt)
  dev = eng.device

  T, n, d = T_TASKS, N_PTS, DIM
  x_all, y_all = synthetic_batch(T, n, d)
  mine = shard_tasks(list(range(T)), rank, world)  # strong scaling
  Tl = len(mine)
  x_host = torch.from_numpy(np.ascontiguousarray(x_all[mine].reshape(Tl * n, d))).to(tdt).pin_memory()
  y_host = torch.from_numpy(np.ascontiguousarray(y_all[mine].reshape(Tl * n))).to(tdt).pin_memory()
  ds = PackedDataset(mine, x_host.to(dev), y_host.to(dev),
                     [n * t for t in range(Tl + 1)])
  mask = 0b110 | (((1 << d) - 1) << 3)

  def make_trainer():
    return AdamTrainer(eng, 0, 1, init_raw(d), mask, d, LR, allreduce=world > 1)

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize(dev)

  def timed(trainer, from_host, steps, graph=True):
    """K steps bracketed by barrier+sync, device-timed with CUDA events.  The
    launch sequence of a step is replayed from a CUDA graph (single GPU)."""
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
      # one read-back of the loss per step (the isfinite host check of
      # gp.py:135-138), pipelined one step behind the enqueue
      prev = trainer.step_pipelined(ds, x_host if from_host else None,
                                    y_host if from_host else None,
                                    use_graph=graph)
      if prev is not None and not math.isfinite(prev):
        raise FloatingPointError("non-finite loss in bench")
    last = trainer.flush()
    if not math.isfinite(last):
      raise FloatingPointError("non-finite loss in bench")
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
      dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms) / steps, last

  # ---- device-resident run: `value`
  tr = make_trainer()
  l0 = eng.launch_count()
  tr.step(ds)
  launches_per_step = eng.launch_count() - l0
  tr.loss()
  for _ in range(max(args.warmup - 1, 2)):
    tr.step(ds, use_graph=True)
    tr.loss()
  sampler = ClockSampler(local)
  sampler.start()
  ms_step, loss = timed(tr, False, args.steps)
  # same K steps once more, eagerly, with the library's per-kernel CUDA events
  # (events cannot be queried inside a graph): roofline numerators
  eng.h.profile_enable(True)
  ms_step_eager, _ = timed(tr, False, args.steps, graph=False)
  prof_ms, prof_cnt = eng.h.profile_read()
  eng.h.profile_enable(False)
  clocks = sampler.stop()

  # ---- end-to-end run through the public trainer with HOST buffers
  tr2 = make_trainer()
  for _ in range(3):
    tr2.step_from_host(ds, x_host, y_host, use_graph=True)
    tr2.loss()
  ms_e2e, _ = timed(tr2, True, args.steps)

  if rank != 0:
    if world > 1:
      dist.destroy_process_group()
    return

  # ---- roofline denominators: measured tensor peak of the engine's precision:
  # cuBLAS DGEMM for fp64; cuBLAS TF32 GEMM / 3 for the fp32 engine (3xTF32)
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
  dgemm_tf = 2 * 8192**3 / best / 1e9
  if f32:
    dgemm_tf /= 3.0
  del a, b

  fl = algorithmic_flops(Tl, n, d)
  traffic = None
  tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
  if os.path.exists(tpath) and not f32:
    try:
      traffic = json.load(open(tpath)).get("k_lauum_grad_dram_bytes_per_launch")
    except Exception:
      traffic = None

  def roof(flops, ms_total, count):
    if not count:
      return None
    ms = ms_total / count
    ach = flops / ms / 1e9
    return {"bound": "tensor", "achieved": ach, "peak": dgemm_tf,
            "unit": "TFLOP/s", "frac": ach / dgemm_tf, "ms_per_launch": ms}

  roofline = roof(fl["lauum_grad"], prof_ms[2], prof_cnt[2]) or {}
  roofline.update({
      "kernel": "k_lauum_grad (largest single launch: K~^-1 = M'M tiles on the "
                "fp64 tensor pipe + gradient contraction)",
      "traffic": traffic,
      "peak_source": ("in-run cuBLAS TF32 GEMM 8192^3 best of 5, divided by 3 "
                      "(the fp32 engine spends 3 TF32 MMAs per product)" if f32
                      else "in-run cuBLAS DGEMM 8192^3 best of 5 (fp64 tensor "
                      "pipe; MEASURED_PEAKS.json holds no fp64 figure)"),
      "flops_per_launch": fl["lauum_grad"],
  })
  roofline_factor = roof(fl["factor_launches"], prof_ms[0], prof_cnt[0]) or {}
  roofline_factor.update({
      "kernel": "k_step x (nblk+1) launches per step: kernel build + blocked "
                "Cholesky + triangular inverse",
      "flops_per_launch_group": fl["factor_launches"]})
  roofline_step = {
      "bound": "tensor", "achieved": fl["step"] / ms_step / 1e9,
      "peak": dgemm_tf, "unit": "TFLOP/s",
      "frac": fl["step"] / ms_step / 1e9 / dgemm_tf,
      "flops_per_step": fl["step"]}

  # ---- CPU baseline on this box's host cores (bounded sample)
  cpu = None
  if not args.no_cpu_baseline:
    sample = 32
    v, dt_s, cores = cpu_port_steps_per_sec(sample, 3, 1, args.dtype)
    cpu = {"value": v, "unit": "steps/s", "cores": cores, "kind": "port",
           "sample": f"{sample} of {T} tasks x 3 steps ({dt_s:.3f} s/step), "
                     f"scaled x{T // sample}; torch-CPU port of the reference "
                     f"step, {args.dtype}, task-batched"}

  line = {
      "metric": METRIC, "value": 1e3 / ms_step, "unit": "steps/s",
      "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
      "ms_per_step": ms_step, "ms_per_step_eager_launches": ms_step_eager,
      "higher_is_better": True, "scaling": "strong",
      "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
      "config": {
          "workload": ("configs[2]" if f32 else "configs[1]") +
                      ": 256 tasks x n=512 x d=8 SE-ARD + constant mean, " +
                      ("fp32" if f32 else "fp64") +
                      " NLL+grad+Adam step (tasks sharded t%N over "
                      "N GPUs, one all-reduce of P+2 scalars per step)",
          "tasks": T, "n": n, "d": d, "tasks_per_gpu": Tl, "lr": LR,
          "l2_policy": "no explicit flush: each step streams its packed L and "
                       "L^-1 tiles (%.0f MB per GPU) through the 126 MB L2, so "
                       "no step starts with its working set resident"
                       % (3 * Tl * 36 * 4096 * esz / 1e6)},
      "final_loss": loss,
      "clocks": clocks,
      "e2e": {"value": 1e3 / ms_e2e, "unit": "steps/s", "ms_per_step": ms_e2e,
              "h2d_bytes_per_step": int((x_host.numel() + y_host.numel()) * esz),
              "d2h_bytes_per_step": esz,
              "path": "gp.AdamTrainer.step_pipelined(ds, x_host, y_host): every "
                      "step uploads its batch from pinned host memory on a copy "
                      "stream into one of two device buffers (overlapping the "
                      "previous step's kernels) and reads its loss back"},
      "gpu_launches": int(launches_per_step * args.steps),
      "gpu_launches_per_step": int(launches_per_step),
      "roofline": roofline,
      "roofline_factor_launches": roofline_factor,
      "roofline_step": roofline_step,
      "section_ms_per_step": {
          "factor": prof_ms[0] / max(prof_cnt[0], 1),
          "alpha": prof_ms[1] / max(prof_cnt[1], 1),
          "lauum_grad": prof_ms[2] / max(prof_cnt[2], 1),
          "reduce": 2 * prof_ms[3] / max(prof_cnt[3], 1)},
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
