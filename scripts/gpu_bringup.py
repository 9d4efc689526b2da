"""GPU bring-up: engine vs oracle on small cases, then first timings.
Run under gpurun:  timeout 600 python scripts/gpu_bringup.py [--big]
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyperbo_b200.engine import Engine, KERNEL_IDS, MEAN_IDS  # noqa: E402
from oracle import hyperbo_oracle as O  # noqa: E402


def raw_vec(model, d):
  ls = np.broadcast_to(np.asarray(model["lengthscale"], dtype=np.float64), (d,))
  return np.concatenate([[model["constant"], model["signal_variance"],
                          model["noise_variance"]], ls])


def grad_vec(g, d):
  ls = np.broadcast_to(np.asarray(g["lengthscale"], dtype=np.float64), (d,))
  return np.concatenate([[g.get("constant", 0.0), g["signal_variance"],
                          g["noise_variance"]], ls])


def rel(a, b):
  a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
  return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300))


def check(eng, cov, ns, d, seed=0):
  ds_np = {}
  for t, n in enumerate(ns):
    ds_np[t] = O.make_task(10 * seed + t, n, d, cov)
  model = O.init_raw_params(d)
  rng = np.random.default_rng(seed)
  model["lengthscale"] = rng.normal(0, 0.3, d)
  mask = 0b110 | (((1 << d) - 1) << 3)
  raw = raw_vec(model, d)
  ds = eng.pack([(k, v[0], v[1]) for k, v in ds_np.items()])
  kid, mid = KERNEL_IDS[cov], MEAN_IDS["constant"]
  # Gram
  x0 = ds_np[0][0]
  K = eng.kernel_matrix(kid, x0, None, raw, mask, add_noise=True).cpu().numpy()
  _, Kref = O.compute_delta_y_and_cov("constant", cov, model, x0, ds_np[0][1],
                                      O.DEFAULT_WARP_FUNC)
  e_k = rel(K, Kref)
  # factorize
  chols, alpha, nll, info = eng.factorize(kid, mid, ds, raw, mask)
  torch.cuda.synchronize()
  e_c = e_a = e_n = 0.0
  for t, n in enumerate(ns):
    c_ref, a_ref, _ = O.solve_gp_linear_system("constant", cov, model, *ds_np[t],
                                               warp_func=O.DEFAULT_WARP_FUNC)
    e_c = max(e_c, rel(chols[t].cpu().numpy(), c_ref))
    e_a = max(e_a, rel(alpha[ds.offs[t]:ds.offs[t + 1]].cpu().numpy(),
                       a_ref.ravel()))
    n_ref = O.nll_sub_dataset("constant", cov, model, *ds_np[t],
                              warp_func=O.DEFAULT_WARP_FUNC)
    e_n = max(e_n, abs(float(nll[t]) - n_ref) / abs(n_ref))
  # nll + grad
  sums = eng.nll_grad(kid, mid, ds, raw, mask).cpu().numpy()
  v_ref, g_ref = O.nll_value_and_grad("constant", cov, model, ds_np,
                                      O.DEFAULT_WARP_FUNC)
  T = len(ns)
  e_v = abs(sums[0] / T - v_ref) / abs(v_ref)
  e_g = rel(sums[1:-1] / T, grad_vec(g_ref, d))
  ok = (max(e_k, e_c, e_a, e_n, e_v) < 1e-9 and e_g < 1e-7) or eng.dtype == torch.float32
  print(f"{'OK ' if ok else 'BAD'} {cov:20s} ns={ns} d={d} K={e_k:.1e} "
        f"chol={e_c:.1e} alpha={e_a:.1e} nll={e_n:.1e} val={e_v:.1e} "
        f"grad={e_g:.1e} info={info.tolist()} cnt={sums[-1]}", flush=True)
  return ok


def check_predict(eng, cov, n, d, nq):
  x, y = O.make_task(77, n, d, cov)
  xq = np.random.default_rng(5).random((nq, d))
  model = O.init_raw_params(d)
  mask = 0b110 | (((1 << d) - 1) << 3)
  raw = raw_vec(model, d)
  kid, mid = KERNEL_IDS[cov], MEAN_IDS["constant"]
  cache, chol, kinvy, nll, info = eng.build_predictor(kid, mid, x, y, raw, mask)
  ds = {0: (x, y), 1: (x[:3], y[:3])}
  mu, var, acq = eng.predict(kid, mid, eng.tensor(x), cache, raw, mask, xq,
                             noise_flag=1.0, var_scale=2.0, acq_id=1,
                             acq_param=float(np.max(y)))
  torch.cuda.synchronize()
  mu_r, var_r = O.gp_predict("constant", cov, model, ds, xq, 0,
                             O.DEFAULT_WARP_FUNC)
  ei_r = O.acquisition("ei", "constant", cov, model, ds, 0, xq,
                       O.DEFAULT_WARP_FUNC)
  e = (rel(mu.cpu().numpy(), mu_r), rel(var.cpu().numpy(), var_r),
       rel(acq.cpu().numpy(), ei_r))
  ok = max(e) < 1e-6
  print(f"{'OK ' if ok else 'BAD'} predict {cov} n={n} nq={nq} mu={e[0]:.1e} "
        f"var={e[1]:.1e} ei={e[2]:.1e}", flush=True)
  return ok


def time_c2(eng, T=256, n=512, d=8, iters=10):
  rng = np.random.default_rng(0)
  x = torch.as_tensor(rng.random((T * n, d)), device="cuda", dtype=eng.dtype)
  y = torch.as_tensor(5 + rng.standard_normal(T * n), device="cuda", dtype=eng.dtype)
  from hyperbo_b200.engine import PackedDataset
  ds = PackedDataset(list(range(T)), x, y, [n * t for t in range(T + 1)])
  model = O.init_raw_params(d)
  raw = eng.tensor(raw_vec(model, d))
  mask = 0b110 | (((1 << d) - 1) << 3)
  sums = torch.empty(3 + d + 2, device="cuda", dtype=eng.dtype)
  for _ in range(3):
    eng.nll_grad(0, 1, ds, raw, mask, sums_out=sums)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(
      enable_timing=True)
  e0.record()
  for _ in range(iters):
    eng.nll_grad(0, 1, ds, raw, mask, sums_out=sums)
  e1.record()
  torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / iters
  flop = T * (n**3 + 4 * n * n + n * n * (3 * d + 8) + n * n * (2 * d + 6))
  print(f"C2-shape nll+grad T={T} n={n} d={d}: {ms:.3f} ms/step  "
        f"{flop / ms / 1e9:.2f} TFLOP/s(algorithmic)  loss={float(sums[0]) / T:.6f}",
        flush=True)
  # factorize only (potrf + z + nll, no inverse)
  for _ in range(2):
    eng.factorize(0, 1, ds, raw, mask, want_chol=False, want_alpha=False)
  torch.cuda.synchronize()
  e0.record()
  for _ in range(iters):
    eng.factorize(0, 1, ds, raw, mask, want_chol=False, want_alpha=False)
  e1.record()
  torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / iters
  print(f"C2-shape potrf-only: {ms:.3f} ms  {T * n**3 / 3 / ms / 1e9:.2f} "
        f"TFLOP/s (n^3/3)", flush=True)


def dgemm_peak():
  a = torch.randn(8192, 8192, device="cuda", dtype=torch.float64)
  b = torch.randn(8192, 8192, device="cuda", dtype=torch.float64)
  for _ in range(2):
    a @ b
  torch.cuda.synchronize()
  best = 1e9
  for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(
        enable_timing=True)
    e0.record()
    a @ b
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
  print(f"cuBLAS DGEMM 8192^3: {2 * 8192**3 / best / 1e9:.2f} TFLOP/s", flush=True)


if __name__ == "__main__":
  print(torch.cuda.get_device_name(0), flush=True)
  F32 = "--f32" in sys.argv
  eng = Engine.get(dtype=torch.float32 if F32 else torch.float64)
  if "--time-only" in sys.argv:
    time_c2(eng, iters=2)
    sys.exit(0)
  ok = True
  ok &= check(eng, "squared_exponential", [40], 3)
  ok &= check(eng, "squared_exponential", [64, 100, 130], 3)
  ok &= check(eng, "matern32", [200, 1, 65], 2)
  ok &= check(eng, "matern52", [512, 300], 8)
  ok &= check_predict(eng, "squared_exponential", 150, 3, 200)
  ok &= check_predict(eng, "matern52", 500, 4, 1000)
  print("ALL OK" if ok else "SOME BAD", flush=True)
  dgemm_peak()
  time_c2(eng)
