"""Generate the committed fixtures under tests/golden/ -- TEST INFRASTRUCTURE.

The reference (JAX) cannot be imported in this image and ships no golden
vectors, so these fixtures are produced by oracle/hyperbo_oracle.py (the numpy
fp64 restatement) and cross-checked, at generation time, against the
independent torch-autograd restatement; generation aborts if the two disagree
by more than 1e-9 relative.  "Parity unpinned" still applies (see the oracle
header): the fixtures pin the ORACLE, and let the GPU tests compare against
fixed files.

  python -m oracle.make_golden        (from the repo root)
"""
import os
import sys

import numpy as np

from oracle import hyperbo_oracle as O
from oracle import hyperbo_oracle_torch as OT

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "tests", "golden")

CASES = {
    # name: (cov, mean, ns, d, ls_seed)
    "c1_se_1x64x1": ("squared_exponential", "constant", [64], 1, 0),
    "ragged_m52_d4": ("matern52", "constant", [37, 64, 130, 1], 4, 1),
    "m32_zero_mean_d2": ("matern32", "zero", [100, 65], 2, 2),
    "se_d8_n200": ("squared_exponential", "constant", [200, 128], 8, 3),
}


def raw_vec(model, d):
  ls = np.broadcast_to(np.asarray(model["lengthscale"], dtype=np.float64), (d,))
  return np.concatenate([[model.get("constant", 0.0), model["signal_variance"],
                          model["noise_variance"]], ls])


def grad_vec(g, d):
  ls = np.broadcast_to(np.asarray(g["lengthscale"], dtype=np.float64), (d,))
  return np.concatenate([[g.get("constant", 0.0), g["signal_variance"],
                          g["noise_variance"]], ls])


def build(name):
  cov, mean, ns, d, seed = CASES[name]
  rng = np.random.default_rng(100 + seed)
  ds = {t: O.make_task(50 * seed + t, n, d, cov) for t, n in enumerate(ns)}
  model = O.init_raw_params(d)
  model["lengthscale"] = rng.normal(0.0, 0.4, d)
  model["signal_variance"] = float(rng.normal(0.0, 0.3))
  if mean == "zero":
    del model["constant"]
  wf = O.DEFAULT_WARP_FUNC
  val, grad = O.nll_value_and_grad(mean, cov, model, ds, wf)
  val_t, grad_t = OT.value_and_grad(mean, cov, model, ds)
  assert abs(val - val_t) <= 1e-9 * abs(val), (name, val, val_t)
  for k in grad:
    if k in grad_t:
      a, b = np.asarray(grad[k], dtype=np.float64), np.asarray(grad_t[k])
      assert np.max(np.abs(a - b)) <= 1e-9 * (np.max(np.abs(b)) + 1e-12), (name, k)
  nll_task = np.array([O.nll_sub_dataset(mean, cov, model, *ds[t], warp_func=wf)
                       for t in range(len(ns))])
  chol0, alpha0, _ = O.solve_gp_linear_system(mean, cov, model, *ds[0],
                                              warp_func=wf)
  xq = rng.random((40, d))
  mu, var = O.gp_predict(mean, cov, model, ds, xq, 0, wf)
  ei = O.acquisition("ei", mean, cov, model, ds, 0, xq, wf)
  pi = O.acquisition("pi", mean, cov, model, ds, 0, xq, wf)
  ucb = O.acquisition("ucb", mean, cov, model, ds, 0, xq, wf)
  out = {
      "cov": cov, "mean": mean, "d": d, "ns": np.array(ns),
      "raw": raw_vec(model, d), "mean_nll": val, "grad": grad_vec(grad, d),
      "nll_task": nll_task, "chol0": chol0, "alpha0": alpha0.ravel(),
      "xq": xq, "mu": mu.ravel(), "var": var.ravel(), "ei": ei.ravel(),
      "pi": pi.ravel(), "ucb": ucb.ravel(),
  }
  for t in range(len(ns)):
    out[f"x{t}"], out[f"y{t}"] = ds[t][0], ds[t][1]
  return out


def make_aligned_dataset(seed, d, spec):
  """Synthetic dataset with ALIGNED sub-datasets (y: n x m, m tasks evaluated on
  the same inputs -- what the reference's EKL objective consumes) plus
  ordinary ones.  spec: list of (n, m, aligned_tag_or_None)."""
  rng = np.random.default_rng(seed)
  ds = {}
  for t, (n, m, tag) in enumerate(spec):
    x = rng.random((n, d))
    base = np.sin(3.0 * x[:, :1]) + 0.5 * x[:, -1:]
    y = 0.3 + base + 0.4 * rng.standard_normal((n, m)) * (1 + x[:, :1])
    ds[t] = (x, y) if tag is None else (x, y, tag)
  return ds


KL_CASES = {
    # name: (cov, mean, d, spec)
    "kl_m52_const_d3": ("matern52", "constant", 3,
                        [(70, 6, 1), (33, 4, "tag"), (40, 1, None), (0, 3, 2)]),
    "kl_se_zero_d2": ("squared_exponential", "zero", 2,
                      [(130, 9, True), (20, 1, None)]),
}


def build_kl(name):
  import functools
  cov, mean, d, spec = KL_CASES[name]
  ds = make_aligned_dataset(7 + len(name), d, spec)
  rng = np.random.default_rng(5)
  model = O.init_raw_params(d)
  model["lengthscale"] = rng.normal(0.0, 0.4, d)
  model["constant"] = 0.2
  if mean == "zero":
    del model["constant"]
  wf = O.DEFAULT_WARP_FUNC
  val, grad = O.kl_value_and_grad(mean, cov, model, ds, wf)
  val_t, grad_t = OT.divergence_value_and_grad(mean, cov, model, ds)
  assert abs(val - val_t) <= 1e-9 * abs(val), (name, val, val_t)
  for k in grad:
    a, b = np.asarray(grad[k], dtype=np.float64), np.asarray(grad_t[k])
    assert np.max(np.abs(a - b)) <= 1e-9 * (np.max(np.abs(b)) + 1e-12), (name, k)
  kl = O.kl_multivariate_normal
  out = {
      "cov": cov, "mean": mean, "d": d, "raw": raw_vec(model, d),
      "spec_n": np.array([s[0] for s in spec]),
      "spec_m": np.array([s[1] for s in spec]),
      "spec_aligned": np.array([s[2] is not None for s in spec]),
      "kl": val, "kl_grad": grad_vec(grad, d),
      "kl_eps": O.multivariate_normal_divergence(
          mean, cov, model, ds, wf, functools.partial(kl, eps=1e-6)),
      "kl_full": O.multivariate_normal_divergence(
          mean, cov, model, ds, wf,
          functools.partial(kl, eps=1e-6, partial=False)),
      "euc": O.multivariate_normal_divergence(
          mean, cov, model, ds, wf, O.euclidean_multivariate_normal),
  }
  for t in ds:
    out[f"x{t}"], out[f"y{t}"] = ds[t][0], ds[t][1]
  return out


if __name__ == "__main__":
  os.makedirs(OUT, exist_ok=True)
  for name in CASES:
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **build(name))
    print("wrote", name)
  for name in KL_CASES:
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **build_kl(name))
    print("wrote", name)
