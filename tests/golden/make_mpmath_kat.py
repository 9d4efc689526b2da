"""Known-answer vectors in 60-digit arithmetic -- TEST INFRASTRUCTURE.

Generates tests/golden/kat_mpmath.json.  This script does NOT import oracle/
(nor the product): it restates the reference's formulas directly in mpmath,
so the fixtures pin BOTH the numpy oracle and the CUDA path against an
independent, higher-precision evaluation of the same mathematics:

  kernels       /root/reference/hyperbo/gp_utils/kernel.py:63-123
  warp          /root/reference/hyperbo/gp_utils/utils.py:28-29,73-81
                (softplus(x) + 1e-10 on lengthscale / signal / noise)
  K~ and y-m    /root/reference/hyperbo/basics/linalg.py:36-69 (jitter 1e-6)
  NLL           /root/reference/hyperbo/gp_utils/objectives.py:144-156,178-195
  gradient      what jax.value_and_grad yields at gp_utils/gp.py:134 -- here by
                high-order numerical differentiation (mpmath.diff at 60 digits,
                exact to >= 25 digits), i.e. independent of the closed form the
                oracle and the CUDA kernels use
  predict       /root/reference/hyperbo/gp_utils/gp.py:242-305,584-619
                (with_noise adds sigma_n^2 WITHOUT the jitter; N/(N-1) inflation)
  acquisition   /root/reference/hyperbo/bo_utils/acfun.py:96-165

    python tests/golden/make_mpmath_kat.py     (~1 min)

Inputs come from numpy PCG64 streams, rounded to 1/1024 so that their decimal
/ binary64 / mpmath representations agree exactly.
"""
import json
import os

import mpmath as mp
import numpy as np

mp.mp.dps = 60
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat_mpmath.json")
JITTER = mp.mpf("1e-6")
EPS_WARP = mp.mpf("1e-10")


def softplus(x):
  return mp.log1p(mp.exp(x))


def warp(v, warped):
  return softplus(v) + EPS_WARP if warped else v


def kern(cov, r2, sv):
  if cov == "squared_exponential":
    return sv * mp.exp(-r2 / 2)
  if cov == "matern32":
    r = mp.sqrt(3) * mp.sqrt(r2)
    return sv * (1 + r) * mp.exp(-r)
  r = mp.sqrt(5) * mp.sqrt(r2)
  return sv * (1 + r + r * r / 3) * mp.exp(-r)


def gram(cov, x1, x2, ls, sv):
  n1, n2, d = len(x1), len(x2), len(ls)
  k = mp.zeros(n1, n2)
  for i in range(n1):
    for j in range(n2):
      r2 = mp.fsum(((x1[i][c] - x2[j][c]) / ls[c]) ** 2 for c in range(d))
      k[i, j] = kern(cov, r2, sv)
  return k


def unpack(raw, d, mean, warped):
  const = raw[0] if mean == "constant" else mp.mpf(0)
  sv = warp(raw[1], warped)
  nv = warp(raw[2], warped)
  ls = [warp(raw[3 + c], warped) for c in range(d)]
  return const, sv, nv, ls


def nll_task(cov, mean, warped, raw, x, y):
  n, d = len(x), len(x[0])
  const, sv, nv, ls = unpack(raw, d, mean, warped)
  k = gram(cov, x, x, ls, sv)
  for i in range(n):
    k[i, i] += nv + JITTER
  r = mp.matrix([yi - const for yi in y])
  chol = mp.cholesky(k)
  z = mp.lu_solve(chol, r)                      # L z = r (L lower: exact solve)
  quad = mp.fsum(zi * zi for zi in z)
  logdet = mp.fsum(mp.log(chol[i, i]) for i in range(n))
  return quad / 2 + logdet + mp.mpf(n) / 2 * mp.log(2 * mp.pi)


def mean_nll(cov, mean, warped, raw, tasks):
  return mp.fsum(nll_task(cov, mean, warped, raw, x, y) for x, y in tasks) / len(tasks)


def predict(cov, mean, warped, raw, x, y, xq, n_tasks):
  n, d = len(x), len(x[0])
  const, sv, nv, ls = unpack(raw, d, mean, warped)
  k = gram(cov, x, x, ls, sv)
  for i in range(n):
    k[i, i] += nv + JITTER
  r = mp.matrix([yi - const for yi in y])
  alpha = mp.lu_solve(k, r)
  ks = gram(cov, x, xq, ls, sv)                 # (n, nq)
  kinv_ks = mp.inverse(k) * ks
  # full posterior covariance as GP.predict(full_cov=True, with_noise=True) returns
  # it (gp.py:295-300, 607-619): noise on the diagonal, then N/(N-1)
  kss = gram(cov, xq, xq, ls, sv)
  cov_full = kss - ks.T * kinv_ks
  for q in range(len(xq)):
    cov_full[q, q] += nv
  if n_tasks > 1:
    cov_full = cov_full * (mp.mpf(n_tasks) / (n_tasks - 1))
  predict.cov_full = [[cov_full[a, b] for b in range(len(xq))] for a in range(len(xq))]
  mu, var = [], []
  for q in range(len(xq)):
    mu.append(mp.fsum(ks[i, q] * alpha[i] for i in range(n)) + const)
    v = sv - mp.fsum(ks[i, q] * kinv_ks[i, q] for i in range(n))
    v += nv                                      # with_noise: no jitter
    if n_tasks > 1:
      v *= mp.mpf(n_tasks) / (n_tasks - 1)       # unbiased
    var.append(v)
  return mu, var, alpha


def acq(mu, var, target):
  ei, pi, ucb = [], [], []
  for m, v in zip(mu, var):
    s = mp.sqrt(v)
    g = (target - m) / s
    ei.append(s * (mp.npdf(g) - g * (1 - mp.ncdf(g))))
    g2 = (target + mp.mpf("0.1") - m) / s
    pi.append(-g2)
    ucb.append(m + 3 * s)
  return ei, pi, ucb


def f(v):
  return float(v)


def fl(vs):
  return [float(v) for v in vs]


def q1024(a):
  return np.round(np.asarray(a) * 1024.0) / 1024.0


def build_case(cid, cov, mean, warped, ns, d, rng, dup=False):
  tasks_np = []
  for n in ns:
    x = q1024(rng.random((n, d)))
    if dup and n > 4:  # duplicate inputs: r = 0 off the diagonal (_safe_sqrt)
      x[n - 1] = x[1]
    yv = q1024(5.0 + np.sin(3.0 * x.sum(axis=1)) + 0.3 * rng.standard_normal(n))
    tasks_np.append((x, yv))
  if warped:
    raw_np = q1024(np.concatenate([[5.1], rng.normal(0, 0.5, 1), [-2.0 + rng.normal(0, 0.3)],
                                   rng.normal(0, 0.5, d)]))
  else:
    raw_np = q1024(np.concatenate([[4.9], [0.8 + 0.5 * rng.random()], [0.05 + 0.1 * rng.random()],
                                   0.4 + rng.random(d)]))
  raw = [mp.mpf(float(v)) for v in raw_np]
  tasks = [([[mp.mpf(float(v)) for v in row] for row in x], [mp.mpf(float(v)) for v in yv])
           for x, yv in tasks_np]
  per_task = [nll_task(cov, mean, warped, raw, x, y) for x, y in tasks]
  val = mp.fsum(per_task) / len(tasks)
  grad = []
  for p in range(3 + d):
    if p == 0 and mean == "zero":
      grad.append(mp.mpf(0))
      continue
    def fp(v, p=p):
      r2 = list(raw)
      r2[p] = v
      return mean_nll(cov, mean, warped, r2, tasks)
    grad.append(mp.diff(fp, raw[p], h=mp.mpf("1e-12")))
  xq_np = q1024(rng.random((6, d)))
  xq = [[mp.mpf(float(v)) for v in row] for row in xq_np]
  x0, y0 = tasks[0]
  mu, var, alpha = predict(cov, mean, warped, raw, x0, y0, xq, len(ns))
  target = max(y0)
  ei, pi, ucb = acq(mu, var, target)
  return {
      "id": cid, "cov": cov, "mean": mean, "warped": bool(warped), "d": d,
      "ns": list(ns), "raw": fl(raw),
      "x": [x.tolist() for x, _ in tasks_np], "y": [yv.tolist() for _, yv in tasks_np],
      "nll_task": fl(per_task), "mean_nll": f(val), "grad": fl(grad),
      "xq": xq_np.tolist(), "alpha0": fl(alpha), "mu": fl(mu), "var": fl(var),
      "cov_full": [fl(row) for row in predict.cov_full],
      "ei": fl(ei), "pi": fl(pi), "ucb": fl(ucb),
  }


def main():
  cases = []
  cid = 0
  for cov in ("squared_exponential", "matern32", "matern52"):
    for mean in ("constant", "zero"):
      for warped in (True, False):
        rng = np.random.Generator(np.random.PCG64(4242 + cid))
        ns = [(3,), (5,), (8, 3)][cid % 3]
        d = [1, 2, 3][(cid // 3) % 3]
        cases.append(build_case(cid, cov, mean, warped, ns, d, rng))
        print("case", cid, cov, mean, warped, ns, d, cases[-1]["mean_nll"], flush=True)
        cid += 1
  # larger cases: more dimensions, ragged multi-task, duplicate points
  extra = [("squared_exponential", "constant", True, (24, 9), 4, False),
           ("matern52", "constant", True, (17, 6, 11), 2, True),
           ("matern32", "constant", True, (12,), 3, True),
           ("matern52", "zero", False, (30,), 5, False)]
  for cov, mean, warped, ns, d, dup in extra:
    rng = np.random.Generator(np.random.PCG64(4242 + cid))
    cases.append(build_case(cid, cov, mean, warped, ns, d, rng, dup))
    print("case", cid, cov, mean, warped, ns, d, cases[-1]["mean_nll"], flush=True)
    cid += 1
  with open(OUT, "w") as fh:
    json.dump({"dps": mp.mp.dps, "generator": "tests/golden/make_mpmath_kat.py",
               "cases": cases}, fh, indent=0)
  print("wrote", OUT)


if __name__ == "__main__":
  main()
