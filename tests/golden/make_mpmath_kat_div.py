"""Known answers for the divergence objectives on ALIGNED data, 60-digit arithmetic --
TEST INFRASTRUCTURE.  Generates tests/golden/kat_mpmath_div.json.  Like
make_mpmath_kat.py (whose kernel / warp helpers it reuses) this script does NOT
import oracle/ nor the product; it restates

  multivariate_normal_divergence  /root/reference/hyperbo/gp_utils/objectives.py:29-101
      mu_data = mean_q y,  cov_data = cov(y, bias=True),
      mu_model = m(x),     cov_model = K(x, x) + noise_variance I      (no jitter)
  partial KL (eps = 0)            /root/reference/hyperbo/gp_utils/utils.py:84-141
      tr(cov1^-1 cov0) + (mu1 - mu0)' cov1^-1 (mu1 - mu0) + logdet cov1
  Euclidean distance              /root/reference/hyperbo/gp_utils/utils.py:151-173
      ||mu0 - mu1||_2 + ||cov0 - cov1||_F

and differentiates both with mpmath.diff (independent of every closed form used by
the oracle, the test double and the CUDA kernels).

    python tests/golden/make_mpmath_kat_div.py      (~1 min)
"""
import json
import os
import sys

import mpmath as mp
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_mpmath_kat import fl, f, gram, q1024, unpack  # noqa: E402

mp.mp.dps = 60
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat_mpmath_div.json")


def moments(y):
  n, m = len(y), len(y[0])
  mu = [mp.fsum(row) / m for row in y]
  cov = mp.zeros(n, n)
  for i in range(n):
    for j in range(n):
      cov[i, j] = mp.fsum((y[i][q] - mu[i]) * (y[j][q] - mu[j]) for q in range(m)) / m
  return mu, cov


def model_moments(cov, mean, warped, raw, x):
  n, d = len(x), len(x[0])
  const, sv, nv, ls = unpack(raw, d, mean, warped)
  k = gram(cov, x, x, ls, sv)
  for i in range(n):
    k[i, i] += nv
  return [const] * n, k


def kl_sub(cov, mean, warped, raw, x, y):
  mu0, cov0 = moments(y)
  mu1, cov1 = model_moments(cov, mean, warped, raw, x)
  n = len(x)
  inv1 = mp.inverse(cov1)
  tr = mp.fsum((inv1 * cov0)[i, i] for i in range(n))
  dv = mp.matrix([mu1[i] - mu0[i] for i in range(n)])
  mah = (dv.T * inv1 * dv)[0, 0]
  chol = mp.cholesky(cov1)
  logdet = 2 * mp.fsum(mp.log(chol[i, i]) for i in range(n))
  return tr + mah + logdet


def euc_sub(cov, mean, warped, raw, x, y):
  mu0, cov0 = moments(y)
  mu1, cov1 = model_moments(cov, mean, warped, raw, x)
  n = len(x)
  a = mp.sqrt(mp.fsum((mu0[i] - mu1[i]) ** 2 for i in range(n)))
  b = mp.sqrt(mp.fsum((cov0[i, j] - cov1[i, j]) ** 2 for i in range(n) for j in range(n)))
  return a + b


def build_case(cid, cov, mean, warped, shapes, d, rng):
  subs_np = []
  for n, m in shapes:
    x = q1024(rng.random((n, d)))
    base = 5.0 + np.sin(3.0 * x.sum(axis=1))
    y = q1024(base[:, None] + 0.4 * rng.standard_normal((n, m)))
    subs_np.append((x, y))
  if warped:
    raw_np = q1024(np.concatenate([[5.1], rng.normal(0, 0.5, 1), [-2.0 + rng.normal(0, 0.3)],
                                   rng.normal(0, 0.5, d)]))
  else:
    raw_np = q1024(np.concatenate([[4.9], [0.8 + 0.5 * rng.random()],
                                   [0.05 + 0.1 * rng.random()], 0.4 + rng.random(d)]))
  raw = [mp.mpf(float(v)) for v in raw_np]
  subs = [([[mp.mpf(float(v)) for v in row] for row in x],
           [[mp.mpf(float(v)) for v in row] for row in y]) for x, y in subs_np]
  out = {"id": cid, "cov": cov, "mean": mean, "warped": bool(warped), "d": d,
         "shapes": [list(s) for s in shapes], "raw": fl(raw),
         "x": [x.tolist() for x, _ in subs_np], "y": [y.tolist() for _, y in subs_np]}
  for name, fn in (("kl", kl_sub), ("euc", euc_sub)):
    def total(r):
      return mp.fsum(fn(cov, mean, warped, r, x, y) for x, y in subs) / len(subs)
    out[name] = f(total(raw))
    grad = []
    for p in range(3 + d):
      if p == 0 and mean == "zero":
        grad.append(mp.mpf(0))
        continue
      def fp(v, p=p):
        r2 = list(raw)
        r2[p] = v
        return total(r2)
      grad.append(mp.diff(fp, raw[p], h=mp.mpf("1e-12")))
    out[name + "_grad"] = fl(grad)
  return out


def main():
  cases, cid = [], 0
  plan = [("squared_exponential", "constant", True, [(6, 4)], 2),
          ("matern32", "constant", True, [(5, 3), (7, 5)], 1),
          ("matern52", "constant", False, [(8, 6)], 3),
          ("matern52", "zero", True, [(6, 3), (4, 4)], 2),
          ("squared_exponential", "zero", False, [(7, 2)], 2),
          ("matern32", "constant", False, [(9, 12)], 2)]
  for cov, mean, warped, shapes, d in plan:
    rng = np.random.Generator(np.random.PCG64(777 + cid))
    cases.append(build_case(cid, cov, mean, warped, shapes, d, rng))
    print("case", cid, cov, mean, warped, shapes, d, cases[-1]["kl"], cases[-1]["euc"],
          flush=True)
    cid += 1
  with open(OUT, "w") as fh:
    json.dump({"dps": mp.mp.dps, "generator": "tests/golden/make_mpmath_kat_div.py",
               "cases": cases}, fh, indent=0)
  print("wrote", OUT)


if __name__ == "__main__":
  main()
