"""Shared helpers for the parity tests (oracle is imported HERE only)."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases(kl=False):
  names = sorted(os.path.basename(p)[:-4]
                 for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
  return [n for n in names if n.startswith("kl_") == kl]


def load_golden_kl(name):
  """Fixtures of the divergence objectives on aligned data (make_golden.py)."""
  z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
  g = {k: z[k] for k in z.files}
  g["cov"], g["mean"], g["d"] = str(g["cov"]), str(g["mean"]), int(g["d"])
  ds = {}
  for t, al in enumerate(g["spec_aligned"]):
    ds[t] = (g[f"x{t}"], g[f"y{t}"], 1) if al else (g[f"x{t}"], g[f"y{t}"])
  g["dataset"] = ds
  return g


def load_golden(name):
  z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
  g = {k: z[k] for k in z.files}
  g["cov"], g["mean"] = str(g["cov"]), str(g["mean"])
  g["d"] = int(g["d"])
  g["ns"] = [int(n) for n in g["ns"]]
  g["dataset"] = {t: (g[f"x{t}"], g[f"y{t}"]) for t in range(len(g["ns"]))}
  return g


def model_from_raw(raw, d, mean):
  m = {"signal_variance": float(raw[1]), "noise_variance": float(raw[2]),
       "lengthscale": np.array(raw[3:3 + d])}
  if mean == "constant":
    m["constant"] = float(raw[0])
  return m


def raw_vec(model, d):
  ls = np.broadcast_to(np.asarray(model["lengthscale"], dtype=np.float64), (d,))
  return np.concatenate([[model.get("constant", 0.0), model["signal_variance"],
                          model["noise_variance"]], ls])


def grad_vec(g, d):
  ls = np.broadcast_to(np.asarray(g["lengthscale"], dtype=np.float64), (d,))
  return np.concatenate([[g.get("constant", 0.0), g["signal_variance"],
                          g["noise_variance"]], ls])


def default_mask(d, mean="constant"):
  return 0b110 | (((1 << d) - 1) << 3)


def rel(a, b):
  a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
  return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300))


def load_kat():
  """mpmath (60-digit) known-answer cases (tests/golden/make_mpmath_kat.py --
  generated WITHOUT importing oracle/)."""
  import json
  with open(os.path.join(GOLDEN_DIR, "kat_mpmath.json")) as fh:
    cases = json.load(fh)["cases"]
  for c in cases:
    c["dataset"] = {t: (np.asarray(x, dtype=np.float64).reshape(-1, c["d"]),
                        np.asarray(y, dtype=np.float64).reshape(-1, 1))
                    for t, (x, y) in enumerate(zip(c["x"], c["y"]))}
    c["raw"] = np.asarray(c["raw"], dtype=np.float64)
    c["xq"] = np.asarray(c["xq"], dtype=np.float64).reshape(-1, c["d"])
    for k in ("nll_task", "grad", "alpha0", "mu", "var", "ei", "pi", "ucb",
              "cov_full"):
      c[k] = np.asarray(c[k], dtype=np.float64)
  return cases


def load_kat_div():
  """mpmath (60-digit) known answers for the divergence objectives on aligned data
  (tests/golden/make_mpmath_kat_div.py -- generated WITHOUT importing oracle/)."""
  import json
  with open(os.path.join(GOLDEN_DIR, "kat_mpmath_div.json")) as fh:
    cases = json.load(fh)["cases"]
  for c in cases:
    c["dataset"] = {
        "a%d" % t: (np.asarray(x, dtype=np.float64).reshape(-1, c["d"]),
                    np.asarray(y, dtype=np.float64), t + 1)
        for t, (x, y) in enumerate(zip(c["x"], c["y"]))}
    c["raw"] = np.asarray(c["raw"], dtype=np.float64)
    for k in ("kl_grad", "euc_grad"):
      c[k] = np.asarray(c[k], dtype=np.float64)
  return cases
