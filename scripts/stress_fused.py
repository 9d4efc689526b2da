"""Stress: many (other batch -> batch A -> batch A) triples in the few-task
regime; every A call must give bit-identical sums (the arithmetic is
deterministic), whatever ran before it."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from hyperbo_b200.engine import Engine
eng = Engine.get()
n, d = 512, 8
raw = np.concatenate([[5.1, 0.0, -4.0], np.linspace(-0.3, 0.4, d)])
mask = 0b110 | (((1 << d) - 1) << 3)
ntr = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(0)
fails = 0
for T in (8, 24, 32, 48):
  x = rng.random((2, T, n, d)); y = 5 + rng.standard_normal((2, T, n, 1))
  pkA = eng.pack([(t, x[0, t], y[0, t]) for t in range(T)])
  pkB = eng.pack([(t, x[1, t], y[1, t]) for t in range(T)])
  ref = eng.nll_grad(0, 1, pkA, raw, mask).cpu().numpy()
  bad = 0
  for trial in range(ntr):
    if trial % 3 == 0: eng.nll_grad(2, 1, pkA, raw * 0.5, mask)
    elif trial % 3 == 1: eng.nll_grad(0, 1, pkB, raw * 0.9, mask)
    else: eng.nll_grad(1, 1, pkB, raw * 1.3, mask).cpu()
    a = eng.nll_grad(0, 1, pkA, raw, mask).cpu().numpy()
    if not np.array_equal(a, ref):
      bad += 1
      if bad <= 3: print("  T", T, "trial", trial, "rel", np.max(np.abs(a - ref)) / np.max(np.abs(ref)), flush=True)
  print("T", T, "failures", bad, "of", ntr, flush=True)
  fails += bad
print("TOTAL failures", fails)
