"""Small end-to-end case for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyperbo_b200.engine import Engine
eng = Engine.get()
rng = np.random.default_rng(0)
d = 3
ns = [130, 64, 200]
tasks = [(t, rng.random((n, d)), 5 + rng.standard_normal((n, 1))) for t, n in enumerate(ns)]
ds = eng.pack(tasks)
raw = np.array([5.1, 0.0, -4.0] + [0.0] * d)
mask = 0b110 | (((1 << d) - 1) << 3)
for kid in (0, 2):
  sums = eng.nll_grad(kid, 1, ds, raw, mask)
  chols, alpha, nll, info = eng.factorize(kid, 1, ds, raw, mask)
cache, chol, kinvy, _, _ = eng.build_predictor(2, 1, tasks[0][1], tasks[0][2], raw, mask)
mu, var, acq = eng.predict(2, 1, eng.tensor(tasks[0][1]), cache, raw, mask, rng.random((100, d)), 1.0, 1.5, 1, 5.0)
k = eng.kernel_matrix(1, tasks[0][1], tasks[1][1], raw, mask)
torch.cuda.synchronize()
print("done", float(sums[0]), float(mu.sum()), float(k.sum()))
