"""compute-sanitizer --tool initcheck driver: one tiny call of each late round-2 entry."""
import numpy as np, torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyperbo_b200.engine import Engine
eng = Engine.get()
rng = np.random.default_rng(0)
d, R, ns = 3, 5, [70, 20]
tasks = [(t, rng.uniform(size=(n, d)), rng.standard_normal(n)) for t, n in enumerate(ns)]
ds = eng.pack(tasks)
raw = np.array([0.3, 0.2, -3.0, 0.1, 0.0, -0.1]); mask = 0b111110
B = rng.standard_normal(sum(ns) * R); cw = rng.uniform(0.5, 1, size=(2, R))
cm = np.array([0, 0, 0, 0, 1], dtype=np.int32)
print("mrhs", eng.nll_grad_mrhs(2, 1, ds, R, B, cw, cm, raw, mask).cpu().numpy()[:3])
print("euc", eng.euclid_grad(2, 1, ds, R, 0.3 * B, ds.y, raw, mask).cpu().numpy()[:3])
x, y = tasks[0][1], tasks[0][2]
cache, chol, alpha, nll, info = eng.build_predictor(2, 1, x, y, raw, mask)
mu, cov = eng.predict_cov(2, 1, eng.tensor(x), cache, raw, mask, rng.uniform(size=(30, d)), 1.0, 1.5)
print("cov", float(cov[0, 0]), float(cov[29, 3]))
