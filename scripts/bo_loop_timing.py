"""BO-iteration latency at BASELINE configs[3] shape (query task n ~ 500, d = 4,
Matern-5/2, 10 000 candidates): the device-resident loop (hb_bo_step: acquisition
sweep + arg-max + rank-1 append, no host sync) against the reference-shaped host
loop (refactorise the task, sweep, arg-max read-back, append) -- both through
bo_utils.bayesopt.simulated_bayesopt."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyperbo_b200.basics import definitions as defs
from hyperbo_b200.bo_utils import acfun, bayesopt
from hyperbo_b200.gp_utils import gp, kernel, mean, utils

d, n0, nq, iters = 4, 500, 10000, 30
rng = np.random.default_rng(0)
def f(x): return 5 + np.sin(3 * x.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((len(x), 1))
train = {t: (rng.random((500, d)),) for t in range(23)}
train = {t: (v[0], f(v[0])) for t, v in train.items()}
xq = rng.random((nq, d)); yq = f(xq)
x0 = rng.random((n0, d)); y0 = f(x0)
model0 = {"constant": 5.1, "lengthscale": np.zeros(d), "signal_variance": 0.0, "noise_variance": -4.0}
out = {}
for mode in ("device", "host"):
  os.environ["HB_BO_DEVICE"] = "1" if mode == "device" else "0"
  times = []
  for rep in range(3):
    dataset = {k: defs.SubDataset(*v) for k, v in train.items()}
    dataset["query"] = defs.SubDataset(x0, y0)
    m = gp.GP(dataset, mean.constant, kernel.matern52, defs.GPParams(model=dict(model0), config={}),
              utils.DEFAULT_WARP_FUNC)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sub = bayesopt.simulated_bayesopt(m, "query", defs.SubDataset(xq, yq), acfun.expected_improvement, iters)
    torch.cuda.synchronize(); times.append((time.perf_counter() - t0) / iters)
    picks = np.asarray(torch.as_tensor(sub.y).cpu())[-iters:].ravel()
  out[mode] = {"us_per_bo_iteration": 1e6 * sorted(times)[1], "last_picks_y": picks[-3:].tolist()}
out["speedup"] = out["host"]["us_per_bo_iteration"] / out["device"]["us_per_bo_iteration"]
out["same_choices"] = out["host"]["last_picks_y"] == out["device"]["last_picks_y"]
out["config"] = {"n0": n0, "d": d, "candidates": nq, "iterations": iters, "kernel": "matern52", "acquisition": "EI"}
print(json.dumps(out))
