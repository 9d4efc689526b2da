"""GP.train(method='lbfgs') wall time: memo of accepted points (always on) and the
speculative line search (HB_LBFGS_SPECULATE=0/1).  Writes gpurun_out/lbfgs_timing.json."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyperbo_b200.basics import definitions as defs, lbfgs as _lbfgs  # noqa: E402
from hyperbo_b200.gp_utils import gp, kernel, mean, utils  # noqa: E402
from oracle import hyperbo_oracle as O  # noqa: E402  (data generator only)

def run(T, n, d, steps, speculate, memo=True):
  os.environ["HB_LBFGS_SPECULATE"] = "1" if speculate else "0"
  ds = {t: defs.SubDataset(*O.make_task(t, n, d, "matern52")) for t in range(T)}
  params = defs.GPParams(model=dict(O.init_raw_params(d)),
                         config={"method": "lbfgs", "max_training_step": steps,
                                 "batch_size": 10**6, "alpha": 1.0})
  m = gp.GP(dataset=ds, mean_func=mean.constant, cov_func=kernel.matern52,
            params=params, warp_func=utils.DEFAULT_WARP_FUNC)
  keep = _lbfgs._Evaluator.__init__.__defaults__
  if not memo:   # reference behaviour: every point is evaluated again
    _lbfgs._Evaluator.__init__.__defaults__ = (None, 0)
  torch.cuda.synchronize(); t0 = time.perf_counter()
  out = m.train()
  torch.cuda.synchronize(); dt = time.perf_counter() - t0
  _lbfgs._Evaluator.__init__.__defaults__ = keep
  return dt, {k: np.asarray(v).tolist() for k, v in out.model.items()}

res = []
for T, n, d in ((12, 150, 4), (24, 500, 4), (256, 512, 8)):
  run(T, n, d, 3, False)  # warm-up (plans, allocations)
  row = {"tasks": T, "n": n, "d": d, "steps": 20}
  for name, spec, memo in (("no_memo", False, False), ("memo", False, True),
                           ("memo+speculate", True, True)):
    dt, model = run(T, n, d, 20, spec, memo)
    row[name + "_s"] = dt
    row[name + "_noise_variance"] = model["noise_variance"]
  res.append(row)
  print(json.dumps(row))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/lbfgs_timing.json", "w"), indent=1)
