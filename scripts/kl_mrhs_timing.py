"""Empirical-KL value + gradient (objectives.kl) per call: hb_nll_grad_mrhs (one
factorisation per aligned sub-dataset) against the m + 2 weighted-task
decomposition.  Writes gpurun_out/kl_mrhs.json."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyperbo_b200.basics import definitions as defs, params_utils  # noqa: E402
from hyperbo_b200.gp_utils import kernel, mean, objectives, utils  # noqa: E402


def bench(n, m, d, nsub, reps=20):
  rng = np.random.default_rng(0)
  dataset = {}
  for s in range(nsub):
    x = rng.uniform(size=(n, d))
    y = np.sin(x.sum(1))[:, None] + 0.3 * rng.standard_normal((n, m))
    dataset[f"a{s}"] = defs.SubDataset(x, y, aligned=s + 1)
  model = {"constant": 0.1, "signal_variance": 0.0, "noise_variance": -3.0,
           "lengthscale": np.zeros(d)}
  out = {"n": n, "m": m, "d": d, "aligned_sub_datasets": nsub}
  vals = {}
  for flag in (True, False):
    objectives.KL_MULTI_RHS = flag
    prog = objectives.compile_objective(objectives.kl, mean.constant,
                                        kernel.matern52, dataset)
    raw, mask, _ = params_utils.pack_raw(model, d, True, utils.DEFAULT_WARP_FUNC)
    raw = prog.eng.tensor(raw)
    for _ in range(3):
      s = prog.sums(raw, mask)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
      s = prog.sums(raw, mask)
    e1.record()
    torch.cuda.synchronize()
    key = "multi_rhs" if flag else "weighted_tasks"
    out[key + "_ms"] = e0.elapsed_time(e1) / reps
    vals[key] = s.cpu().numpy()
  out["speedup"] = out["weighted_tasks_ms"] / out["multi_rhs_ms"]
  a, b = vals["multi_rhs"], vals["weighted_tasks"]
  out["max_rel_diff"] = float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
  objectives.KL_MULTI_RHS = True
  return out


if __name__ == "__main__":
  res = [bench(512, 20, 8, 1), bench(512, 20, 8, 8), bench(2048, 20, 8, 1),
         bench(1024, 50, 4, 2)]
  os.makedirs("gpurun_out", exist_ok=True)
  with open("gpurun_out/kl_mrhs.json", "w") as f:
    json.dump(res, f, indent=1)
  for r in res:
    print(json.dumps(r))
