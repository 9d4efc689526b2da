"""cProfile of GP(...).train() (Adam, 256 x 512 x 8, K steps): where the fixed host
overhead of the reference-shaped entry point goes."""
import cProfile, pstats, sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyperbo_b200.basics import definitions as defs
from hyperbo_b200.gp_utils import gp as gpm, kernel, mean, objectives, utils
T, n, d, K = 256, 512, 8, int(sys.argv[1]) if len(sys.argv) > 1 else 20
rng = np.random.default_rng(0)
dataset = {t: defs.SubDataset(rng.random((n, d)), 5 + rng.standard_normal((n, 1))) for t in range(T)}

def train_once(steps):
  params = defs.GPParams(
      model={"constant": 5.1, "lengthscale": np.zeros(d), "signal_variance": 0.0,
             "noise_variance": -4.0},
      config={"method": "adam", "learning_rate": 1e-3, "max_training_step": steps,
              "batch_size": n + 1, "objective": objectives.nll})
  model = gpm.GP(dataset, mean.constant, kernel.squared_exponential, params,
                 utils.DEFAULT_WARP_FUNC)
  torch.cuda.synchronize(); t0 = time.perf_counter()
  model.train(key=0)
  torch.cuda.synchronize()
  return time.perf_counter() - t0

train_once(3)
print("wall K=%d: %.1f ms" % (K, 1e3 * train_once(K)))
print("wall K=0-ish (1 step): %.1f ms" % (1e3 * train_once(1)))
pr = cProfile.Profile(); pr.enable(); train_once(K); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
