"""Profiling target: W warm-up calls + 1 call of hb_nll_grad_batched on the
bench workload (T x 512 x 8, fp64 unless --f32), through the Python engine.
14 launches per call at T = 256 (k_prep, 9 x k_step, k_alpha, k_lauum_grad,
k_reduce_task, k_reduce_final):  ncu -s 14*W -c 14."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyperbo_b200.engine import Engine, PackedDataset  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 256
W = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dt = torch.float32 if "--f32" in sys.argv else torch.float64
n, d = 512, 8
rng = np.random.default_rng(0)
eng = Engine.get(0, dtype=dt)
x = torch.as_tensor(rng.random((T * n, d)), device="cuda", dtype=dt)
y = torch.as_tensor(5 + rng.standard_normal(T * n), device="cuda", dtype=dt)
ds = PackedDataset(list(range(T)), x, y, [n * t for t in range(T + 1)])
raw = np.array([5.1, 0.0, -4.0] + [0.0] * d)
mask = 0b110 | (((1 << d) - 1) << 3)
for _ in range(W + 1):
  sums = eng.nll_grad(0, 1, ds, raw, mask)
torch.cuda.synchronize()
print("mean nll", float(sums[0] / sums[-1]))
