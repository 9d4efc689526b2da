"""initcheck control experiment: the round-1 launch-per-column path (k_alpha reads the M
tiles that k_step wrote with cp.async.bulk shared->global copies)."""
import numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyperbo_b200.engine import Engine
eng = Engine.get()
eng.h.set_option("fused", 0)
rng = np.random.default_rng(0)
d = 3
tasks = [(t, rng.uniform(size=(n, d)), rng.standard_normal(n)) for t, n in enumerate([300, 280])]
ds = eng.pack(tasks)
raw = np.array([0.3, 0.2, -3.0, 0.1, 0.0, -0.1]); mask = 0b111110
print("nll", eng.nll_grad(2, 1, ds, raw, mask).cpu().numpy()[:3])
