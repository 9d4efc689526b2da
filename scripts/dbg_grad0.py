import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from oracle import hyperbo_oracle as O
from tests import helpers as H
from hyperbo_b200.engine import Engine
eng = Engine.get()
n, d = 512, 8
model = O.init_raw_params(d)
model["lengthscale"] = np.linspace(-0.3, 0.4, d)
raw, mask = H.raw_vec(model, d), H.default_mask(d)
for T in (32, 48):
  ds = {t: O.make_task(t, n, d) for t in range(T)}
  pk = eng.pack([(k, v[0], v[1]) for k, v in ds.items()])
  per = np.stack([eng.nll_grad(0, 1, eng.pack([(t, ds[t][0], ds[t][1])]), raw, mask).cpu().numpy() for t in range(T)])
  eng.nll_grad(2, 1, pk, raw * 0.5, mask)
  runs = [eng.nll_grad(0, 1, pk, raw, mask).cpu().numpy() for _ in range(3)]
  print("T", T, "rel err per run", [float("%.3g" % H.rel(r[:-1], per.sum(0)[:-1])) for r in runs], flush=True)
