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
NAMES = ["L", "M", "W", "z", "alpha", "apart_rpart", "gpart", "gtask", "logdet", "nll_task"]
PER = {"L": 4096, "M": 4096, "W": 4096, "z": 64, "alpha": 64, "apart_rpart": 64, "gpart": 34, "gtask": 34, "logdet": 1, "nll_task": 1}
def snap(T):
  tiles = T * 36
  sizes = [tiles * 4096, tiles * 4096, tiles * 4096, T * 512, T * 512, 2 * tiles * 64, tiles * 34, T * 34, T * 8, T]
  out = {}
  for w, (name, cnt) in enumerate(zip(NAMES, sizes)):
    a = np.zeros(cnt, dtype=np.float64)
    eng.h.debug_read(w, a.ctypes.data, a.nbytes)
    out[name] = a
  return out
for T in (32, 48):
  ds = {t: O.make_task(t, n, d) for t in range(T)}
  pk = eng.pack([(k, v[0], v[1]) for k, v in ds.items()])
  per = np.stack([eng.nll_grad(0, 1, eng.pack([(t, ds[t][0], ds[t][1])]), raw, mask).cpu().numpy() for t in range(T)])
  # poison what earlier calls left in the workspace: a batch with other params
  eng.nll_grad(2, 1, pk, raw * 0.5, mask)
  runs, snaps = [], []
  for _ in range(3):
    runs.append(eng.nll_grad(0, 1, pk, raw, mask).cpu().numpy())
    snaps.append(snap(T))
  errs = [H.rel(r[:-1], per.sum(0)[:-1]) for r in runs]
  print("T", T, "rel err per run", [float("%.3g" % e) for e in errs], flush=True)
  if errs[0] > 1e-10:
    for name in NAMES:
      x, y = snaps[0][name], snaps[1][name]
      bad = np.nonzero(x != y)[0]
      if len(bad):
        units = sorted(set((bad // PER[name]).tolist()))
        print("   ", name, "differs in", len(bad), "elements; n_units", len(units), "units", units[:30], flush=True)
