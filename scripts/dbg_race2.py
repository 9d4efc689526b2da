"""Debug: the dbg_grad scenario (Engine path, same packed batch for the poison
call and the checked call) with buffer dumps on failure."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from oracle import hyperbo_oracle as O
from tests import helpers as H
from hyperbo_b200.engine import Engine
eng = Engine.get()
n, d = 512, 8
T = int(sys.argv[1]) if len(sys.argv) > 1 else 32
MODE = sys.argv[2] if len(sys.argv) > 2 else "same"
model = O.init_raw_params(d)
model["lengthscale"] = np.linspace(-0.3, 0.4, d)
raw, mask = H.raw_vec(model, d), H.default_mask(d)
ds = {t: O.make_task(t, n, d) for t in range(T)}
pk = eng.pack([(k, v[0], v[1]) for k, v in ds.items()])
NAMES = ["L", "M", "W", "z", "alpha", "apart_rpart", "gpart", "gtask", "logdet", "nll_task"]
PER = {"L": 4096, "M": 4096, "W": 4096, "z": 64, "alpha": 64, "apart_rpart": 64, "gpart": 34, "gtask": 34, "logdet": 1, "nll_task": 1}
def snap():
  tiles = T * 36
  sizes = [tiles * 4096, tiles * 4096, tiles * 4096, T * 512, T * 512, 2 * tiles * 64, tiles * 34, T * 34, T * 8, T]
  out = {}
  for w, (name, cnt) in enumerate(zip(NAMES, sizes)):
    a = np.zeros(cnt, dtype=np.float64)
    eng.h.debug_read(w, a.ctypes.data, a.nbytes)
    out[name] = a
  return out
nfail = 0
for trial in range(24):
  if MODE == "fresh":   # a new plan every trial: T changes
    T = 24 + trial
    pk = eng.pack([(k, ds[k % 32][0], ds[k % 32][1]) for k in range(T)])
  if MODE == "singles":
    for t in range(8):
      eng.nll_grad(0, 1, eng.pack([(t, ds[t][0], ds[t][1])]), raw, mask).cpu()
  eng.nll_grad(2, 1, pk, raw * (0.5 + 0.01 * trial), mask)     # poison, not synced
  a = eng.nll_grad(0, 1, pk, raw, mask).cpu().numpy(); sa = snap()
  b = eng.nll_grad(0, 1, pk, raw, mask).cpu().numpy(); sb = snap()
  err = np.max(np.abs(a - b)) / np.max(np.abs(b))
  if err > 1e-12:
    nfail += 1
    print("trial", trial, "FAIL rel", err, "sums idx", np.nonzero(np.abs(a - b) > 1e-9 * np.abs(b).max())[0].tolist(), flush=True)
    for name in NAMES:
      x, y = sa[name], sb[name]
      bad = np.nonzero(x != y)[0]
      if len(bad):
        units = sorted(set((bad // PER[name]).tolist()))
        print("   ", name, "differs in", len(bad), "elements; units", units[:24], "n_units", len(units), flush=True)
    if nfail >= 3: break
print("failures", nfail)
