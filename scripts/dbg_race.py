"""Debug: hunt the first-call-after-poison race: compare every internal buffer of
a failing call with the same call repeated."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.path.join(ROOT, "hyperbo_b200", "libhyperbo_b200.so"))
lib.hb_debug_read.restype = ctypes.c_int64
h = ctypes.c_void_p()
assert lib.hb_create(ctypes.byref(h), 0, 0) == 0
P = lambda t: ctypes.c_void_p(t.data_ptr())
n, d = 512, 8
mask = 0b110 | (((1 << d) - 1) << 3)
NAMES = ["L", "M", "W", "z", "alpha", "apart_rpart", "gpart", "gtask", "logdet", "nll_task"]
def call(T, seed, kid, scale):
  rng = np.random.default_rng(seed)
  x = torch.as_tensor(rng.random((T * n, d)), device="cuda")
  y = torch.as_tensor(5 + rng.standard_normal(T * n), device="cuda")
  raw = torch.as_tensor(np.concatenate([[5.1, 0.0, -4.0], np.linspace(-0.3, 0.4, d)]) * scale, device="cuda")
  sums = torch.zeros(3 + d + 2, device="cuda", dtype=torch.float64)
  offs = (ctypes.c_int64 * (T + 1))(*[n * t for t in range(T + 1)])
  rc = lib.hb_nll_grad_batched(h, kid, 1, T, offs, d, P(x), P(y), P(raw), ctypes.c_uint64(mask), P(sums), None, None, None)
  torch.cuda.synchronize()
  return sums.cpu().numpy()
def snap(T):
  out = {}
  tiles = T * 36
  sizes = [tiles * 4096, tiles * 4096, tiles * 4096, T * 512, T * 512, 2 * tiles * 64, tiles * 34, T * 34, T * 8, T]
  for w, (name, cnt) in enumerate(zip(NAMES, sizes)):
    a = np.zeros(cnt, dtype=np.float64)
    lib.hb_debug_read(h, w, a.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(a.nbytes))
    out[name] = a
  return out
T = int(sys.argv[1]) if len(sys.argv) > 1 else 32
nfail = 0
for trial in range(30):
  call(T, 100 + trial, 2, 0.5)            # poison (other data, kernel, params)
  a = call(T, 7, 0, 1.0); sa = snap(T)
  b = call(T, 7, 0, 1.0); sb = snap(T)
  err = np.max(np.abs(a - b)) / np.max(np.abs(b))
  if err > 1e-12:
    nfail += 1
    print("trial", trial, "FAIL rel", err, "sums diff idx", np.nonzero(np.abs(a - b) > 1e-9 * np.abs(b).max())[0].tolist(), flush=True)
    for name in NAMES:
      x, y = sa[name], sb[name]
      bad = np.nonzero(~(np.abs(x - y) <= 1e-12 * (np.abs(y) + 1e-300)) & ~((x == y)))[0]
      if len(bad):
        per = {"L": 4096, "M": 4096, "W": 4096, "z": 64, "alpha": 64, "apart_rpart": 64, "gpart": 34, "gtask": 34, "logdet": 1, "nll_task": 1}[name]
        units = sorted(set((bad // per).tolist()))
        print("   ", name, "differs in", len(bad), "elements; units", units[:20], "(unit = %d elems)" % per, flush=True)
    if nfail >= 3: break
print("failures", nfail)
