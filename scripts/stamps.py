"""Debug: per-CTA phase timeline of k_lauum_grad (build with -DHB_STAMPS)."""
import ctypes, os, subprocess, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = "/tmp/libhb_stamps.so"
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17",
                       "-Xcompiler", "-fPIC", "-shared", "-DHB_STAMPS", "-o", so,
                       *[os.path.join(ROOT, "hyperbo_b200/csrc", u)
                         for u in ("hb_capi.cu", "hb_f64.cu", "hb_f32.cu")]])
lib = ctypes.CDLL(so)
h = ctypes.c_void_p()
assert lib.hb_create(ctypes.byref(h), 0, 0) == 0
T, n, d = 256, 512, 8
rng = np.random.default_rng(0)
x = torch.as_tensor(rng.random((T * n, d)), device="cuda")
y = torch.as_tensor(5 + rng.standard_normal(T * n), device="cuda")
raw = torch.tensor([5.1, 0, -4] + [0.0] * d, device="cuda", dtype=torch.float64)
sums = torch.zeros(3 + d + 2, device="cuda", dtype=torch.float64)
offs = (ctypes.c_int64 * (T + 1))(*[n * t for t in range(T + 1)])
mask = 0b110 | (((1 << d) - 1) << 3)
def P(t): return ctypes.c_void_p(t.data_ptr())
for _ in range(3):
  rc = lib.hb_nll_grad_batched(h, 0, 1, T, offs, d, P(x), P(y), P(raw), ctypes.c_uint64(mask), P(sums), None, None, None)
  assert rc == 0
torch.cuda.synchronize()
ntile = 36
buf = np.zeros((T * ntile, 8), dtype=np.int64)
lib.hb_debug_stamps(h, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)), ctypes.c_int64(buf.size))
t0 = buf[:, 0]
mma = buf[:, 2] - buf[:, 0]; exch = buf[:, 3] - buf[:, 2]; epi = buf[:, 4] - buf[:, 3]; red = buf[:, 5] - buf[:, 4]
klen = buf[:, 7]
print("cycles (mean): start->mma_done %.0f  exchange %.0f  epilogue %.0f  reduce %.0f  total %.0f" % (mma.mean(), exch.mean(), epi.mean(), red.mean(), (buf[:,5]-buf[:,0]).mean()))
for k in range(1, 9):
  m = klen == k
  print(" k=%d tiles: n=%d  mma %.0f  exch %.0f  epi %.0f  red %.0f" % (k, m.sum(), mma[m].mean(), exch[m].mean(), epi[m].mean(), red[m].mean()))
span = buf[:, 5].max() - buf[:, 0].min()
print("kernel span cycles", span, " sum CTA cycles / (span*296) = %.2f" % ((buf[:,5]-buf[:,0]).sum() / (span * 296.0)))
