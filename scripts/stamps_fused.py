"""Debug: per-item timeline of the persistent kernel k_fused (HB_STAMPS build).

  python scripts/stamps_fused.py [T] [out.npy]
"""
import ctypes, os, subprocess, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = "/tmp/libhb_stamps_fused.so"
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17",
                       "-Xcompiler", "-fPIC", "-shared", "-DHB_STAMPS", "-o", so,
                       *[os.path.join(ROOT, "hyperbo_b200/csrc", u)
                         for u in ("hb_capi.cu", "hb_f64.cu", "hb_f32.cu")]])
lib = ctypes.CDLL(so)
h = ctypes.c_void_p()
assert lib.hb_create(ctypes.byref(h), 0, 0) == 0
T = int(sys.argv[1]) if len(sys.argv) > 1 else 32
n, d = 512, 8
rng = np.random.default_rng(0)
x = torch.as_tensor(rng.random((T * n, d)), device="cuda")
y = torch.as_tensor(5 + rng.standard_normal(T * n), device="cuda")
raw = torch.tensor([5.1, 0, -4] + [0.0] * d, device="cuda", dtype=torch.float64)
sums = torch.zeros(3 + d + 2, device="cuda", dtype=torch.float64)
offs = (ctypes.c_int64 * (T + 1))(*[n * t for t in range(T + 1)])
mask = 0b110 | (((1 << d) - 1) << 3)
def P(t): return ctypes.c_void_p(t.data_ptr())
for _ in range(3):
  assert lib.hb_nll_grad_batched(h, 0, 1, T, offs, d, P(x), P(y), P(raw), ctypes.c_uint64(mask), P(sums), None, None, None) == 0
torch.cuda.synchronize()
nitems = T * (8 + 28 + 28 + 36 + 1)
buf = np.zeros((nitems, 8), dtype=np.int64)
lib.hb_debug_stamps(h, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)), ctypes.c_int64(buf.size))
if len(sys.argv) > 2:
  np.save(sys.argv[2], buf)
t0 = buf[:, 0].min()
task = buf[:, 7] >> 32
kind = (buf[:, 7] >> 24) & 255
ia = (buf[:, 7] >> 12) & 4095
ib = buf[:, 7] & 4095
names = ["DIAG", "PANEL", "TRTRI", "LAUUM"]
print("kernel span %.1f us, items %d" % ((buf[:, 5].max() - t0) / 1e3, nitems))
for k in range(4):
  m = kind == k
  dur = (buf[m, 5] - buf[m, 0]) / 1e3
  print("%s: n=%d mean %.1f us  max %.1f  sum %.0f us" % (names[k], m.sum(), dur.mean(), dur.max(), dur.sum()))
# chain of task 0
print("task 0 timeline (us since kernel start): start | pre-stream | deps/stream end | (flagD) | potrf end | end")
m = task == 0
order = np.argsort(buf[m, 5])
rows = buf[m][order]
for r in rows:
  k = (r[7] >> 24) & 255
  if k == 3: continue
  a, b = (r[7] >> 12) & 4095, r[7] & 4095
  print("%-5s (%d,%d) sm%3d  " % (names[k], a, b, r[6]) + " ".join("%8.1f" % ((v - t0) / 1e3) if v else "       -" for v in r[:6]))
lm = m & (kind == 3)
print("task 0 LAUUM: first start %.1f last end %.1f" % ((buf[lm, 0].min() - t0) / 1e3, (buf[lm, 5].max() - t0) / 1e3))
