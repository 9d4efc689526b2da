"""Debug: per-CTA phase timeline of k_step (build with -DHB_STAMPS; potrf+trtri path)."""
import ctypes, os, subprocess
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
  assert lib.hb_nll_grad_batched(h, 0, 1, T, offs, d, P(x), P(y), P(raw), ctypes.c_uint64(mask), P(sums), None, None, None) == 0
torch.cuda.synchronize()
nb = 8
buf = np.zeros(((nb + 1) * T * 8, 8), dtype=np.int64)
lib.hb_debug_stamps(h, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)), ctypes.c_int64(buf.size))
buf = buf.reshape(nb + 1, T, 8, 8)
for jj in range(nb + 1):
  j = jj - 1
  b = buf[jj]
  out = []
  for kind, name in ((0, "diag"), (1, "panel"), (2, "trtri")):
    m = (b[:, :, 7] == kind) & (b[:, :, 0] != 0)
    if kind == 0 and j >= 0: m = m & (b[:, :, 5] != 0)
    if not m.any(): continue
    r = b[m]
    if kind == 1:
      out.append("panel n=%d stream %.0f keval %.0f trsm+store %.0f" % (m.sum(), (r[:,1]-r[:,0]).mean(), (r[:,2]-r[:,1]).mean(), (r[:,3]-r[:,2]).mean()))
    elif kind == 2:
      out.append("trtri n=%d stream %.0f rest %.0f" % (m.sum(), (r[:,1]-r[:,0]).mean(), (r[:,2]-r[:,1]).mean()))
    else:
      st = np.where(r[:,3] != 0, r[:,3], r[:,0])
      out.append("diag n=%d panelpart %.0f syrkstream %.0f keval %.0f potrf64 %.0f" % (m.sum(), (st-r[:,0]).mean(), (r[:,4]-st).mean(), (r[:,5]-r[:,4]).mean(), (r[:,6]-r[:,5]).mean()))
  print("j=%d: " % j + " | ".join(out))
