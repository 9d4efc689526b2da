"""Debug: phase timeline of the look-ahead CTA (role 0) of k_step (HB_STAMPS build)."""
import ctypes, os, subprocess, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = "/tmp/libhb_stamps0.so"
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
offs = (ctypes.c_int64 * (T + 1))(*[n * t for t in range(T + 1)])
mask = 0b110 | (((1 << d) - 1) << 3)
nll = torch.zeros(T, device="cuda", dtype=torch.float64)
alpha = torch.zeros(T * n, device="cuda", dtype=torch.float64)
def P(t): return ctypes.c_void_p(t.data_ptr())
for _ in range(3):
  assert lib.hb_factorize_batched(h, 0, 1, T, offs, d, P(x), P(y), P(raw), ctypes.c_uint64(mask), None, P(alpha), P(nll), None, None) == 0
torch.cuda.synchronize()
nb = 8
buf = np.zeros(((nb + 1) * T * 8, 8), dtype=np.int64)
lib.hb_debug_stamps(h, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)), ctypes.c_int64(buf.size))
buf = buf.reshape(nb + 1, T, 8, 8)
for jj in range(nb):
  r = buf[jj][:, 0, :]
  r = r[r[:, 0] != 0]
  if len(r) == 0: continue
  s = r[:, :7].astype(np.float64)
  tail = (r[:, 7] >> 8).astype(np.float64)
  if jj == 0:
    print("j=-1: keval+... %.0f potrf64 %.0f tail %.0f" % ((s[:,5]-s[:,0]).mean(), (s[:,6]-s[:,5]).mean(), tail.mean()))
    continue
  print("j=%d: panel-stream %.0f | xchg+pre+keval %.0f | trsm+xchg+store %.0f | syrk+diag-stream %.0f | xchg+pre+keval %.0f | potrf64 %.0f | logdet+z+store %.0f | total %.0f"
        % (jj - 1, (s[:,1]-s[:,0]).mean(), (s[:,2]-s[:,1]).mean(), (s[:,3]-s[:,2]).mean(), (s[:,4]-s[:,3]).mean(),
           (s[:,5]-s[:,4]).mean(), (s[:,6]-s[:,5]).mean(), tail.mean(), (s[:,6]-s[:,0]+tail).mean()))
  # the other roles of this launch: durations by kind
  for kind, name in ((1, "panel"), (2, "trtri")):
    rr = buf[jj].reshape(-1, 8)
    m = ((rr[:, 7] & 255) == kind) & (rr[:, 0] != 0)
    if m.any():
      q = rr[m]
      last = np.where(kind == 1, q[:, 3], q[:, 2])
      print("      %s n=%d mean duration %.0f max %.0f" % (name, m.sum(), (last - q[:, 0]).mean(), (last - q[:, 0]).max()))
