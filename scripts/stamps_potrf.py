"""Debug: phase breakdown of the in-CTA 64x64 diagonal factorisation (HB_STAMPS build)."""
import ctypes, os, subprocess, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = "/tmp/libhb_stamps.so"
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17",
                       "-Xcompiler", "-fPIC", "-shared", "-DHB_STAMPS", "-DHB_STAMPS_POTRF", "-o", so,
                       *[os.path.join(ROOT, "hyperbo_b200/csrc", u)
                         for u in ("hb_capi.cu", "hb_f64.cu", "hb_f32.cu")]])
lib = ctypes.CDLL(so)
h = ctypes.c_void_p()
assert lib.hb_create(ctypes.byref(h), 0, 0) == 0
T = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n, d = 512, 8
rng = np.random.default_rng(0)
x = torch.as_tensor(rng.random((T * n, d)), device="cuda")
y = torch.as_tensor(5 + rng.standard_normal(T * n), device="cuda")
raw = torch.tensor([5.1, 0, -4] + [0.0] * d, device="cuda", dtype=torch.float64)
offs = (ctypes.c_int64 * (T + 1))(*[n * t for t in range(T + 1)])
mask = 0b110 | (((1 << d) - 1) << 3)
nll = torch.zeros(T, device="cuda", dtype=torch.float64)
def P(t): return ctypes.c_void_p(t.data_ptr())
for _ in range(3):  # factorise-only path: no lauum stamps overwrite the buffer
  assert lib.hb_factorize_batched(h, 0, 1, T, offs, d, P(x), P(y), P(raw), ctypes.c_uint64(mask), None, None, P(nll), None, None) == 0
torch.cuda.synchronize()
nb = 8
buf = np.zeros(((nb + 1) * T * 8, 8), dtype=np.int64)
lib.hb_debug_stamps(h, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)), ctypes.c_int64(buf.size))
buf = buf.reshape(nb + 1, T, 8, 8)
for jj in range(nb):
  r = buf[jj][:, 0, :]      # role 0 (look-ahead) of every task
  r = r[r[:, 0] != 0]
  if len(r) == 0: continue
  tot = r[:, 6] - r[:, 5]
  print("j=%d diag blocks=%d  potrf64 total %.0f  [potrf16 %.0f | trsm+trtri16 %.0f | syrk16 %.0f | inverse+zero %.0f]  whole CTA %.0f"
        % (jj - 1, len(r), tot.mean(), r[:, 1].mean(), r[:, 2].mean(), r[:, 3].mean(), r[:, 4].mean(), (r[:, 6] - r[:, 0]).mean()))
