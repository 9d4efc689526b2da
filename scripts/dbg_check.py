"""Debug: first call after a different batch, with the fast-path verifier."""
import ctypes, os, subprocess, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = "/tmp/libhb_verify.so"
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17",
                       "-Xcompiler", "-fPIC", "-shared", "-DHB_FUSED_VERIFY", "-o", so,
                       *[os.path.join(ROOT, "hyperbo_b200/csrc", u) for u in ("hb_capi.cu", "hb_f64.cu", "hb_f32.cu")]])
lib = ctypes.CDLL(so)
h = ctypes.c_void_p()
assert lib.hb_create(ctypes.byref(h), 0, 0) == 0
P = lambda t: ctypes.c_void_p(t.data_ptr())
n, d = 512, 8
mask = 0b110 | (((1 << d) - 1) << 3)
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
for T in (32, 48):
  call(T, 1, 2, 0.5)               # poison
  a = call(T, 2, 0, 1.0)
  b = call(T, 2, 0, 1.0)
  print("T", T, "first vs second call rel diff", np.max(np.abs(a - b)) / np.max(np.abs(b)), flush=True)
