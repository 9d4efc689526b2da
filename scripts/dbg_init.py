import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import ctypes, os, subprocess
ROOT = '/root/repo'
so = "/tmp/libhb_nospec.so"
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
                       "-Xcompiler", "-fPIC", "-shared", "-DHB_NO_SPECULATION", "-o", so,
                       *[os.path.join(ROOT, "hyperbo_b200/csrc", u) for u in ("hb_capi.cu", "hb_f64.cu", "hb_f32.cu")]])
lib = ctypes.CDLL(so)
h = ctypes.c_void_p()
assert lib.hb_create(ctypes.byref(h), 0, 0) == 0
n, d, T = 512, 8, int(sys.argv[1]) if len(sys.argv) > 1 else 32
rng = np.random.default_rng(0)
raw = np.concatenate([[5.1, 0.0, -4.0], np.linspace(-0.3, 0.4, d)])
mask = 0b110 | (((1 << d) - 1) << 3)
x = rng.random((T, n, d)); y = 5 + rng.standard_normal((T, n, 1))
xt = torch.as_tensor(x.reshape(-1, d), device="cuda"); yt = torch.as_tensor(y.reshape(-1), device="cuda")
rt = torch.as_tensor(raw, device="cuda"); sums = torch.zeros(3 + d + 2, device="cuda", dtype=torch.float64)
offs = (ctypes.c_int64 * (T + 1))(*[n * t for t in range(T + 1)])
P = lambda t: ctypes.c_void_p(t.data_ptr())
for rep in range(2):
  rc = lib.hb_nll_grad_batched(h, 0, 1, T, offs, d, P(xt), P(yt), P(rt), ctypes.c_uint64(mask), P(sums), None, None, None)
  torch.cuda.synchronize()
  print(rc, sums.cpu().numpy())
