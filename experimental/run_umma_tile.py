"""EXPERIMENTAL (round-2 starting point; never run yet -- see umma_tf32_tile.cu).

Builds umma_tf32_tile.cu for sm_100a and checks the tcgen05 TF32 tile product
against torch for both operand roles:

    python experimental/run_umma_tile.py            (on a B200; ~1 s)

Exit code 0 = both roles within TF32 tolerance.  Run it under a short
`timeout`: a wrong descriptor can hang the MMA-completion wait.
"""
import ctypes
import os
import subprocess
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libumma_tf32_tile.so")


def elem_off(r, c):  # hb_common.cuh: offset of (r, c) inside a packed 64x64 tile
  return ((((r >> 3) << 4) + (c >> 2)) << 5) + ((r & 7) << 2) + (c & 3)


def pack(mat):
  out = np.empty(4096, dtype=np.float32)
  r, c = np.meshgrid(np.arange(64), np.arange(64), indexing="ij")
  out[elem_off(r, c)] = mat
  return out


def main():
  subprocess.check_call([
      "nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17",
      "-Xcompiler", "-fPIC", "-shared", "-cudart", "static", "-o", SO,
      os.path.join(HERE, "umma_tf32_tile.cu")])
  lib = ctypes.CDLL(SO)
  rng = np.random.default_rng(0)
  ok = True
  for role in (0, 1):
    a = rng.standard_normal((64, 64)).astype(np.float32)
    b = rng.standard_normal((64, 64)).astype(np.float32)
    # role 0: tiles hold A[m][k], B[n][k] -> D = A B';  role 1: A[k][m], B[k][n] -> D = A' B
    want = a @ b.T if role == 0 else a.T @ b
    ta = torch.from_numpy(pack(a)).cuda()
    tb = torch.from_numpy(pack(b)).cuda()
    d = torch.zeros((64, 64), dtype=torch.float32, device="cuda")
    rc = lib.hb_exp_umma_tf32_tile(ctypes.c_void_p(ta.data_ptr()),
                                   ctypes.c_void_p(tb.data_ptr()),
                                   ctypes.c_void_p(d.data_ptr()), role, None)
    torch.cuda.synchronize()
    err = float(np.max(np.abs(d.cpu().numpy() - want)) / np.max(np.abs(want)))
    print(f"role {role}: launch rc={rc}  max rel err vs fp32 = {err:.3e} "
          f"(TF32 inputs: expect ~1e-3)")
    ok = ok and rc == 0 and err < 5e-3
  # MN-major descriptor variants (role 1): which (LBO, SBO, K step) describes the
  # packed tile read as [k][mn]?
  a = rng.standard_normal((64, 64)).astype(np.float32)
  b = rng.standard_normal((64, 64)).astype(np.float32)
  want = a.T @ b
  ta, tb = torch.from_numpy(pack(a)).cuda(), torch.from_numpy(pack(b)).cuda()
  for lbo, sbo, kstep in ((2048, 128, 2048), (128, 2048, 2048), (128, 1024, 2048),
                          (1024, 128, 2048), (2048, 128, 1024), (128, 2048, 1024),
                          (128, 128, 2048), (2048, 2048, 2048), (4096, 128, 2048),
                          (128, 4096, 2048), (256, 2048, 2048), (2048, 256, 2048)):
    d = torch.zeros((64, 64), dtype=torch.float32, device="cuda")
    rc = lib.hb_exp_umma_tf32_tile_mn(ctypes.c_void_p(ta.data_ptr()), ctypes.c_void_p(tb.data_ptr()),
                                      ctypes.c_void_p(d.data_ptr()), lbo, sbo, kstep, None)
    torch.cuda.synchronize()
    got = d.cpu().numpy()
    err = float(np.max(np.abs(got - want)) / np.max(np.abs(want)))
    # how many output entries are right: hints at which dimension is mis-strided
    good = float(np.mean(np.abs(got - want) < 2e-2 * np.max(np.abs(want))))
    print(f"MN variant lbo={lbo} sbo={sbo} kstep={kstep}: rc={rc} err={err:.3e} good={good:.2f}")
  return 0 if ok else 1


if __name__ == "__main__":
  sys.exit(main())
