"""In-tree build of the native code (no JIT cache: the .so files travel with the
repo snapshot to the GPU box).

  libhyperbo_b200.so   the C-ABI product (nvcc, sm_100a only)
  _C.*.so              thin pybind11 forwarding layer (g++), links the above
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhyperbo_b200.so")
EXT = os.path.join(HERE, "_C" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo",
    "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
]


def _newer(target, sources):
  if not os.path.exists(target):
    return True
  t = os.path.getmtime(target)
  return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
  proc = subprocess.run(cmd, capture_output=True, text=True)
  if proc.returncode != 0:
    raise RuntimeError("build failed: %s\n%s\n%s" % (" ".join(cmd), proc.stdout,
                                                      proc.stderr))


def build(force: bool = False, verbose: bool = False) -> None:
  nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
  cu_src = [os.path.join(CSRC, f) for f in (
      "hb_capi.cu", "hb_common.cuh", "hb_device.inc", "hb_kernels.inc",
      "hb_host.inc")]
  hdr = os.path.join(HERE, "..", "include", "hyperbo_b200.h")
  if force or _newer(LIB, cu_src + [hdr]):
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB, os.path.join(CSRC, "hb_capi.cu")]
    if verbose:
      print(" ".join(cmd), file=sys.stderr)
    _run(cmd)
  pyb = os.path.join(CSRC, "hb_pybind.cpp")
  if force or _newer(EXT, [pyb, hdr, LIB]):
    import pybind11
    cmd = [
        os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-fPIC", "-shared",
        "-fvisibility=hidden", "-I", pybind11.get_include(), "-I",
        sysconfig.get_paths()["include"], pyb, "-o", EXT, "-L", HERE,
        "-lhyperbo_b200", "-Wl,-rpath,$ORIGIN",
    ]
    if verbose:
      print(" ".join(cmd), file=sys.stderr)
    _run(cmd)


if __name__ == "__main__":
  build(force="--force" in sys.argv, verbose=True)
  print("built", LIB, EXT)
