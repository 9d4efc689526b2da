"""In-tree build of the native code (no JIT cache: the .so files travel with the
repo snapshot to the GPU box).

  libhyperbo_b200.so   the C-ABI product (nvcc, sm_100a only; three
                       translation units compiled in parallel)
  _C.*.so              thin pybind11 forwarding layer (g++), links the above
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhyperbo_b200.so")
EXT = os.path.join(HERE, "_C" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo",
    "-std=c++17", "-Xcompiler", "-fPIC", "-Xfatbin", "-compress-all",
]
# translation units of the C-ABI library: the kernels are compiled once per
# engine precision, in parallel
UNITS = ("hb_capi.cu", "hb_f64.cu", "hb_f32.cu")


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
  cu_src = [os.path.join(CSRC, f) for f in UNITS + (
      "hb_internal.cuh", "hb_common.cuh", "hb_device.inc", "hb_kernels.inc",
      "hb_host.inc", "hb_fused.inc", "hb_mrhs.inc")]
  hdr = os.path.join(HERE, "..", "include", "hyperbo_b200.h")
  if force or _newer(LIB, cu_src + [hdr]):
    # objects go to a scratch directory: only the linked .so stays in-tree
    obj_dir = tempfile.mkdtemp(prefix="hb_build_")
    procs, objs = [], []
    for unit in UNITS:
      obj = os.path.join(obj_dir, unit[:-3] + ".o")
      cmd = [nvcc] + NVCC_FLAGS + ["-c", "-o", obj, os.path.join(CSRC, unit)]
      if verbose:
        print(" ".join(cmd), file=sys.stderr)
      procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE,
                                          stderr=subprocess.PIPE, text=True)))
      objs.append(obj)
    for cmd, proc in procs:
      out, err = proc.communicate()
      if proc.returncode != 0:
        raise RuntimeError("build failed: %s\n%s\n%s" % (" ".join(cmd), out, err))
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared",
           "-cudart", "static", "-o", LIB] + objs
    if verbose:
      print(" ".join(cmd), file=sys.stderr)
    _run(cmd)
    shutil.rmtree(obj_dir, ignore_errors=True)
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
