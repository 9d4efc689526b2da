"""Debug: cycle counts of potrf64_blocked's phases on an otherwise idle GPU and
with every SM slot busy (HB_STAMPS build of the library, built beforehand as
hyperbo_b200/libhb_stamps.so:  nvcc ... -DHB_STAMPS)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.path.join(ROOT, "hyperbo_b200", "libhb_stamps.so"))
out = (ctypes.c_longlong * 8)()
names = ["total", "potrf16 x4", "trsm+inv16", "trailing upd", "zero+inverse"]
# variants of the 16x16 pivot block (hb_kernels.inc, potrf16_warp_v): 0 = shipped,
# 1 = Newton step folded into the scaling, 2 = both half-warps, 3 = smem broadcast
for f32 in (0, 1):
  for pv in (0, 1, 2, 3):
    for grid in (1, 444):
      assert lib.hb_debug_potrf64(f32, grid | (pv << 16), 20, out) == 0
      print("fp32" if f32 else "fp64", "variant", pv, "CTAs", grid,
            " ".join("%s=%d" % (n, out[i]) for i, n in enumerate(names)), "chk", out[5])
