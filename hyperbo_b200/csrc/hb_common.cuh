// hyperbo_b200 device building blocks (sm_100a).
//
// Data layout in HBM (DESIGN.md "Data layout"): every task's symmetric /
// triangular matrix (K~ -> L, and L^{-1}) is stored as PACKED LOWER-TRIANGULAR
// 64x64 TILES; tile (i,j), i >= j, lives at tile slot i(i+1)/2 + j and is one
// contiguous 64*64*sizeof(Real) block so that it moves with ONE 1-D TMA bulk
// copy (cp.async.bulk -> SASS UBLKCP).  Inside a tile, elements are grouped in
// 8x4 micro-blocks (8 rows x 4 cols, row-major inside, 32 scalars contiguous);
// micro-blocks are ordered [row_block(8)][col_block(16)].  One micro-block is
// exactly one warp-wide MMA operand fragment (lane = 4*(row&7) + (col&3)), so
// every fragment load is a single conflict-free, fully coalesced shared-memory
// request, for both the K-major and the MN-major operand role.  The same
// ordering is the tcgen05 SWIZZLE_NONE K-major canonical layout for 4-byte
// types (8 rows x 16 B core matrices), which is what the fp32 path needs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace hb {

constexpr int TB = 64;                 // tile edge
constexpr int TILE_ELEMS = TB * TB;    // 4096 scalars per tile
constexpr int NTHREADS = 256;          // 8 warps per CTA
constexpr int MAX_DIM = 32;
constexpr int HALF_ELEMS = TILE_ELEMS / 2;  // one k-half of a tile (64 x 32)
constexpr int STAGE_ELEMS = TILE_ELEMS;     // A half + B half
constexpr int NSTAGE = 3;                   // ring stages

struct TaskDesc {
  int n;               // true number of points
  int nblk;            // ceil(n / 64)
  long long xoff;      // first row of this task in X / y
  long long voff;      // offset into 64-padded per-task vectors (z, alpha)
  long long tile_off;  // first tile slot of this task in the packed buffers
  long long chol_off;  // element offset of this task's (n,n) row-major factor
  int theta_idx;       // which hyper-parameter set (second batch axis; 0 if one)
  int pad_;
};

__host__ __device__ __forceinline__ int tri_idx(int i, int j) {
  return i * (i + 1) / 2 + j;
}
// offset of element (r, c) inside a 64x64 tile
__host__ __device__ __forceinline__ int elem_off(int r, int c) {
  return ((((r >> 3) << 4) + (c >> 2)) << 5) + ((r & 7) << 2) + (c & 3);
}

#ifdef HB_FUSED_DEBUG
__device__ int g_dbg[8 * 1024];
#define HB_DBG_AT(n_) do { if (threadIdx.x == 0) g_dbg[blockIdx.x * 8 + 2] = (n_); } while (0)
#else
#define HB_DBG_AT(n_)
#endif
// ------------------------------------------------------------------ PTX ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(count)
               : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  const uint32_t a = smem_u32(bar);
#ifdef HB_FUSED_DEBUG
  unsigned long long spins_ = 0;
#endif
  while (!done) {
#ifdef HB_FUSED_DEBUG
    if (++spins_ > (1ull << 26)) {
      if ((threadIdx.x & 31) <= 1)
        printf("mbarrier timeout: block %d thread %d bar %u parity %u | item %d desc %x t0-at %d\n",
               blockIdx.x, threadIdx.x, a, parity, g_dbg[blockIdx.x * 8], g_dbg[blockIdx.x * 8 + 1],
               g_dbg[blockIdx.x * 8 + 2]);
      break;
    }
#endif
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
  }
}
// 1-D TMA bulk copy global -> shared, completion on an mbarrier (UBLKCP.S.G)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem,
                                         uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// 1-D TMA bulk copy shared -> global (bulk async-group completion)
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem,
                                         uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::
                   "l"(dst_gmem),
               "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read_all() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_all() {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
// order generic-proxy smem accesses before later async-proxy (TMA) accesses
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ------------------------------------------------------- fragment loads ----
enum Major { KMAJOR = 0, MNMAJOR = 1 };

// Warp roles inside the 256-thread CTA for one 64x64 output tile:
//   wk = warp >> 2 : split-K half of every k range
//   q              : 32x32 quadrant, wm = q >> 1 (rows), wn = q & 1 (cols)
// Warps w and w+4 share an SM sub-partition (w & 3).  The wk = 1 warp of a
// sub-partition takes quadrant 3 - (w & 3), so each sub-partition owns one
// warp of quadrant q and one of 3 - q: when triangular operands let the
// (wm = 1) or (wn = 0) warps skip half of a k range, the remaining tensor work
// stays balanced over the four sub-partitions.
// acc[fm][fn][e]: rows 32wm + 8fm + g, cols 32wn + 8fn + 2t + e.
struct FullK {};  // tag: the full-K warp mapping below
struct WarpPos {
  int warp, lane, wk, q, wm, wn, g, t;
  bool fullk;
  __device__ __forceinline__ WarpPos() {
    warp = threadIdx.x >> 5;
    lane = threadIdx.x & 31;
    wk = warp >> 2;
    q = wk ? 3 - (warp & 3) : (warp & 3);
    wm = q >> 1;
    wn = q & 1;
    g = lane >> 2;
    t = lane & 3;
    fullk = false;
  }
  // Full-K mapping (no split-K exchange, half the accumulator registers): warp
  // = one 16-row group r4 = 2 wm + wk of one 32-column half wn, over the whole
  // k range; acc[fi][fn][e] IS own[fi][fn][e] (rows 16 r4 + 8 fi + g, the same
  // own_row / own_col formulas).  The two warps of a sub-partition (w, w + 4)
  // take row groups (r4, 3 - r4) and opposite column halves, so triangular
  // clipping (now at 16-row granularity) stays balanced over the sub-partitions.
  __device__ __forceinline__ explicit WarpPos(FullK) {
    warp = threadIdx.x >> 5;
    lane = threadIdx.x & 31;
    const int s = warp & 3, hi = warp >> 2;
    int r4 = s >> 1;          // sub-partitions 0,1 -> row group 0; 2,3 -> 1
    wn = s & 1;
    if (hi) { r4 = 3 - r4; wn = 1 - wn; }
    wm = r4 >> 1;
    wk = r4 & 1;
    q = 2 * wm + wn;
    g = lane >> 2;
    t = lane & 3;
    fullk = true;
  }
};

// structural-zero flags of a tile product (per-warp k-range clipping)
enum TriFlags {
  TRI_NONE = 0,
  TRI_A_KLE = 1,    // A[m][k] = 0 for k > m  (lower-triangular, K-major role)
  TRI_A_KGE = 2,    // A[m][k] = 0 for k < m  (lower-triangular, MN-major role)
  TRI_B_KLE = 4,    // B[k][n] = 0 for k > n
  TRI_B_KGE = 8,    // B[k][n] = 0 for k < n
  TRI_SYM_LOWER = 16  // symmetric output, strictly-upper quadrant not needed
};
// clip the k-block range [lo, hi) (units of 4, tile-global) for this warp
__device__ __forceinline__ void tri_clip(int flags, const WarpPos& w, int& lo,
                                         int& hi) {
  if (w.fullk) {  // 16-row groups
    const int r4 = 2 * w.wm + w.wk;
    if (flags & TRI_A_KLE) hi = min(hi, 4 * (r4 + 1));
    if (flags & TRI_A_KGE) lo = max(lo, 4 * r4);
  } else {
    if (flags & TRI_A_KLE) hi = min(hi, 8 * (w.wm + 1));
    if (flags & TRI_A_KGE) lo = max(lo, 8 * w.wm);
  }
  if (flags & TRI_B_KLE) hi = min(hi, 8 * (w.wn + 1));
  if (flags & TRI_B_KGE) lo = max(lo, 8 * w.wn);
  if ((flags & TRI_SYM_LOWER) && w.wm == 0 && w.wn == 1) hi = lo;
}

// symmetric output (TRI_SYM_LOWER), full-K mapping: how many of the warp's four
// 8-column blocks hold entries on or below the diagonal of its 16 rows
__device__ __forceinline__ int sym_nfn(int flags, const WarpPos& w) {
  if (!(flags & TRI_SYM_LOWER) || !w.fullk) return 4;
  const int r4 = 2 * w.wm + w.wk;
  return max(0, min(4, 2 * (r4 + 1) - 4 * w.wn));
}

// offset of the MN-major fragment (k-block kl, idx-block blk8) inside a buffer
// whose rows are the contraction index and whose row-blocks hold 16 col-blocks
__device__ __forceinline__ int mn_off(int blk8, int kl, int g, int t) {
  return ((((kl >> 1) << 4) + (blk8 << 1) + (g >> 2)) << 5) +
         ((((kl & 1) << 2) + t) << 2) + (g & 3);
}


// theta buffer layout (doubles for every engine precision), written by k_prep
constexpr int TH_CONST = 0, TH_SV = 1, TH_NV = 2, TH_LS = 3;
constexpr int TH_INVLS = TH_LS + MAX_DIM;            // 35
constexpr int TH_CHAIN = TH_INVLS + MAX_DIM;         // 67 (3 + MAX_DIM entries)
constexpr int TH_DIAG = 110;  // noise_variance + jitter (what is added to diag K)
constexpr int TH_SIZE = 128;
constexpr double JITTER = 1e-6;    // basics/linalg.py:42
constexpr double EPS_WARP = 1e-10; // gp_utils/utils.py:28,73
constexpr int GP_STRIDE = 2 + MAX_DIM;  // [sv, nv, ls...]
constexpr int HB_XCHG_MAX = 64;         // scalars per peer-memory all-reduce
constexpr int HB_XBUF_BYTES = 128 + 2 * HB_XCHG_MAX * 8;
constexpr int LPT_GROUP_MAX = 256;      // tasks per launch-order group (<= T)
__host__ __device__ inline int xstride(int d) { return d | 1; }

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// cvt.rna.tf32: round an fp32 value to TF32 (10-bit mantissa), kept in a b32
__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

// counter-based generator of the device-side sub-sampling (hb_subsample)
__host__ __device__ __forceinline__ unsigned long long hb_mix64(unsigned long long z) {
  z += 0x9e3779b97f4a7c15ULL;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ unsigned hb_feistel_perm(unsigned i, unsigned n,
                                                             unsigned long long key) {
  unsigned hb = 1;  // half width in bits: 2^(2 hb) >= n
  while ((1ull << (2 * hb)) < n) ++hb;
  const unsigned hmask = (1u << hb) - 1u;
  unsigned x = i;
  do {
    unsigned l = x >> hb, r = x & hmask;
#pragma unroll
    for (int rnd = 0; rnd < 6; ++rnd) {
      const unsigned f = (unsigned)hb_mix64(key + ((unsigned long long)rnd << 56) + r) & hmask;
      const unsigned nl = r;
      r = l ^ f;
      l = nl;
    }
    x = (l << hb) | r;
  } while (x >= n);  // cycle walking keeps the map a bijection of [0, n)
  return x;
}


}  // namespace hb
