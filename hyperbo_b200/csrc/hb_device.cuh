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

namespace hb {

constexpr int TB = 64;                 // tile edge
constexpr int TILE_ELEMS = TB * TB;    // 4096 scalars per tile
constexpr int NTHREADS = 256;          // 8 warps per CTA
constexpr int MAX_DIM = 32;

struct TaskDesc {
  int n;               // true number of points
  int nblk;            // ceil(n / 64)
  long long xoff;      // first row of this task in X / y
  long long voff;      // offset into 64-padded per-task vectors (z, alpha)
  long long tile_off;  // first tile slot of this task in the packed buffers
  long long chol_off;  // element offset of this task's (n,n) row-major factor
};

__host__ __device__ __forceinline__ int tri_idx(int i, int j) {
  return i * (i + 1) / 2 + j;
}
// offset of element (r, c) inside a 64x64 tile
__host__ __device__ __forceinline__ int elem_off(int r, int c) {
  return ((((r >> 3) << 4) + (c >> 2)) << 5) + ((r & 7) << 2) + (c & 3);
}

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
  while (!done) {
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

// fp64 tensor-core MMA (DMMA.8x8x4): C(8x8) += A(8x4,row) * B(4x8,col)
// lane = 4*g + t :  a = A[g][t],  b = B[t][g],  c = {C[g][2t], C[g][2t+1]}
__device__ __forceinline__ void mma_884(double (&c)[2], double a, double b) {
  asm volatile(
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, "
      "{%0,%1};"
      : "+d"(c[0]), "+d"(c[1])
      : "d"(a), "d"(b));
}

// ------------------------------------------------------- fragment loads ----
enum Major { KMAJOR = 0, MNMAJOR = 1 };

// Operand stored as tile[idx][k] (idx = output row/col, k = contraction):
// the 8(idx) x 4(k) fragment of idx-block `blk8`, k-block `kb` is micro-block
// (blk8, kb): 32 contiguous scalars, lane-ordered.
template <typename Real>
__device__ __forceinline__ Real frag_kmajor(const Real* tile, int blk8, int kb,
                                            int lane) {
  return tile[(((blk8 << 4) + kb) << 5) + lane];
}
// Operand stored as tile[k][idx]: element (k = 4kb + t, idx = 8 blk8 + g).
template <typename Real>
__device__ __forceinline__ Real frag_mnmajor(const Real* tile, int blk8, int kb,
                                             int lane) {
  const int g = lane >> 2, t = lane & 3;
  return tile[((((kb >> 1) << 4) + (blk8 << 1) + (g >> 2)) << 5) +
              ((((kb & 1) << 2) + t) << 2) + (g & 3)];
}

// Warp roles inside the 256-thread CTA for one 64x64 output tile:
//   wk = warp >> 2 : split-K half of every k range
//   q              : 32x32 quadrant, wm = q >> 1 (rows), wn = q & 1 (cols)
// Warps w and w+4 share an SM sub-partition (w & 3).  The wk = 1 warp of a
// sub-partition takes quadrant 3 - (w & 3), so each sub-partition owns one
// warp of quadrant q and one of 3 - q: when triangular operands let the
// (wm = 1) or (wn = 0) warps skip half of a k range, the remaining tensor work
// stays balanced over the four sub-partitions.
// acc[fm][fn][e]: rows 32wm + 8fm + g, cols 32wn + 8fn + 2t + e.
struct WarpPos {
  int warp, lane, wk, q, wm, wn, g, t;
  __device__ __forceinline__ WarpPos() {
    warp = threadIdx.x >> 5;
    lane = threadIdx.x & 31;
    wk = warp >> 2;
    q = wk ? 3 - (warp & 3) : (warp & 3);
    wm = q >> 1;
    wn = q & 1;
    g = lane >> 2;
    t = lane & 3;
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
  if (flags & TRI_A_KLE) hi = min(hi, 8 * (w.wm + 1));
  if (flags & TRI_A_KGE) lo = max(lo, 8 * w.wm);
  if (flags & TRI_B_KLE) hi = min(hi, 8 * (w.wn + 1));
  if (flags & TRI_B_KGE) lo = max(lo, 8 * w.wn);
  if ((flags & TRI_SYM_LOWER) && w.wm == 0 && w.wn == 1) hi = lo;
}

// offset of the MN-major fragment (k-block kl, idx-block blk8) inside a buffer
// whose rows are the contraction index and whose row-blocks hold 16 col-blocks
__device__ __forceinline__ int mn_off(int blk8, int kl, int g, int t) {
  return ((((kl >> 1) << 4) + (blk8 << 1) + (g >> 2)) << 5) +
         ((((kl & 1) << 2) + t) << 2) + (g & 3);
}

// acc += A_half * B_half over the local k-blocks [kl_lo, kl_hi) of one k-half
// (8 k-blocks of 4).  K-major operand: micro-block (blk8, kl) at
// (blk8 * rbs + kl) * 32 (rbs = 8 in a pipeline stage, 16 in a resident tile
// whose pointer is pre-offset by 8h micro-blocks).  MN-major operand: rows
// 0..31 of the half, 16 col-blocks per row-block (pointer pre-offset by
// h * 2048 in a resident tile).  The two split-K warp groups share the range.
template <int AM, int BM>
__device__ __forceinline__ void half_mma(double (&acc)[4][4][2],
                                         const double* __restrict__ Ah, int a_rbs,
                                         const double* __restrict__ Bh, int b_rbs,
                                         const WarpPos& w, int kl_lo = 0,
                                         int kl_hi = 8) {
  const int kmid = kl_lo + ((kl_hi - kl_lo + 1) >> 1);
  const int k0 = w.wk ? kmid : kl_lo, k1 = w.wk ? kl_hi : kmid;
#pragma unroll 2
  for (int kl = k0; kl < k1; ++kl) {
    double a[4], b[4];
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      a[f] = (AM == KMAJOR) ? Ah[((((4 * w.wm + f) * a_rbs) + kl) << 5) + w.lane]
                            : Ah[mn_off(4 * w.wm + f, kl, w.g, w.t)];
      b[f] = (BM == KMAJOR) ? Bh[((((4 * w.wn + f) * b_rbs) + kl) << 5) + w.lane]
                            : Bh[mn_off(4 * w.wn + f, kl, w.g, w.t)];
    }
#pragma unroll
    for (int fm = 0; fm < 4; ++fm)
#pragma unroll
      for (int fn = 0; fn < 4; ++fn) mma_884(acc[fm][fn], a[fm], b[fn]);
  }
}

// both operands are full tiles resident in shared memory (tile layout)
template <int AM, int BM>
__device__ __forceinline__ void resident_mma(double (&acc)[4][4][2],
                                             const double* At, const double* Bt,
                                             const WarpPos& w,
                                             int flags = TRI_NONE) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    int lo = 8 * h, hi = 8 * h + 8;
    tri_clip(flags, w, lo, hi);
    if (lo < hi)
      half_mma<AM, BM>(acc, (AM == KMAJOR) ? At + ((8 * h) << 5) : At + h * 2048,
                       16, (BM == KMAJOR) ? Bt + ((8 * h) << 5) : Bt + h * 2048,
                       16, w, lo - 8 * h, hi - 8 * h);
  }
}

__device__ __forceinline__ void acc_zero(double (&acc)[4][4][2]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
}

// After the k loop the two split-K halves hold partial sums of the same
// quadrant.  Exchange through shared memory so that warp (wk, wq) ends up
// owning the fully reduced rows 32wm + 16wk + 8fi + g (fi = 0,1) of its
// quadrant:  own[fi][fn][e].  `xc` is a TILE_ELEMS scratch buffer.
__device__ __forceinline__ void splitk_exchange(double (&acc)[4][4][2],
                                                double (&own)[2][4][2],
                                                double* xc, const WarpPos& w) {
  const int wq = w.q;
  // send the half this warp does NOT keep to slot (wq, dest wk = 1 - wk);
  // static register indices only (a runtime index would spill acc to local)
  {
    double2* dst = reinterpret_cast<double2*>(xc) +
                   (((wq * 2 + (1 - w.wk)) * 8) << 5) + w.lane;
#pragma unroll
    for (int fi = 0; fi < 2; ++fi)
#pragma unroll
      for (int fn = 0; fn < 4; ++fn) {
        const double2 v = w.wk ? make_double2(acc[fi][fn][0], acc[fi][fn][1])
                               : make_double2(acc[2 + fi][fn][0],
                                              acc[2 + fi][fn][1]);
        dst[(fi * 4 + fn) << 5] = v;
      }
  }
  __syncthreads();
  {
    const double2* src = reinterpret_cast<const double2*>(xc) +
                         (((wq * 2 + w.wk) * 8) << 5) + w.lane;
#pragma unroll
    for (int fi = 0; fi < 2; ++fi)
#pragma unroll
      for (int fn = 0; fn < 4; ++fn) {
        const double2 v = src[(fi * 4 + fn) << 5];
        own[fi][fn][0] = (w.wk ? acc[2 + fi][fn][0] : acc[fi][fn][0]) + v.x;
        own[fi][fn][1] = (w.wk ? acc[2 + fi][fn][1] : acc[fi][fn][1]) + v.y;
      }
  }
  fence_async_smem();
  __syncthreads();
}

// row / col (inside the 64x64 tile) of own[fi][fn][e]
__device__ __forceinline__ int own_row(const WarpPos& w, int fi) {
  return 32 * w.wm + 16 * w.wk + 8 * fi + w.g;
}
__device__ __forceinline__ int own_col(const WarpPos& w, int fn, int e) {
  return 32 * w.wn + 8 * fn + 2 * w.t + e;
}

// write own[][][] into a shared-memory tile buffer in tile layout
__device__ __forceinline__ void own_to_tile(const double (&own)[2][4][2],
                                            double* tile, const WarpPos& w) {
#pragma unroll
  for (int fi = 0; fi < 2; ++fi)
#pragma unroll
    for (int fn = 0; fn < 4; ++fn) {
      const int r = own_row(w, fi), c = own_col(w, fn, 0);
      *reinterpret_cast<double2*>(tile + elem_off(r, c)) =
          make_double2(own[fi][fn][0], own[fi][fn][1]);
    }
}

// ----------------------------------------------------- streamed tile GEMM ---
constexpr int HALF_ELEMS = TILE_ELEMS / 2;  // one k-half of a tile (64 x 32)
constexpr int STAGE_ELEMS = TILE_ELEMS;     // A half + B half
constexpr int NSTAGE = 3;                   // ring = 3 x 32 KiB (fp64)

struct TilePair {
  const double* a;  // global tile of operand A
  const double* b;  // global tile of operand B (== a: reuse the A half)
  int kb_lo, kb_hi; // non-zero k-block range of this product, within [0,16)
  int flags;        // TriFlags: per-warp clipping for triangular operands
};

struct Pipe {
  uint64_t* bars;    // [NSTAGE] stage barriers + [1] aux barrier
  double* ring;      // NSTAGE x STAGE_ELEMS; also reused as R0/R1/R2 scratch
  uint32_t parmask;  // bit s: parity the next wait on barrier s expects
};

__device__ __forceinline__ void pipe_wait(Pipe& p, int s) {
  mbar_wait(&p.bars[s], (p.parmask >> s) & 1u);
  p.parmask ^= (1u << s);
}
// one whole tile global -> resident smem buffer on the aux barrier (thread 0)
__device__ __forceinline__ void load_tile_async(Pipe& p, double* dst,
                                                const double* src) {
  mbar_expect_tx(&p.bars[NSTAGE], TILE_ELEMS * 8);
  bulk_g2s(dst, src, TILE_ELEMS * 8, &p.bars[NSTAGE]);
}

// acc += sum_k A_k * B_k with the tiles streamed from global memory (L2) in
// k-HALVES through a 3-stage TMA-bulk / mbarrier ring: two half-steps (= one
// full tile product) are always in flight behind the one feeding the tensor
// pipe.  A K-major half is 8 row-block segments of 2 KiB, an MN-major half is
// one contiguous 16 KiB segment; warp 0's lanes issue the copies.  The ring
// must be idle on entry (a __syncthreads() since its last generic use) and is
// idle again on return.  hook(k, h, Ahalf) runs while a half is resident.
// pre() runs right after the first three half-steps have been issued, i.e.
// while their TMA copies are in flight (used for the X-block / vector loads).
// tail(stage) is called ONCE, by all threads, as soon as the stream has no more
// half-steps to issue: `stage` is then free for an extra resident load that
// rides behind the stream (e.g. the W tile of the lauum epilogue).  Returns the
// stage handed to tail (or -1): it is NOT idle on return.
template <int AM, int BM, class Fn, class Hook, class Pre, class Tail>
__device__ __forceinline__ int stream_gemm(double (&acc)[4][4][2], int K, Fn fn,
                                           Pipe& p, Hook hook, const WarpPos& w,
                                           Pre pre, Tail tail) {
  int ei = 0, ni = 0, tail_stage = -1;
  auto issue_next = [&]() {
    while (ei < 2 * K) {
      const int k = ei >> 1, h = ei & 1;
      const TilePair tp = fn(k);
      ++ei;
      if (max(tp.kb_lo, 8 * h) >= min(tp.kb_hi, 8 * h + 8)) continue;
      const int s = ni % NSTAGE;
      ++ni;
      if (w.warp == 0) {
        double* sa = p.ring + s * STAGE_ELEMS;
        double* sb = sa + HALF_ELEMS;
        const bool lb = tp.b != tp.a;
        if (w.lane == 0)
          mbar_expect_tx(&p.bars[s], (lb ? 2u : 1u) * HALF_ELEMS * 8u);
        __syncwarp();
        if (AM == KMAJOR) {
          if (w.lane < 8)
            bulk_g2s(sa + ((w.lane * 8) << 5), tp.a + ((w.lane * 16 + 8 * h) << 5),
                     2048, &p.bars[s]);
        } else if (w.lane == 0) {
          bulk_g2s(sa, tp.a + h * HALF_ELEMS, HALF_ELEMS * 8, &p.bars[s]);
        }
        if (lb) {
          if (BM == KMAJOR) {
            if (w.lane >= 8 && w.lane < 16)
              bulk_g2s(sb + (((w.lane - 8) * 8) << 5),
                       tp.b + (((w.lane - 8) * 16 + 8 * h) << 5), 2048, &p.bars[s]);
          } else if (w.lane == 8) {
            bulk_g2s(sb, tp.b + h * HALF_ELEMS, HALF_ELEMS * 8, &p.bars[s]);
          }
        }
      }
      return;
    }
    if (tail_stage < 0) {
      tail_stage = ni % NSTAGE;
      tail(tail_stage);
    }
  };
  issue_next();
  issue_next();
  issue_next();
  pre();
  int nc = 0;
  for (int ec = 0; ec < 2 * K; ++ec) {
    const int k = ec >> 1, h = ec & 1;
    const TilePair tp = fn(k);
    const int lo = max(tp.kb_lo, 8 * h), hi = min(tp.kb_hi, 8 * h + 8);
    if (lo >= hi) continue;
    const int s = nc % NSTAGE;
    ++nc;
    pipe_wait(p, s);
    const double* As = p.ring + s * STAGE_ELEMS;
    const double* Bs = (tp.b == tp.a) ? As : As + HALF_ELEMS;
    {
      int wlo = lo, whi = hi;
      tri_clip(tp.flags, w, wlo, whi);
      if (wlo < whi) half_mma<AM, BM>(acc, As, 8, Bs, 8, w, wlo - 8 * h, whi - 8 * h);
    }
    hook(k, h, As);
    __syncthreads();
    issue_next();
  }
  return tail_stage;
}

struct NoHook {
  __device__ __forceinline__ void operator()(int, int, const double*) const {}
};
struct NoPre {
  __device__ __forceinline__ void operator()() const {}
};
struct NoTail {
  __device__ __forceinline__ void operator()(int) const {}
};
template <int AM, int BM, class Fn, class Hook, class Pre>
__device__ __forceinline__ void stream_gemm(double (&acc)[4][4][2], int K, Fn fn,
                                            Pipe& p, Hook hook, const WarpPos& w,
                                            Pre pre) {
  stream_gemm<AM, BM>(acc, K, fn, p, hook, w, pre, NoTail());
}
template <int AM, int BM, class Fn, class Hook>
__device__ __forceinline__ void stream_gemm(double (&acc)[4][4][2], int K, Fn fn,
                                            Pipe& p, Hook hook,
                                            const WarpPos& w) {
  stream_gemm<AM, BM>(acc, K, fn, p, hook, w, NoPre(), NoTail());
}

// ------------------------------------------------------------ reductions ---
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// deterministic block sum; `red` has >= 8 doubles; result valid in all threads
__device__ __forceinline__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < NTHREADS / 32; ++i) s += red[i];
  return s;
}

}  // namespace hb
