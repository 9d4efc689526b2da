// EXPERIMENTAL -- round-2 starting point, NOT part of the product build and NOT
// yet run on hardware (written when this round's GPU budget was spent; it only
// has been compiled for sm_100a).  See DESIGN.md section 9, item 1.
//
// One CTA computes D(64x64) = A * B' for two 64x64 fp32 tiles held in the
// engine's packed tile layout (8x4 micro-blocks, [row_block][col_block]) with
// ONE chain of tcgen05.mma kind::tf32 instructions (M = 64, N = 64, K = 8 each),
// accumulating in tensor memory, and reads the accumulator back with
// tcgen05.ld.16x256b.x4 -- which, for M = 64, hands warp w rows 16 (w % 4) ..
// +15 in the m16n8 accumulator register order, i.e. the own[fi][fn][e]
// ownership of the engine's full-K warp mapping.
//
//   role 0: both operands K-major  (A[m][k], B[n][k]: rows = M / N)
//   role 1: both operands MN-major (A[k][m], B[k][n]: rows = contraction), the
//           k_lauum_grad case
//
// Descriptor fields follow cute/arch/mma_sm100_desc.hpp (SmemDescriptor,
// InstrDescriptor) of the CUTLASS copy vendored in this image.
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr int TILE_ELEMS = 64 * 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// SWIZZLE_NONE shared-memory matrix descriptor (version 1 = Blackwell)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes,
                                              uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);             // start address  [0,14)
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;   // leading offset [16,30)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;   // stride offset  [32,46)
  d |= (uint64_t)1 << 46;                             // version        [46,48)
  return d;                                           // layout_type = 0 (no swizzle)
}

// instruction descriptor: D fp32, A / B tf32, M = 64, N = 64
__device__ __forceinline__ uint32_t make_idesc(int a_mn_major, int b_mn_major) {
  uint32_t i = 0;
  i |= 1u << 4;                      // c_format = F32
  i |= 2u << 7;                      // a_format = TF32
  i |= 2u << 10;                     // b_format = TF32
  i |= (uint32_t)a_mn_major << 15;   // a_major: 0 = K, 1 = MN
  i |= (uint32_t)b_mn_major << 16;   // b_major
  i |= (64u >> 3) << 17;             // n_dim
  i |= (64u >> 4) << 24;             // m_dim
  return i;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}

__global__ void __launch_bounds__(256) k_umma_tile(const float* __restrict__ A,
                                                   const float* __restrict__ B,
                                                   float* __restrict__ D, int role,
                                                   unsigned lbo_mn, unsigned sbo_mn,
                                                   unsigned kstep_mn) {
  __shared__ __align__(128) float sA[TILE_ELEMS];
  __shared__ __align__(128) float sB[TILE_ELEMS];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int e = threadIdx.x; e < TILE_ELEMS; e += blockDim.x) {
    sA[e] = A[e];
    sB[e] = B[e];
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {  // one warp allocates 64 TMEM columns (power of two >= 32)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::
                     "r"(smem_u32(&tmem_base_smem)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // generic-proxy writes of sA / sB must be visible to the async (tensor) proxy
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_d = tmem_base_smem;

  if (warp == 1 && lane == 0) {
    const uint32_t idesc = make_idesc(role, role);
    // K-major role : next core matrix along K = next col-block (128 B) = LBO;
    //                next 8 rows = next row-block (16 col-blocks * 128 B) = SBO;
    //                K advances by 8 scalars = 2 col-blocks = 256 B per instruction.
    // MN-major role: next 4 MN columns = next col-block (128 B) = SBO;
    //                next 8 K rows = next row-block (2048 B) = LBO;
    //                K advances by 8 rows = one row-block = 2048 B per instruction.
    const uint32_t lbo = role ? lbo_mn : 128u, sbo = role ? sbo_mn : 2048u;
    const uint32_t kstep = role ? kstep_mn : 256u;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint64_t da = make_desc(smem_u32(sA) + k * kstep, lbo, sbo);
      const uint64_t db = make_desc(smem_u32(sB) + k * kstep, lbo, sbo);
      umma_tf32(tmem_d, da, db, idesc, k > 0);
    }
    // arrives on the mbarrier once every MMA above has completed
    asm volatile(
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::
            "r"(smem_u32(&bar))
        : "memory");
  }
  {  // every thread waits for phase 0 of the barrier
    uint32_t done = 0;
    while (!done)
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(smem_u32(&bar))
          : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;");

  // warp w: rows 16 (w % 4) .. +15 live in lanes 32 (w % 4) .. +15; columns
  // 32 (w / 4) .. +31.  16x256b.x4: reg[4 fn + 2 fi + e] = D[8 fi + g][8 fn + 2 t + e].
  const int q = warp & 3, wn = warp >> 2, g = lane >> 2, t = lane & 3;
  uint32_t r[16];
  const uint32_t taddr = tmem_d + ((uint32_t)(32 * q) << 16) + (uint32_t)(32 * wn);
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]),
        "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
        "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int fn = 0; fn < 4; ++fn)
#pragma unroll
    for (int fi = 0; fi < 2; ++fi)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int row = 16 * q + 8 * fi + g, col = 32 * wn + 8 * fn + 2 * t + e;
        D[row * 64 + col] = __uint_as_float(r[4 * fn + 2 * fi + e]);
      }

  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem_d));
}

}  // namespace

// A_tile / B_tile: device pointers to 4096 floats in the packed tile layout;
// D: device pointer to 64 x 64 row-major floats.  role: 0 K-major, 1 MN-major.
extern "C" int hb_exp_umma_tf32_tile(const float* A_tile, const float* B_tile, float* D,
                                     int role, void* stream) {
  k_umma_tile<<<1, 256, 0, (cudaStream_t)stream>>>(A_tile, B_tile, D, role, 2048u, 128u, 2048u);
  return (int)cudaGetLastError();
}
// MN-major descriptor experiments: explicit LBO / SBO / per-instruction K step
extern "C" int hb_exp_umma_tf32_tile_mn(const float* A_tile, const float* B_tile, float* D,
                                        unsigned lbo, unsigned sbo, unsigned kstep,
                                        void* stream) {
  k_umma_tile<<<1, 256, 0, (cudaStream_t)stream>>>(A_tile, B_tile, D, 1, lbo, sbo, kstep);
  return (int)cudaGetLastError();
}
