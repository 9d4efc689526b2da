// hyperbo_b200 kernels: batched GP NLL + gradient on packed 64x64 tiles.
//
// Reference arithmetic restated (paths under /root/reference/hyperbo/):
//   kernel build        gp_utils/kernel.py:63-123, basics/linalg.py:36-69
//   Cholesky + solves   basics/linalg.py:29-33,72-110,139-171
//   NLL value           gp_utils/objectives.py:144-156,178-195
//   gradient            closed form of jax.value_and_grad at gp_utils/gp.py:134
//   Adam                optax.adam as used at gp_utils/gp.py:124,143-144
//
// Algorithm per task (all tasks of a call advance together, one launch per
// block column): left-looking blocked Cholesky on 64x64 tiles whose trailing
// tile products run on the fp64 tensor pipe (DMMA), with the kernel matrix
// evaluated on the fly (K~ never hits memory), the triangular inverse
// M = L^{-1} built row by row in the SAME launches, then K~^{-1} = M'M formed
// tile by tile and contracted against dK/dtheta in the epilogue (K~^{-1} never
// hits memory either).
#pragma once
#include "hb_device.cuh"

namespace hb {

// theta buffer layout (doubles), written by k_prep
constexpr int TH_CONST = 0, TH_SV = 1, TH_NV = 2, TH_LS = 3;
constexpr int TH_INVLS = TH_LS + MAX_DIM;            // 35
constexpr int TH_CHAIN = TH_INVLS + MAX_DIM;         // 67 (3 + MAX_DIM entries)
constexpr int TH_SIZE = 128;
constexpr double JITTER = 1e-6;    // basics/linalg.py:42
constexpr double EPS_WARP = 1e-10; // gp_utils/utils.py:28,73

struct Params {
  const TaskDesc* tasks;
  int T, d;
  int kernel_id, mean_id;
  int with_trtri;
  const double* X;
  const double* y;
  const double* theta;
  double* Lt;       // packed tiles of L
  double* Mt;       // packed tiles of M = L^{-1} (diag tiles: inv of diag blocks)
  double* Wt;       // packed tiles of the pair weights W (dK/dl = W D^2/l^3);
                    // nullptr when no gradient is requested
  double* zz;       // per task: z'z (= r' K~^{-1} r)
  double* z;        // L^{-1} r, 64-padded per task
  double* alpha;    // K~^{-1} r, 64-padded per task
  double* logdet;   // per (task, block): sum_k log L_kk of that diagonal block
  double* asum;     // per (task, block): sum of alpha over the block
  double* nll_task; // per task
  double* gpart;    // per tile slot: [2 + MAX_DIM] gradient partials
  double* gtask;    // per task: [2 + MAX_DIM]
  int* info;        // per task: 0 or failing column + 1
  unsigned* bad;    // per task scratch: min failing column + 1, ~0u if none
  long long* stamps;  // debug: per-CTA phase timestamps (HB_STAMPS builds only)
};

constexpr int GP_STRIDE = 2 + MAX_DIM;  // [sv, nv, ls...]
constexpr int LPT_GROUP_MAX = 256;      // tasks per launch-order group (<= T)

// shared-memory map (bytes) of the tile kernels.  Everything large lives in the
// 96 KiB ring: while no stream is in flight its three 32 KiB stages double as
// R0 (split-K exchange), R1 / R2 (resident operand / output tiles).  ~106 KiB
// per CTA at d = 8 => 2 CTAs per SM (one CTA's epilogue / diagonal block
// overlaps the other's tensor-pipe work).
constexpr int SM_BARS = 0;                              // 4 mbarriers
constexpr int SM_RED = 64;                              // 16 doubles
constexpr int SM_VEC = 192;                             // 2 x 64 doubles
constexpr int SM_RING = 1280;                           // 3 x 32 KiB
constexpr int SM_X = SM_RING + NSTAGE * STAGE_ELEMS * 8;
__host__ __device__ inline int xstride(int d) { return d | 1; }
__host__ inline size_t step_smem_bytes(int d) {
  return SM_X + 2 * 64 * xstride(d) * 8;
}
// k_lauum_grad additionally keeps 8 x GP_STRIDE reduction slots after the X blocks
__host__ inline size_t lauum_smem_bytes(int d) {
  return SM_X + 2 * 64 * xstride(d) * 8 + 8 * GP_STRIDE * 8;
}
constexpr size_t PREDICT_SMEM_BYTES = SM_X;

__device__ __forceinline__ Pipe make_pipe(unsigned char* smem) {
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BARS);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s <= NSTAGE; ++s) mbar_init(&bars[s], 1);
    fence_mbar_init();
  }
  return Pipe{bars, reinterpret_cast<double*>(smem + SM_RING), 0u};
}

// ------------------------------------------------------------------ prep ---
__device__ __forceinline__ double softplus(double x) {
  return x > 0.0 ? x + log1p(exp(-x)) : log1p(exp(x));
}
__device__ __forceinline__ double sigmoid(double x) {
  return 1.0 / (1.0 + exp(-x));
}

// raw -> theta (params_utils.retrieve_params + utils.DEFAULT_WARP_FUNC), the
// chain-rule factors d theta / d raw, and per-call state reset.
__global__ void k_prep(const double* __restrict__ raw, uint64_t warp_mask,
                       int d, int mean_id, double* __restrict__ theta,
                       unsigned* bad, int T) {
  const int p = threadIdx.x;
  if (p < 3 + d) {
    const double r = raw[p];
    const bool wp = (warp_mask >> p) & 1ull;
    double v = wp ? softplus(r) + EPS_WARP : r;
    double ch = wp ? sigmoid(r) : 1.0;
    if (p == 0 && mean_id == 0) { v = 0.0; ch = 0.0; }
    theta[TH_CHAIN + p] = ch;
    if (p < 3) theta[p] = v;
    else {
      theta[TH_LS + (p - 3)] = v;
      theta[TH_INVLS + (p - 3)] = 1.0 / v;
    }
  }
  for (int t = threadIdx.x; t < T; t += blockDim.x) bad[t] = 0xffffffffu;
}

// ------------------------------------------------------- kernel functions ---
// k and the pair weight W with dK/dl_k = W * Delta_k^2 / l_k^3
template <int KID>
__device__ __forceinline__ void kern_eval(double r2, double sv, double& k,
                                          double& wgt) {
  if (KID == 0) {  // squared_exponential, kernel.py:63-81
    k = sv * exp(-0.5 * r2);
    wgt = k;
  } else if (KID == 1) {  // matern32, kernel.py:84-102
    const double r = sqrt(3.0 * r2);
    const double e = exp(-r);
    k = sv * (1.0 + r) * e;
    wgt = 3.0 * sv * e;
  } else {  // matern52, kernel.py:105-123
    const double r = sqrt(5.0 * r2);
    const double e = exp(-r);
    k = sv * (1.0 + r + r * r * (1.0 / 3.0)) * e;
    wgt = (5.0 / 3.0) * sv * (1.0 + r) * e;
  }
}

// load the 64 x d block `blk` of a task's inputs, scaled by 1/lengthscale,
// rows beyond n zero-filled.  xs[r * DP + k].
__device__ __forceinline__ void load_xblock(double* xs, const double* X,
                                            long long row0, int nvalid, int d,
                                            int DP, const double* theta) {
  for (int e = threadIdx.x; e < 64 * d; e += NTHREADS) {
    const int r = e / d, k = e - r * d;
    xs[r * DP + k] =
        (r < nvalid) ? X[(row0 + r) * d + k] * theta[TH_INVLS + k] : 0.0;
  }
}

// own[fi][fn][e] <- K~ tile (bi, bj) evaluated from the scaled input blocks.
// Padding rows/cols (>= n) make K~ the identity there.  If `sub`, computes
// K~ - own instead.  wout (optional) receives the pair weights W.
template <int KID, bool SUB>
__device__ __forceinline__ void ktile_eval(double (&own)[2][4][2],
                                           const double* xi, const double* xj,
                                           int d, int DP, int row0, int col0,
                                           int n, double sv, double diag_add,
                                           const WarpPos& w,
                                           double* wtile = nullptr) {
  double r2[2][4][2];
#pragma unroll
  for (int fi = 0; fi < 2; ++fi)
#pragma unroll
    for (int fn = 0; fn < 4; ++fn) r2[fi][fn][0] = r2[fi][fn][1] = 0.0;
  for (int k = 0; k < d; ++k) {
    double xr[2], xc[4][2];
#pragma unroll
    for (int fi = 0; fi < 2; ++fi) xr[fi] = xi[own_row(w, fi) * DP + k];
#pragma unroll
    for (int fn = 0; fn < 4; ++fn)
#pragma unroll
      for (int e = 0; e < 2; ++e) xc[fn][e] = xj[own_col(w, fn, e) * DP + k];
#pragma unroll
    for (int fi = 0; fi < 2; ++fi)
#pragma unroll
      for (int fn = 0; fn < 4; ++fn)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const double df = xr[fi] - xc[fn][e];
          r2[fi][fn][e] = fma(df, df, r2[fi][fn][e]);
        }
  }
#pragma unroll
  for (int fi = 0; fi < 2; ++fi)
#pragma unroll
    for (int fn = 0; fn < 4; ++fn) {
      double wv[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int gr = row0 + own_row(w, fi), gc = col0 + own_col(w, fn, e);
        double k, wg;
        kern_eval<KID>(r2[fi][fn][e], sv, k, wg);
        if (gr == gc) k += diag_add;
        if (gr >= n || gc >= n) {
          k = (gr == gc) ? 1.0 : 0.0;
          wg = 0.0;  // padding never contributes to the gradient
        }
        wv[e] = wg;
        own[fi][fn][e] = SUB ? k - own[fi][fn][e] : k;
      }
      // keep the pair weights for the gradient contraction (k_lauum_grad)
      if (wtile)
        *reinterpret_cast<double2*>(
            wtile + elem_off(own_row(w, fi), own_col(w, fn, 0))) =
            make_double2(wv[0], wv[1]);
    }
}

// ------------------------------------------- in-CTA diagonal block (64x64) ---
// Blocked right-looking Cholesky of the SPD block held in TILE LAYOUT in `At`
// (becomes L, strict upper zeroed) plus its inverse into `Mt` (zeroed on
// entry), b = 16:
//   (1) warp 0 factors the 16x16 diagonal sub-block in registers (one row per
//       lane, column broadcast by shuffles; every lane tracks the running
//       diagonal so the pivot needs no extra broadcast),
//   (2) the rows below solve x L_D' = a one row per thread, while warp 7
//       inverts L_D (needed only for the final inverse, off the critical path),
//   (3) the trailing 16x16 blocks are updated on the tensor pipe (DMMA), one
//       warp per block, operands read straight from the tile layout.
// Then L^{-1} is assembled block row by block row with DMMA products.
__device__ __forceinline__ void potrf16_warp(double* At, int b, double* rsv,
                                             int lane) {
  // (An LDL'-style variant -- reciprocal instead of rsqrt in the pivot chain,
  // scaling deferred -- measured SLOWER, 5.6K vs 3.9K cycles per block: a
  // single warp is bound by its instruction count here, not by the chain.)
  const int r = lane & 15;  // lanes 16..31 mirror lanes 0..15 (no writes)
  double a[16], dd[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) a[c] = At[elem_off(16 * b + r, 16 * b + c)];
#pragma unroll
  for (int c = 0; c < 16; ++c) dd[c] = __shfl_sync(0xffffffffu, a[c], c);
  double rsk[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const double rs = rsqrt(dd[k]);  // NaN for a negative pivot: propagates
    rsk[k] = rs;
    const double l = a[k] * rs;  // L[r][k]
    a[k] = l;
#pragma unroll
    for (int c = k + 1; c < 16; ++c) {
      const double lc = __shfl_sync(0xffffffffu, l, c);  // L[c][k]
      a[c] = fma(-l, lc, a[c]);
      dd[c] = fma(-lc, lc, dd[c]);
    }
  }
  __syncwarp();  // the mirrored lanes 16..31 have read these rows too
  if (lane < 16) {
#pragma unroll
    for (int c = 0; c < 16; ++c)
      At[elem_off(16 * b + r, 16 * b + c)] = (c <= r) ? a[c] : 0.0;
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 16; ++k) rsv[16 * b + k] = rsk[k];
  }
}

// one row below the diagonal sub-block: x L_D' = a  (forward substitution)
__device__ __forceinline__ void trsm16_row(double* At, int b, int row,
                                           const double* rsv) {
  double a[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) a[c] = At[elem_off(row, 16 * b + c)];
#pragma unroll
  for (int m = 0; m < 16; ++m) {
    const double x = a[m] * rsv[16 * b + m];
    a[m] = x;
#pragma unroll
    for (int c = m + 1; c < 16; ++c)
      a[c] = fma(-x, At[elem_off(16 * b + c, 16 * b + m)], a[c]);  // broadcast
  }
#pragma unroll
  for (int c = 0; c < 16; ++c) At[elem_off(row, 16 * b + c)] = a[c];
}

// W = L_D^{-1}: lane c owns column c
__device__ __forceinline__ void trtri16_warp(const double* At, double* Mt, int b,
                                             const double* rsv, int lane) {
  if (lane >= 16) return;
  const int c = lane;
  double wv[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) wv[r] = (r == c) ? 1.0 : 0.0;
#pragma unroll
  for (int m = 0; m < 16; ++m) {
    const double x = wv[m] * rsv[16 * b + m];
    wv[m] = x;
#pragma unroll
    for (int r = m + 1; r < 16; ++r)
      wv[r] = fma(-x, At[elem_off(16 * b + r, 16 * b + m)], wv[r]);
  }
#pragma unroll
  for (int r = 0; r < 16; ++r) Mt[elem_off(16 * b + r, 16 * b + c)] = wv[r];
}

// 16x16 block product on the tensor pipe, one warp: c[fm][fn] (rows 8fm+g,
// cols 8fn+2t+e of the block) += A(16 x 16) * B(16 x 16), A K-major at
// micro-blocks (a_blk8 + fm, a_kb + kq); B K-major at (b_blk8 + fn, b_kb + kq)
// or MN-major (rows = contraction) at k-block b_kb + kq, col-block b_blk8 + fn.
template <int BM>
__device__ __forceinline__ void blk16_mma(double (&c)[2][2][2], const double* At,
                                          int a_blk8, int a_kb, const double* Bt,
                                          int b_blk8, int b_kb, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int kq = 0; kq < 4; ++kq) {
    double a[2], bb[2];
#pragma unroll
    for (int f = 0; f < 2; ++f) {
      a[f] = At[((((a_blk8 + f) << 4) + a_kb + kq) << 5) + lane];
      bb[f] = (BM == KMAJOR) ? Bt[((((b_blk8 + f) << 4) + b_kb + kq) << 5) + lane]
                             : Bt[mn_off(b_blk8 + f, b_kb + kq, g, t)];
    }
#pragma unroll
    for (int fm = 0; fm < 2; ++fm)
#pragma unroll
      for (int fn = 0; fn < 2; ++fn) mma_884(c[fm][fn], a[fm], bb[fn]);
  }
}

__device__ __forceinline__ void potrf64_blocked(double* At, double* Mt,
                                                double* rsv, const WarpPos& w,
                                                long long* dbg = nullptr) {
#ifdef HB_STAMPS
  long long t0_ = clock64(), tp_ = 0, tt_ = 0, ts_ = 0, t1_;
#define HB_PT(acc_) t1_ = clock64(); acc_ += t1_ - t0_; t0_ = t1_
#else
#define HB_PT(acc_)
#endif
  for (int b = 0; b < 4; ++b) {
    if (w.warp == 0) potrf16_warp(At, b, rsv, w.lane);
    __syncthreads();
    HB_PT(tp_);
    const int nbelow = 48 - 16 * b;
    if ((int)threadIdx.x < nbelow)
      trsm16_row(At, b, 16 * (b + 1) + threadIdx.x, rsv);
    else if (w.warp == 7)
      trtri16_warp(At, Mt, b, rsv, w.lane);
    __syncthreads();
    HB_PT(tt_);
    // trailing blocks (r16, c16), b < c16 <= r16 <= 3, one warp each
    const int nb = 3 - b;
    if (w.warp < nb * (nb + 1) / 2) {
      int rr = 0, cc = w.warp;
      while (cc > rr) { cc -= rr + 1; ++rr; }
      const int r16 = b + 1 + rr, c16 = b + 1 + cc;
      double c[2][2][2] = {};
      blk16_mma<KMAJOR>(c, At, 2 * r16, 4 * b, At, 2 * c16, 4 * b, w.lane);
#pragma unroll
      for (int fm = 0; fm < 2; ++fm)
#pragma unroll
        for (int fn = 0; fn < 2; ++fn) {
          double2* ptr = reinterpret_cast<double2*>(
              At + elem_off(16 * r16 + 8 * fm + w.g, 16 * c16 + 8 * fn + 2 * w.t));
          double2 v = *ptr;
          v.x -= c[fm][fn][0];
          v.y -= c[fm][fn][1];
          *ptr = v;
        }
    }
    __syncthreads();
    HB_PT(ts_);
  }
#ifdef HB_STAMPS
  if (dbg && threadIdx.x == 0) { dbg[0] = tp_; dbg[1] = tt_; dbg[2] = ts_; dbg[3] = clock64(); }
#endif
  // zero the strictly upper 16x16 blocks of L (stale K~ values)
  for (int e = threadIdx.x; e < TILE_ELEMS; e += NTHREADS) {
    const int r = e >> 6, c = e & 63;
    if ((c >> 4) > (r >> 4)) At[elem_off(r, c)] = 0.0;
  }
  __syncthreads();
  // block rows of M = L^{-1}:  M(i,j) = -W_i * sum_{k=j}^{i-1} L(i,k) M(k,j)
  for (int i = 1; i < 4; ++i) {
    if (w.warp < i) {
      const int j = w.warp;
      double tacc[2][2][2] = {};
      for (int k = j; k < i; ++k)
        blk16_mma<MNMAJOR>(tacc, At, 2 * i, 4 * k, Mt, 2 * j, 4 * k, w.lane);
      // stage T in M(i,j) (tile layout) to feed it back as an MN-major operand
#pragma unroll
      for (int fm = 0; fm < 2; ++fm)
#pragma unroll
        for (int fn = 0; fn < 2; ++fn)
          *reinterpret_cast<double2*>(
              Mt + elem_off(16 * i + 8 * fm + w.g, 16 * j + 8 * fn + 2 * w.t)) =
              make_double2(tacc[fm][fn][0], tacc[fm][fn][1]);
      __syncwarp();
      double racc[2][2][2] = {};
      blk16_mma<MNMAJOR>(racc, Mt, 2 * i, 4 * i, Mt, 2 * j, 4 * i, w.lane);
      __syncwarp();
#pragma unroll
      for (int fm = 0; fm < 2; ++fm)
#pragma unroll
        for (int fn = 0; fn < 2; ++fn)
          *reinterpret_cast<double2*>(
              Mt + elem_off(16 * i + 8 * fm + w.g, 16 * j + 8 * fn + 2 * w.t)) =
              make_double2(-racc[fm][fn][0], -racc[fm][fn][1]);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------- step kernel --
// One launch per block column j (j = -1 .. nblk_max-1).  State on entry:
// L(:, <j), L(j,j), M(j,j) = L(j,j)^{-1}, z_{<=j} and rows < j of M are final.
// Roles (blockIdx.x) of task blockIdx.y:
//   role 0 .. np-1 : panel tile i = j+1+role:  L(i,j) = (K(i,j) - sum_{k<j}
//                    L(i,k) L(j,k)') M(j,j)'            [tensor-pipe GEMMs]
//        role 0 then also factors the NEXT diagonal block (look-ahead):
//                    A = K~(i,i) - sum_{k<=j} L(i,k) L(i,k)',  L(i,i) = chol(A),
//                    M(i,i) = L(i,i)^{-1},  z_i = M(i,i) (r_i - sum L(i,k) z_k)
//   then j roles   : row j of M:  M(j,c) = -M(j,j) sum_{k=c}^{j-1} L(j,k) M(k,c)
template <int KID>
__global__ void __launch_bounds__(NTHREADS, 2) k_step(Params P, int j,
                                                       int nroles,
                                                       int LPT_GROUP) {
  extern __shared__ __align__(128) unsigned char smem[];
  // 1-D grid.  Tasks are taken in groups of LPT_GROUP = min(T, 256) (one group
  // for the usual batch sizes: measured best, 2.11 vs 2.18 / 2.33 ms per step
  // for groups of 64 / 16 at 256 x 512 x 8); inside a group the
  // order is role-major with roles sorted by decreasing length (look-ahead
  // CTA, panels, then the trtri roles from the longest k-loop to the
  // shortest): long CTAs start first, so the launch tail is one short CTA,
  // while the CTAs in flight belong to ~2 groups whose tiles stay L2-resident.
  const int gsz = LPT_GROUP * nroles;
  const int grp = blockIdx.x / gsz;
  const int within = blockIdx.x - grp * gsz;
  const int role = within / LPT_GROUP;
  const int task = grp * LPT_GROUP + (within - role * LPT_GROUP);
  if (task >= P.T) return;
  const TaskDesc td = P.tasks[task];
  const int nblk = td.nblk;
  if (nblk == 0) return;
  const int np = (j < 0) ? 1 : max(0, nblk - 1 - j);
  const int nt = (P.with_trtri && j >= 1 && j < nblk) ? j : 0;
  if (role >= np + nt) return;

  const WarpPos w;
#ifdef HB_STAMPS
  long long sk_[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  sk_[0] = clock64();
#define HB_SK(k) sk_[k] = clock64()
#define HB_SK_OUT(kind)                                                        \
  if (threadIdx.x == 0 && P.stamps) {                                          \
    long long* o = P.stamps + (((size_t)(j + 1) * P.T + task) * 8 +      \
                               role) * 8;                                \
    for (int q_ = 0; q_ < 7; ++q_) o[q_] = sk_[q_];                            \
    o[7] = kind;                                                               \
  }
#else
#define HB_SK(k)
#define HB_SK_OUT(kind)
#endif
  double* red = reinterpret_cast<double*>(smem + SM_RED);
  double* vec = reinterpret_cast<double*>(smem + SM_VEC);
  const int DP = xstride(P.d);
  double* xi = reinterpret_cast<double*>(smem + SM_X);
  double* xj = xi + 64 * DP;
  Pipe pipe = make_pipe(smem);
  double* R0 = pipe.ring;
  double* R1 = pipe.ring + STAGE_ELEMS;
  double* R2 = pipe.ring + 2 * STAGE_ELEMS;
  __syncthreads();
  constexpr uint32_t TILE_BYTES = TILE_ELEMS * sizeof(double);

  const double* Lt = P.Lt + td.tile_off * TILE_ELEMS;
  double* Ltw = P.Lt + td.tile_off * TILE_ELEMS;
  const double* Mt = P.Mt + td.tile_off * TILE_ELEMS;
  double* Mtw = P.Mt + td.tile_off * TILE_ELEMS;
  const double sv = P.theta[TH_SV];
  const double nv = P.theta[TH_NV];
  double acc[4][4][2];
  double own[2][4][2];

  // ------------------------------------------------------------ trtri role
  if (role >= np) {
    const int c = role - np;
    acc_zero(acc);
    stream_gemm<KMAJOR, MNMAJOR>(
        acc, j - c,
        [&](int kk) {
          const int k = c + kk;
          return TilePair{Lt + (size_t)tri_idx(j, k) * TILE_ELEMS,
                          Mt + (size_t)tri_idx(k, c) * TILE_ELEMS, 0, 16,
                          kk == 0 ? TRI_B_KGE : TRI_NONE};
        },
        pipe, NoHook(), w);
    HB_SK(1);
    if (threadIdx.x == 0)
      load_tile_async(pipe, R2, Mt + (size_t)tri_idx(j, j) * TILE_ELEMS);
    splitk_exchange(acc, own, R0, w);
    own_to_tile(own, R1, w);  // S as [k'][n]: the MN-major B operand
    __syncthreads();
    pipe_wait(pipe, NSTAGE);
    acc_zero(acc);
    resident_mma<KMAJOR, MNMAJOR>(acc, R2, R1, w, TRI_A_KLE);
    splitk_exchange(acc, own, R0, w);
#pragma unroll
    for (int fi = 0; fi < 2; ++fi)
#pragma unroll
      for (int fn = 0; fn < 4; ++fn) {
        own[fi][fn][0] = -own[fi][fn][0];
        own[fi][fn][1] = -own[fi][fn][1];
      }
    own_to_tile(own, R1, w);
    fence_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
      bulk_s2g(Mtw + (size_t)tri_idx(j, c) * TILE_ELEMS, R1, TILE_BYTES);
      bulk_commit();
      bulk_wait_read_all();
    }
    HB_SK(2);
    HB_SK_OUT(2);
    return;
  }

  // ------------------------------------------------------------ panel role
  const int i = (j < 0) ? 0 : j + 1 + role;
  auto load_x = [&]() {
    load_xblock(xi, P.X, td.xoff + 64LL * i, min(64, td.n - 64 * i), P.d, DP,
                P.theta);
    if (j >= 0)
      load_xblock(xj, P.X, td.xoff + 64LL * j, min(64, td.n - 64 * j), P.d, DP,
                  P.theta);
  };
  if (j < 0) load_x();
  if (j >= 0) {
    acc_zero(acc);
    stream_gemm<KMAJOR, KMAJOR>(
        acc, j,
        [&](int k) {
          return TilePair{Lt + (size_t)tri_idx(i, k) * TILE_ELEMS,
                          Lt + (size_t)tri_idx(j, k) * TILE_ELEMS, 0, 16,
                          TRI_NONE};
        },
        pipe, NoHook(), w, load_x);
    HB_SK(1);
    if (threadIdx.x == 0)
      load_tile_async(pipe, R2, Mt + (size_t)tri_idx(j, j) * TILE_ELEMS);
    splitk_exchange(acc, own, R0, w);  // its barriers also publish xi / xj
    ktile_eval<KID, true>(
        own, xi, xj, P.d, DP, 64 * i, 64 * j, td.n, sv, 0.0, w,
        P.Wt ? P.Wt + (td.tile_off + tri_idx(i, j)) * TILE_ELEMS : nullptr);
    own_to_tile(own, R1, w);
    __syncthreads();
    HB_SK(2);
    pipe_wait(pipe, NSTAGE);
    acc_zero(acc);
    resident_mma<KMAJOR, KMAJOR>(acc, R1, R2, w, TRI_B_KLE);  // (K - S) M(j,j)'
    splitk_exchange(acc, own, R0, w);
    own_to_tile(own, R1, w);
    fence_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
      bulk_s2g(Ltw + (size_t)tri_idx(i, j) * TILE_ELEMS, R1, TILE_BYTES);
      bulk_commit();
      if (role != 0) bulk_wait_read_all();
    }
    HB_SK(3);
    if (role != 0) {
      HB_SK_OUT(1);
      return;
    }
  }

  // ------------------------------------- look-ahead: factor diagonal block i
  // A = K~(i,i) - sum_{k<=j} L(i,k) L(i,k)'.  The k = j term uses the L(i,j)
  // tile still resident in R1; the k < j terms stream through the ring.
  // rhs partial: lane (g,t) of warp w accumulates sum_k L(i,k)[8w+g][.] z_k[.]
  double rhs_part = 0.0;
  const double* zt = P.z + td.voff;
  acc_zero(acc);
  if (j >= 0) {
    resident_mma<KMAJOR, KMAJOR>(acc, R1, R1, w, TRI_SYM_LOWER);
    const double* zk = zt + 64 * j;
#pragma unroll
    for (int cb = 0; cb < 16; ++cb)
      rhs_part = fma(R1[(((w.warp << 4) + cb) << 5) + w.lane], zk[4 * cb + w.t],
                     rhs_part);
    if (threadIdx.x == 0) bulk_wait_read_all();  // the store has read R1
    __syncthreads();
  }
  auto matvec_hook = [&](int k, int h, const double* Ah) {
    const double* zk = zt + 64 * k + 32 * h;
#pragma unroll
    for (int cb = 0; cb < 8; ++cb)
      rhs_part = fma(Ah[(((w.warp << 3) + cb) << 5) + w.lane], zk[4 * cb + w.t],
                     rhs_part);
  };
  stream_gemm<KMAJOR, KMAJOR>(
      acc, max(j, 0),
      [&](int k) {
        const double* a = Lt + (size_t)tri_idx(i, k) * TILE_ELEMS;
        return TilePair{a, a, 0, 16, TRI_SYM_LOWER};
      },
      pipe, matvec_hook, w);
  HB_SK(4);
  splitk_exchange(acc, own, R0, w);
  ktile_eval<KID, true>(
      own, xi, xi, P.d, DP, 64 * i, 64 * i, td.n, sv, nv + JITTER, w,
      P.Wt ? P.Wt + (td.tile_off + tri_idx(i, i)) * TILE_ELEMS : nullptr);
  own_to_tile(own, R1, w);  // A block in tile layout
  for (int e = threadIdx.x; e < TILE_ELEMS; e += NTHREADS) R2[e] = 0.0;
  // rhs = (y - m) - sum_k L(i,k) z_k
  rhs_part += __shfl_xor_sync(0xffffffffu, rhs_part, 1);
  rhs_part += __shfl_xor_sync(0xffffffffu, rhs_part, 2);
  if (w.t == 0) {
    const int r = 8 * w.warp + w.g;
    const int gr = 64 * i + r;
    const double yv = (gr < td.n) ? P.y[td.xoff + gr] - P.theta[TH_CONST] : 0.0;
    vec[r] = yv - rhs_part;
  }
  __syncthreads();
  HB_SK(5);
#ifdef HB_STAMPS
  __shared__ long long pdbg_[4];
  potrf64_blocked(R1, R2, vec + 64, w, pdbg_);
  HB_SK(6);
  if (threadIdx.x == 0) { sk_[1] = pdbg_[0]; sk_[2] = pdbg_[1]; sk_[3] = pdbg_[2]; sk_[4] = sk_[6] - pdbg_[3]; }
#else
  potrf64_blocked(R1, R2, vec + 64, w);
#endif

  // log-determinant part and breakdown detection from the diagonal of L
  double ld_part = 0.0;
  if (threadIdx.x < 64) {
    const double l = R1[elem_off(threadIdx.x, threadIdx.x)];
    ld_part = log(l);
    const bool bad = !(l > 0.0) || !isfinite(l);
    const unsigned m = __ballot_sync(0xffffffffu, bad);
    if (m && (threadIdx.x & 31) == 0)
      atomicMin(&P.bad[task],
                (unsigned)(64 * i + (threadIdx.x & 32) + __ffs(m)));
  }
  fence_async_smem();
  ld_part = block_sum(ld_part, red);
  if (threadIdx.x == 0) {
    P.logdet[td.voff / 64 + i] = ld_part;
    bulk_s2g(Ltw + (size_t)tri_idx(i, i) * TILE_ELEMS, R1, TILE_BYTES);
    bulk_s2g(Mtw + (size_t)tri_idx(i, i) * TILE_ELEMS, R2, TILE_BYTES);
    bulk_commit();
  }
  // z_i = M(i,i) rhs : warp w handles rows 8w + g
  {
    double s = 0.0;
#pragma unroll
    for (int cb = 0; cb < 16; ++cb)
      s = fma(R2[(((w.warp << 4) + cb) << 5) + w.lane], vec[4 * cb + w.t], s);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (w.t == 0) P.z[td.voff + 64 * i + 8 * w.warp + w.g] = s;
  }
  if (threadIdx.x == 0) bulk_wait_read_all();
  HB_SK_OUT(0);
}

// ------------------------------------------------------------ alpha kernel --
// alpha = M' z (= K~^{-1} r), block c per CTA; CTA c == 0 also finishes the
// per-task NLL value  .5 z'z + sum log L_ii + .5 n log 2pi.
__global__ void __launch_bounds__(NTHREADS) k_alpha(Params P) {
  __shared__ double red[16];
  __shared__ double part[8][64];
  const TaskDesc td = P.tasks[blockIdx.y];
  const int c = blockIdx.x;
  if (c >= td.nblk) return;
  const WarpPos w;
  const double* Mt = P.Mt + td.tile_off * TILE_ELEMS;
  const double* zt = P.z + td.voff;
  // warp w handles row-blocks rb == w of every tile (k, c), k >= c; lane
  // (g,t) accumulates column 4cb + t over rows 8rb + g.
  double a16[16];
#pragma unroll
  for (int cb = 0; cb < 16; ++cb) a16[cb] = 0.0;
  for (int k = c; k < (P.with_trtri ? td.nblk : 0); ++k) {
    const double* tile = Mt + (size_t)tri_idx(k, c) * TILE_ELEMS;
    const double zv = zt[64 * k + 8 * w.warp + w.g];
#pragma unroll
    for (int cb = 0; cb < 16; ++cb)
      a16[cb] = fma(tile[(((w.warp << 4) + cb) << 5) + w.lane], zv, a16[cb]);
  }
  // reduce over g (lane bits 2..4), then over warps through smem
#pragma unroll
  for (int cb = 0; cb < 16; ++cb) {
    double v = a16[cb];
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    if (w.g == 0) part[w.warp][4 * cb + w.t] = v;
  }
  __syncthreads();
  double av = 0.0;
  if (threadIdx.x < 64) {
#pragma unroll
    for (int q = 0; q < 8; ++q) av += part[q][threadIdx.x];
    P.alpha[td.voff + 64 * c + threadIdx.x] = av;
  }
  const double s = block_sum(threadIdx.x < 64 ? av : 0.0, red);
  if (threadIdx.x == 0) P.asum[td.voff / 64 + c] = s;
  if (c == 0) {
    double zz = 0.0;
    for (int e = threadIdx.x; e < 64 * td.nblk; e += NTHREADS)
      zz = fma(zt[e], zt[e], zz);
    zz = block_sum(zz, red);
    if (threadIdx.x == 0) {
      double ld = 0.0;
      for (int b = 0; b < td.nblk; ++b) ld += P.logdet[td.voff / 64 + b];
      const unsigned bad = P.bad[blockIdx.y];
      double v = 0.5 * zz + ld + 0.5 * td.n * 1.8378770664093453;  // log(2 pi)
      if (bad != 0xffffffffu) v = __longlong_as_double(0x7ff8000000000000LL);
      P.info[blockIdx.y] = (bad != 0xffffffffu) ? (int)bad : 0;
      P.nll_task[blockIdx.y] = v;
      P.zz[blockIdx.y] = zz;
    }
  }
}

// ------------------------------------------------------- lauum + gradient ---
// CTA (task, tile (i,j)): S = (K~^{-1})(i,j) = sum_{k>=i} M(k,i)' M(k,j) on the
// tensor pipe, then the gradient contraction of G = .5 (S - alpha alpha') with
// dK/dtheta evaluated on the fly from X.  Partials per tile:
//   [0] <G, K>   [1] tr G   [2+k] <G o W, ((x_k - x'_k)/l_k)^2>
// (symmetric counterpart of an off-diagonal tile folded in by a factor 2).
template <int KID>
__global__ void __launch_bounds__(NTHREADS, 2) k_lauum_grad(Params P,
                                                             int ntile_max,
                                                             int LPT_GROUP) {
  extern __shared__ __align__(128) unsigned char smem[];
  // 1-D grid, TASK-major: the CTAs of a task are adjacent in launch order so its
  // M and W tiles are fetched from HBM once and re-read from L2 (measured: 0.32
  // GB vs 1.5 GB of DRAM reads per launch against a tile-major order, at equal
  // speed); inside a task the longest k-loops (small i) come first.
  (void)LPT_GROUP;
  const int task = blockIdx.x / ntile_max;
  const int tsel = blockIdx.x - task * ntile_max;
  if (task >= P.T) return;
  const TaskDesc td = P.tasks[task];
  const int nblk = td.nblk;
  const int ntile = nblk * (nblk + 1) / 2;
  if (tsel >= ntile) return;
  int i = 0, j = tsel;
  while (j > i) { j -= i + 1; ++i; }

  const WarpPos w;
#ifdef HB_STAMPS
  long long st_[6];
  st_[0] = clock64();
#define HB_STAMP(k) st_[k] = clock64()
#else
#define HB_STAMP(k)
#endif
  double* vec = reinterpret_cast<double*>(smem + SM_VEC);
  const int DP = xstride(P.d);
  double* xi = reinterpret_cast<double*>(smem + SM_X);
  double* xj = xi + 64 * DP;
  double* red2 = xj + 64 * DP;
  Pipe pipe = make_pipe(smem);
  __syncthreads();
  const double* Mt = P.Mt + td.tile_off * TILE_ELEMS;
  auto load_inputs = [&]() {
    load_xblock(xi, P.X, td.xoff + 64LL * i, min(64, td.n - 64 * i), P.d, DP,
                P.theta);
    load_xblock(xj, P.X, td.xoff + 64LL * j, min(64, td.n - 64 * j), P.d, DP,
                P.theta);
    if (threadIdx.x < 64)
      vec[threadIdx.x] = P.alpha[td.voff + 64 * i + threadIdx.x];
    else if (threadIdx.x < 128)
      vec[threadIdx.x] = P.alpha[td.voff + 64 * j + threadIdx.x - 64];
  };

  double acc[4][4][2];
  double own[2][4][2];
  acc_zero(acc);
  const int wstage = stream_gemm<MNMAJOR, MNMAJOR>(
      acc, nblk - i,
      [&](int kk) {
        const int k = i + kk;
        int fl = (i == j) ? TRI_SYM_LOWER : TRI_NONE;
        if (kk == 0) fl |= TRI_A_KGE | (i == j ? TRI_B_KGE : 0);
        return TilePair{Mt + (size_t)tri_idx(k, i) * TILE_ELEMS,
                        Mt + (size_t)tri_idx(k, j) * TILE_ELEMS, 0, 16, fl};
      },
      pipe, NoHook(), w, load_inputs,
      [&](int stage) {  // the W tile rides behind the last M half-steps
        if (threadIdx.x == 0)
          load_tile_async(pipe, pipe.ring + stage * STAGE_ELEMS,
                          P.Wt + (td.tile_off + tri_idx(i, j)) * TILE_ELEMS);
      });
  HB_STAMP(2);
  const double* Ws = pipe.ring + wstage * STAGE_ELEMS;
  splitk_exchange(acc, own, pipe.ring + ((wstage + 1) % NSTAGE) * STAGE_ELEMS, w);
  HB_STAMP(3);
  pipe_wait(pipe, NSTAGE);

  // G = .5 (S - alpha alpha');  d nll/d l_k  needs  sum G W D_k^2  with the
  // pair weights W saved by the factorisation (no second exp / distance pass);
  // d nll/d noise = tr G;  d nll/d signal follows in closed form from tr G
  // (k_reduce_final), so no kernel value is needed here at all.
  const double mult = (i == j) ? 1.0 : 2.0;
  double g_nv = 0.0;
  double gw[2][4][2];
#pragma unroll
  for (int fi = 0; fi < 2; ++fi)
#pragma unroll
    for (int fn = 0; fn < 4; ++fn) {
      const int r = own_row(w, fi), c0 = own_col(w, fn, 0);
      const double2 wv = *reinterpret_cast<const double2*>(Ws + elem_off(r, c0));
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = c0 + e;
        const int gr = 64 * i + r, gc = 64 * j + c;
        double G = 0.5 * (own[fi][fn][e] - vec[r] * vec[64 + c]);
        // off-diagonal tiles stand for their mirror image too; a diagonal
        // tile contributes its lower triangle (x2) and its diagonal (the
        // strictly upper part was not computed)
        double me = mult;
        if (i == j) me = (r > c) ? 2.0 : (r == c ? 1.0 : 0.0);
        if (me == 0.0 || gr >= td.n || gc >= td.n) G = 0.0;
        if (gr == gc) g_nv += G;
        gw[fi][fn][e] = me * G * (e ? wv.y : wv.x);
      }
    }
  const double g_sv = 0.0;  // slot kept for layout; see k_reduce_final
  HB_STAMP(4);
  // per-thread partials for the lengthscale terms (two independent chains per
  // dimension), then ONE batched butterfly over all partials so the shuffle /
  // add latencies of the 2 + d reductions overlap instead of serialising
  const int np = 2 + P.d;
  for (int p0 = 0; p0 < np; p0 += 6) {
    double part[6];
#pragma unroll
    for (int u = 0; u < 6; ++u) {
      const int p = p0 + u;
      double v = 0.0;
      if (p == 0) v = g_sv;
      else if (p == 1) v = g_nv;
      else if (p < np) {
        const int k = p - 2;
        double v0 = 0.0, v1 = 0.0;
        const double xr0 = xi[own_row(w, 0) * DP + k];
        const double xr1 = xi[own_row(w, 1) * DP + k];
#pragma unroll
        for (int fn = 0; fn < 4; ++fn)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const double xc = xj[own_col(w, fn, e) * DP + k];
            const double d0 = xr0 - xc, d1 = xr1 - xc;
            v0 = fma(gw[0][fn][e], d0 * d0, v0);
            v1 = fma(gw[1][fn][e], d1 * d1, v1);
          }
        v = v0 + v1;
      }
      part[u] = v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int u = 0; u < 6; ++u)
        part[u] += __shfl_xor_sync(0xffffffffu, part[u], o);
    if (w.lane == 0) {
#pragma unroll
      for (int u = 0; u < 6; ++u)
        if (p0 + u < np) red2[w.warp * GP_STRIDE + p0 + u] = part[u];
    }
  }
  __syncthreads();
  if ((int)threadIdx.x < np) {
    double sum = 0.0;
#pragma unroll
    for (int q = 0; q < NTHREADS / 32; ++q) sum += red2[q * GP_STRIDE + threadIdx.x];
    P.gpart[(td.tile_off + tsel) * GP_STRIDE + threadIdx.x] = sum;
  }
#ifdef HB_STAMPS
  HB_STAMP(5);
  if (threadIdx.x == 0 && P.stamps) {
    long long* o = P.stamps + ((size_t)task * ntile_max + tsel) * 8;
    unsigned smid;
    asm("mov.u32 %0, %%smid;" : "=r"(smid));
    o[0] = st_[0]; o[2] = st_[2]; o[3] = st_[3]; o[4] = st_[4]; o[5] = st_[5];
    o[6] = smid; o[7] = nblk - i;
  }
#endif
}

// per-task sum of the tile partials (fixed order -> deterministic)
__global__ void k_reduce_task(Params P) {
  const TaskDesc td = P.tasks[blockIdx.x];
  const int ntile = td.nblk * (td.nblk + 1) / 2;
  const int p = threadIdx.x;
  if (p >= 2 + P.d) return;
  double s = 0.0;
  for (int t = 0; t < ntile; ++t) s += P.gpart[(td.tile_off + t) * GP_STRIDE + p];
  P.gtask[(size_t)blockIdx.x * GP_STRIDE + p] = s;
}

// sums over tasks + chain rule.  out[0] = sum nll, out[1+p] = sum d nll/d raw_p,
// out[1+P] = #non-empty tasks.  One CTA: warp q accumulates output q (tasks
// strided over its lanes, fixed order -> bit-reproducible), one shuffle tree.
__global__ void __launch_bounds__(1024) k_reduce_final(
    Params P, double* __restrict__ out, double* __restrict__ nll_task_out) {
  const int np = 3 + P.d;
  const int nwarp = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double sv = P.theta[TH_SV];
  for (int q = warp; q <= np + 1; q += nwarp) {
    double v = 0.0;
    for (int t = lane; t < P.T; t += 32) {
      const TaskDesc td = P.tasks[t];
      if (td.n == 0) continue;
      if (q == 0) v += P.nll_task[t];
      else if (q == np + 1) v += 1.0;
      else if (q == 1) {  // constant: -sum alpha
        for (int b = 0; b < td.nblk; ++b) v -= P.asum[td.voff / 64 + b];
      } else if (q == 2) {
        // signal variance: tr(G K) = .5 (n - z'z) - (nv + eps) tr G
        v += 0.5 * (td.n - P.zz[t]) -
             (P.theta[TH_NV] + JITTER) * P.gtask[(size_t)t * GP_STRIDE + 1];
      } else v += P.gtask[(size_t)t * GP_STRIDE + (q - 2)];
    }
    v = warp_sum(v);
    if (lane == 0) {
      if (q >= 1 && q <= np && v != 0.0) {  // (an empty batch stays exactly 0)
        const int p = q - 1;
        if (p == 1) v /= sv;                                    // tr(G K)/sv
        if (p >= 3) v *= P.theta[TH_INVLS + (p - 3)];           // /l_k
        v *= P.theta[TH_CHAIN + p];
      }
      out[q] = v;
    }
  }
  if (nll_task_out)
    for (int t = threadIdx.x; t < P.T; t += blockDim.x)
      nll_task_out[t] = P.tasks[t].n ? P.nll_task[t] : 0.0;
}

// optax.adam step with the accept / stop semantics of gp_utils/gp.py:135-146
__global__ void k_adam(int P_, double* raw, double* m, double* v,
                       double* accepted, const double* sums, double* scal,
                       double lr, double b1, double b2, double eps, int tie_ls) {
  const int p = threadIdx.x;
  const double cnt = sums[1 + P_];
  const double loss = cnt > 0.0 ? sums[0] / cnt : 0.0;
  const double t_old = scal[1];
  const bool stopped = scal[2] != 0.0;
  const bool fin = isfinite(loss);
  __syncthreads();
  if (p == 0) {
    scal[0] = loss;
    if (!stopped) {
      if (fin) { scal[1] = t_old + 1.0; scal[3] += 1.0; }
      else scal[2] = 1.0;
    }
  }
  if (p < P_ && !stopped && fin) {
    const double t = t_old + 1.0;
    double g = cnt > 0.0 ? sums[1 + p] / cnt : 0.0;
    if (tie_ls && p >= 3) {
      g = 0.0;
      for (int k = 3; k < P_; ++k) g += sums[1 + k];
      g = cnt > 0.0 ? g / cnt : 0.0;
    }
    const double r0 = raw[p];
    accepted[p] = r0;
    const double mn = b1 * m[p] + (1.0 - b1) * g;
    const double vn = b2 * v[p] + (1.0 - b2) * g * g;
    m[p] = mn;
    v[p] = vn;
    const double mhat = mn / (1.0 - pow(b1, t));
    const double vhat = vn / (1.0 - pow(b2, t));
    raw[p] = r0 - lr * mhat / (sqrt(vhat) + eps);
  }
}

// -------------------------------------------------- stand-alone Gram matrix --
// kernel.covariance_matrix.matrix_map: out (n1, n2) row-major; HBM-write bound.
template <int KID>
__global__ void __launch_bounds__(NTHREADS) k_kernel_matrix(
    const double* __restrict__ X1, long long n1, const double* __restrict__ X2,
    long long n2, int d, const double* __restrict__ theta, int add_noise,
    double jitter, double* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int DP = xstride(d);
  double* x1 = reinterpret_cast<double*>(smem);
  double* x2 = x1 + 64 * DP;
  const long long r0 = 64LL * blockIdx.y, c0 = 64LL * blockIdx.x;
  load_xblock(x1, X1, r0, (int)min(64LL, n1 - r0), d, DP, theta);
  load_xblock(x2, X2, c0, (int)min(64LL, n2 - c0), d, DP, theta);
  __syncthreads();
  const double sv = theta[TH_SV];
  const double diag_add = add_noise ? theta[TH_NV] + jitter : 0.0;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  double r2[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) r2[a][b] = 0.0;
  for (int k = 0; k < d; ++k) {
    double xr[4], xc[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) xr[a] = x1[(ty + 16 * a) * DP + k];
#pragma unroll
    for (int b = 0; b < 4; ++b) xc[b] = x2[(tx + 16 * b) * DP + k];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const double df = xr[a] - xc[b];
        r2[a][b] = fma(df, df, r2[a][b]);
      }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const long long r = r0 + ty + 16 * a, c = c0 + tx + 16 * b;
      if (r < n1 && c < n2) {
        double k, wg;
        kern_eval<KID>(r2[a][b], sv, k, wg);
        if (r == c) k += diag_add;
        out[r * n2 + c] = k;
      }
    }
}

__global__ void k_fill(double* out, long long n, const double* theta, int idx,
                       double add) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) out[i] = theta[idx] + add;
}

// packed L tiles -> row-major (n, n) lower factors with zeroed upper triangle
__global__ void __launch_bounds__(NTHREADS) k_unpack_chol(Params P,
                                                          double* __restrict__ out) {
  const TaskDesc td = P.tasks[blockIdx.y];
  const int ntile = td.nblk * (td.nblk + 1) / 2;
  if ((int)blockIdx.x >= ntile) return;
  int i = 0, j = blockIdx.x;
  while (j > i) { j -= i + 1; ++i; }
  const double* tile = P.Lt + (td.tile_off + blockIdx.x) * TILE_ELEMS;
  double* o = out + td.chol_off;
  const long long n = td.n;
  for (int e = threadIdx.x; e < TILE_ELEMS; e += NTHREADS) {
    const int r = e >> 6, c = e & 63;
    const long long gr = 64LL * i + r, gc = 64LL * j + c;
    if (gr < n && gc < n) {
      o[gr * n + gc] = tile[elem_off(r, c)];
      if (i != j) o[gc * n + gr] = 0.0;
    }
  }
}

// compact 64-padded per-task vectors into the caller's (sum n,) layout
__global__ void k_unpad_vec(Params P, const double* __restrict__ src,
                            double* __restrict__ dst) {
  const TaskDesc td = P.tasks[blockIdx.y];
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < td.n;
       e += gridDim.x * blockDim.x)
    dst[td.xoff + e] = src[td.voff + e];
}

__global__ void k_copy_scalars(const double* src, double* dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}
__global__ void k_copy_info(const int* src, int* dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

// ------------------------------------------------------------------ predict --
// gp.predict (gp_utils/gp.py:242-305) against the cached M = L^{-1} tiles:
//   K*  = k(X_obs, X_q)                      k_kstar   (tiles to L2-resident scratch)
//   mu  = K*' alpha + m                      k_kstar   (partials per obs block)
//   V   = L^{-1} K* = M K*                   k_predict_gemm (tensor pipe)
//   var = k(xq,xq) - colsum(V o V)           k_predict_gemm / k_predict_final
// followed by the acquisition epilogue (bo_utils/acfun.py:96-142).
struct PredParams {
  const double* X;      // (n, d) observations
  const double* Xq;     // (nq, d) queries of this pass
  const double* theta;
  const double* Mt;     // cache: packed M tiles
  const double* alpha;  // cache: 64-padded alpha
  double* kst;          // scratch: [nqc][nblk] tiles
  double* mupart;       // scratch: [nblk][nqc*64]
  double* vpart;        // scratch: [nblk][nqc*64]
  int n, nblk, d;
  long long nq;
  int nqc;
};

template <int KID>
__global__ void __launch_bounds__(NTHREADS) k_kstar(PredParams Q) {
  extern __shared__ __align__(128) unsigned char smem[];
  double* tile = reinterpret_cast<double*>(smem);
  double* part = tile + TILE_ELEMS;          // [8][64]
  double* al = part + 8 * 64;                // [64]
  const int DP = xstride(Q.d);
  double* xi = al + 64;
  double* xj = xi + 64 * DP;
  const int k = blockIdx.x, qc = blockIdx.y;
  load_xblock(xi, Q.X, 64LL * k, min(64, Q.n - 64 * k), Q.d, DP, Q.theta);
  load_xblock(xj, Q.Xq, 64LL * qc, (int)min(64LL, Q.nq - 64LL * qc), Q.d, DP,
              Q.theta);
  if (threadIdx.x < 64) al[threadIdx.x] = Q.alpha[64 * k + threadIdx.x];
  __syncthreads();
  const double sv = Q.theta[TH_SV];
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  double r2[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) r2[a][b] = 0.0;
  for (int kk = 0; kk < Q.d; ++kk) {
    double xr[4], xc[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) xr[a] = xi[(ty + 16 * a) * DP + kk];
#pragma unroll
    for (int b = 0; b < 4; ++b) xc[b] = xj[(tx + 16 * b) * DP + kk];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const double df = xr[a] - xc[b];
        r2[a][b] = fma(df, df, r2[a][b]);
      }
  }
  double mu[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int r = ty + 16 * a, c = tx + 16 * b;
      double kv, wg;
      kern_eval<KID>(r2[a][b], sv, kv, wg);
      if (64 * k + r >= Q.n) kv = 0.0;
      tile[elem_off(r, c)] = kv;
      mu[b] = fma(kv, al[r], mu[b]);
    }
  // reduce mu over the 16 ty values: lanes differ in ty by bit 4, then warps
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    double v = mu[b] + __shfl_xor_sync(0xffffffffu, mu[b], 16);
    if (lane < 16) part[warp * 64 + tx + 16 * b] = v;
  }
  fence_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    bulk_s2g(Q.kst + ((size_t)qc * Q.nblk + k) * TILE_ELEMS, tile,
             TILE_ELEMS * sizeof(double));
    bulk_commit();
  }
  if (threadIdx.x < 64) {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += part[q * 64 + threadIdx.x];
    Q.mupart[(size_t)k * Q.nqc * 64 + 64 * qc + threadIdx.x] = s;
  }
  if (threadIdx.x == 0) bulk_wait_all();
}

__global__ void __launch_bounds__(NTHREADS, 2) k_predict_gemm(PredParams Q) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int i = Q.nblk - 1 - blockIdx.x;  // heavy rows first
  const int qc = blockIdx.y;
  const WarpPos w;
  Pipe pipe = make_pipe(smem);
  double* part = pipe.ring + STAGE_ELEMS;  // [4][64] in R1 (ring idle by then)
  __syncthreads();
  double acc[4][4][2];
  double own[2][4][2];
  acc_zero(acc);
  stream_gemm<KMAJOR, MNMAJOR>(
      acc, i + 1,
      [&](int k) {
        return TilePair{Q.Mt + (size_t)tri_idx(i, k) * TILE_ELEMS,
                        Q.kst + ((size_t)qc * Q.nblk + k) * TILE_ELEMS, 0, 16,
                        k == i ? TRI_A_KLE : TRI_NONE};
      },
      pipe, NoHook(), w);
  splitk_exchange(acc, own, pipe.ring, w);
  // column sums of V o V: reduce over fi, g (lane bits 2..4), then (wm, wk)
#pragma unroll
  for (int fn = 0; fn < 4; ++fn)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      double v = own[0][fn][e] * own[0][fn][e] + own[1][fn][e] * own[1][fn][e];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (w.g == 0) part[(w.wm * 2 + w.wk) * 64 + own_col(w, fn, e)] = v;
    }
  __syncthreads();
  if (threadIdx.x < 64)
    Q.vpart[(size_t)i * Q.nqc * 64 + 64 * qc + threadIdx.x] =
        part[threadIdx.x] + part[64 + threadIdx.x] + part[128 + threadIdx.x] +
        part[192 + threadIdx.x];
}

// standard normal pdf / cdf as jax.scipy.stats.norm (acfun.py:109-110)
__device__ __forceinline__ double acq_eval(int acq_id, double mu, double var,
                                           double param) {
  const double sd = sqrt(var);
  if (acq_id == 3) return mu + param * sd;           // ucb_sub
  const double gamma = (param - mu) / sd;
  if (acq_id == 2) return -gamma;                    // probability_of_improvement_sub
  const double pdf = 0.3989422804014327 * exp(-0.5 * gamma * gamma);
  const double cdf = 0.5 * erfc(-gamma * 0.7071067811865476);
  return (pdf - gamma * (1.0 - cdf)) * sd;           // expected_improvement_sub
}

__global__ void k_predict_final(PredParams Q, double noise_flag, double scale,
                                int acq_id, double acq_param, double* mu_out,
                                double* var_out, double* acq_out) {
  const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (q >= Q.nq) return;
  double mu = Q.theta[TH_CONST], v = 0.0;
  const size_t ld = (size_t)Q.nqc * 64;
  for (int k = 0; k < Q.nblk; ++k) {
    mu += Q.mupart[k * ld + q];
    v += Q.vpart[k * ld + q];
  }
  const double var = (Q.theta[TH_SV] - v + noise_flag * Q.theta[TH_NV]) * scale;
  if (mu_out) mu_out[q] = mu;
  if (var_out) var_out[q] = var;
  if (acq_out) acq_out[q] = acq_eval(acq_id, mu, var, acq_param);
}

// prior prediction (no observations, gp.py:275-282) shares k_predict_final with
// nblk = 0.  acfun_sub alone on given vectors:
__global__ void k_acq(int acq_id, double param, long long nq, const double* mu,
                      const double* var, double* out) {
  const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (q < nq) out[q] = acq_eval(acq_id, mu[q], var[q], param);
}

}  // namespace hb
