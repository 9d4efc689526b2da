// hyperbo_b200 C ABI (include/hyperbo_b200.h): host-side planning, workspace
// arena and kernel launches.  Everything is enqueued on the caller's stream.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/hyperbo_b200.h"
#include "hb_kernels.cuh"

using namespace hb;

namespace {

struct Buf {
  void* p = nullptr;
  size_t cap = 0;
};

struct Plan {
  std::vector<int64_t> offs;
  int T = 0, d = 0;
  std::vector<TaskDesc> tasks;
  TaskDesc* tasks_d = nullptr;
  size_t tasks_cap = 0;
  int nblk_max = 0;
  long long total_tiles = 0, total_blocks = 0, sum_n = 0, chol_elems = 0;
  uint64_t stamp = 0;
  bool uploaded = false;
};

constexpr int NPLAN = 4;

}  // namespace

struct hb_handle_s {
  int device = 0;
  int dtype = HB_F64;
  std::string err;
  int64_t launches = 0;
  uint64_t clock = 0;
  Plan plans[NPLAN];
  Buf theta, Lt, Mt, Wt, zz, z, alpha, logdet, asum, nll_task, gpart, gtask, info, bad,
      sums, kst, mupart, vpart, pcache, stamps;
  bool attr_set = false;
  int smem_d = -1;
  bool profiling = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_ev[HB_PROFILE_SECTIONS];
};

namespace {

#define HB_CUDA(call)                                                         \
  do {                                                                        \
    cudaError_t e_ = (call);                                                  \
    if (e_ != cudaSuccess) {                                                  \
      h->err = std::string(#call) + ": " + cudaGetErrorString(e_);            \
      return HB_ERR_CUDA;                                                     \
    }                                                                         \
  } while (0)

int fail(hb_handle_t h, int code, const char* msg) {
  if (h) h->err = msg;
  return code;
}

int ensure(hb_handle_t h, Buf& b, size_t bytes) {
  if (bytes <= b.cap) return HB_OK;
  if (b.p) HB_CUDA(cudaFree(b.p));
  b.p = nullptr;
  b.cap = 0;
  size_t want = bytes + bytes / 8 + 256;
  HB_CUDA(cudaMalloc(&b.p, want));
  b.cap = want;
  return HB_OK;
}

size_t total_ws(hb_handle_t h) {
  const Buf* all[] = {&h->theta, &h->Lt,    &h->Mt,  &h->Wt, &h->zz,  &h->z,    &h->alpha,
                      &h->logdet, &h->asum, &h->nll_task, &h->gpart, &h->gtask,
                      &h->info,  &h->bad,   &h->sums,  &h->kst,  &h->mupart,
                      &h->vpart, &h->pcache};
  size_t s = 0;
  for (auto* b : all) s += b->cap;
  for (auto& p : h->plans) s += p.tasks_cap;
  return s;
}

// find or build the plan for (T, offs, d); uploads descriptors when new
int get_plan(hb_handle_t h, int T, const int64_t* offs, int d, cudaStream_t st,
             Plan** out) {
  ++h->clock;
  Plan* lru = &h->plans[0];
  for (auto& p : h->plans) {
    if (p.uploaded && p.T == T && p.d == d && (int)p.offs.size() == T + 1 &&
        std::memcmp(p.offs.data(), offs, sizeof(int64_t) * (T + 1)) == 0) {
      p.stamp = h->clock;
      *out = &p;
      return HB_OK;
    }
    if (p.stamp < lru->stamp) lru = &p;
  }
  Plan& p = *lru;
  p.uploaded = false;
  p.T = T;
  p.d = d;
  p.offs.assign(offs, offs + T + 1);
  p.tasks.resize(T);
  p.nblk_max = 0;
  long long tiles = 0, blocks = 0, chol = 0;
  for (int t = 0; t < T; ++t) {
    const int64_t n = offs[t + 1] - offs[t];
    if (n < 0 || n > (1 << 20)) return fail(h, HB_ERR_BAD_ARG, "bad offs");
    TaskDesc& td = p.tasks[t];
    td.n = (int)n;
    td.nblk = (int)((n + TB - 1) / TB);
    td.xoff = offs[t];
    td.voff = blocks * TB;
    td.tile_off = tiles;
    td.chol_off = chol;
    tiles += (long long)td.nblk * (td.nblk + 1) / 2;
    blocks += td.nblk;
    chol += n * n;
    p.nblk_max = std::max(p.nblk_max, td.nblk);
  }
  p.total_tiles = tiles;
  p.total_blocks = blocks;
  p.sum_n = offs[T] - offs[0];
  p.chol_elems = chol;
  const size_t bytes = sizeof(TaskDesc) * (size_t)std::max(T, 1);
  if (bytes > p.tasks_cap) {
    if (p.tasks_d) HB_CUDA(cudaFree(p.tasks_d));
    p.tasks_d = nullptr;
    HB_CUDA(cudaMalloc(&p.tasks_d, bytes));
    p.tasks_cap = bytes;
  }
  if (T > 0)
    HB_CUDA(cudaMemcpyAsync(p.tasks_d, p.tasks.data(), sizeof(TaskDesc) * T,
                            cudaMemcpyHostToDevice, st));
  p.stamp = h->clock;
  p.uploaded = true;
  *out = &p;
  return HB_OK;
}

int ensure_ws(hb_handle_t h, const Plan& p, bool grad) {
  int rc;
  const size_t tiles = (size_t)std::max<long long>(p.total_tiles, 1);
  const size_t blocks = (size_t)std::max<long long>(p.total_blocks, 1);
  const size_t T = (size_t)std::max(p.T, 1);
  if ((rc = ensure(h, h->theta, TH_SIZE * 8))) return rc;
  if ((rc = ensure(h, h->Lt, tiles * TILE_ELEMS * 8))) return rc;
  if ((rc = ensure(h, h->Mt, tiles * TILE_ELEMS * 8))) return rc;
  if ((rc = ensure(h, h->z, blocks * TB * 8))) return rc;
  if ((rc = ensure(h, h->alpha, blocks * TB * 8))) return rc;
  if ((rc = ensure(h, h->logdet, blocks * 8))) return rc;
  if ((rc = ensure(h, h->asum, blocks * 8))) return rc;
  if ((rc = ensure(h, h->nll_task, T * 8))) return rc;
  if ((rc = ensure(h, h->zz, T * 8))) return rc;
  if ((rc = ensure(h, h->info, T * 4))) return rc;
  if ((rc = ensure(h, h->bad, T * 4))) return rc;
  if ((rc = ensure(h, h->sums, (3 + MAX_DIM + 2) * 8))) return rc;
#ifdef HB_STAMPS
  if ((rc = ensure(h, h->stamps, (size_t)(p.nblk_max + 1) * T * 8 * 64 + tiles * 64 + 1024))) return rc;
#endif
  if (grad) {
    if ((rc = ensure(h, h->Wt, tiles * TILE_ELEMS * 8))) return rc;
    if ((rc = ensure(h, h->gpart, tiles * GP_STRIDE * 8))) return rc;
    if ((rc = ensure(h, h->gtask, T * GP_STRIDE * 8))) return rc;
  }
  return HB_OK;
}

template <int KID>
int set_attrs_k(hb_handle_t h, int d) {
  HB_CUDA(cudaFuncSetAttribute(k_step<KID>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)step_smem_bytes(MAX_DIM)));
  HB_CUDA(cudaFuncSetAttribute(k_lauum_grad<KID>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)lauum_smem_bytes(MAX_DIM)));
  // two ~106 KiB CTAs per SM need the full shared-memory carveout
  HB_CUDA(cudaFuncSetAttribute(k_step<KID>,
                               cudaFuncAttributePreferredSharedMemoryCarveout,
                               cudaSharedmemCarveoutMaxShared));
  HB_CUDA(cudaFuncSetAttribute(k_lauum_grad<KID>,
                               cudaFuncAttributePreferredSharedMemoryCarveout,
                               cudaSharedmemCarveoutMaxShared));
  HB_CUDA(cudaFuncSetAttribute(k_kstar<KID>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)((TILE_ELEMS + 8 * 64 + 64 +
                                      2 * 64 * xstride(MAX_DIM)) * 8)));
  (void)d;
  return HB_OK;
}

int set_attrs(hb_handle_t h) {
  if (h->attr_set) return HB_OK;
  int rc;
  if ((rc = set_attrs_k<0>(h, 0))) return rc;
  if ((rc = set_attrs_k<1>(h, 0))) return rc;
  if ((rc = set_attrs_k<2>(h, 0))) return rc;
  HB_CUDA(cudaFuncSetAttribute(k_predict_gemm,
                               cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)PREDICT_SMEM_BYTES));
  HB_CUDA(cudaFuncSetAttribute(k_predict_gemm,
                               cudaFuncAttributePreferredSharedMemoryCarveout,
                               cudaSharedmemCarveoutMaxShared));
  h->attr_set = true;
  return HB_OK;
}

int check_common(hb_handle_t h, int kernel_id, int mean_id, int d) {
  if (!h) return HB_ERR_BAD_ARG;
  if (h->dtype != HB_F64)
    return fail(h, HB_ERR_UNSUPPORTED, "only HB_F64 handles are implemented");
  if (kernel_id < 0 || kernel_id > 2) return fail(h, HB_ERR_BAD_ARG, "kernel_id");
  if (mean_id < 0 || mean_id > 1) return fail(h, HB_ERR_BAD_ARG, "mean_id");
  if (d < 1) return fail(h, HB_ERR_BAD_ARG, "d < 1");
  if (d > MAX_DIM) return fail(h, HB_ERR_UNSUPPORTED, "d > HB_MAX_DIM");
  return HB_OK;
}

Params make_params(hb_handle_t h, const Plan& p, int kernel_id, int mean_id,
                   const void* X, const void* y, int with_trtri,
                   bool grad = false) {
  Params P;
  P.tasks = p.tasks_d;
  P.T = p.T;
  P.d = p.d;
  P.kernel_id = kernel_id;
  P.mean_id = mean_id;
  P.with_trtri = with_trtri;
  P.X = (const double*)X;
  P.y = (const double*)y;
  P.theta = (const double*)h->theta.p;
  P.Lt = (double*)h->Lt.p;
  P.Mt = (double*)h->Mt.p;
  P.Wt = grad ? (double*)h->Wt.p : nullptr;
  P.zz = (double*)h->zz.p;
  P.z = (double*)h->z.p;
  P.alpha = (double*)h->alpha.p;
  P.logdet = (double*)h->logdet.p;
  P.asum = (double*)h->asum.p;
  P.nll_task = (double*)h->nll_task.p;
  P.gpart = (double*)h->gpart.p;
  P.gtask = (double*)h->gtask.p;
  P.info = (int*)h->info.p;
  P.bad = (unsigned*)h->bad.p;
  P.stamps = (long long*)h->stamps.p;
  return P;
}

#define HB_LAUNCH_CHECK()                                  \
  do {                                                     \
    ++h->launches;                                         \
    cudaError_t e_ = cudaGetLastError();                   \
    if (e_ != cudaSuccess) {                               \
      h->err = std::string("launch: ") + cudaGetErrorString(e_); \
      return HB_ERR_CUDA;                                  \
    }                                                      \
  } while (0)

struct Section {
  hb_handle_t h;
  int id;
  cudaStream_t st;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  Section(hb_handle_t h_, int id_, cudaStream_t st_) : h(h_), id(id_), st(st_) {
    if (!h->profiling) return;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) {
      e0 = e1 = nullptr;
      return;
    }
    cudaEventRecord(e0, st);
  }
  ~Section() {
    if (!e0) return;
    cudaEventRecord(e1, st);
    h->prof_ev[id].emplace_back(e0, e1);
  }
};

template <int KID>
int run_factor_k(hb_handle_t h, const Plan& p, const Params& P, cudaStream_t st) {
  const size_t smem = step_smem_bytes(p.d);
  for (int j = -1; j < p.nblk_max; ++j) {
    const int np = (j < 0) ? 1 : std::max(0, p.nblk_max - 1 - j);
    const int nt = (P.with_trtri && j >= 1) ? j : 0;
    if (np + nt == 0) continue;
    const int nroles = np + nt;
    const int lpt = std::min(p.T, LPT_GROUP_MAX);
    const int ngrp = (p.T + lpt - 1) / lpt;
    dim3 grid((unsigned)nroles * lpt * ngrp);
    k_step<KID><<<grid, NTHREADS, smem, st>>>(P, j, nroles, lpt);
    HB_LAUNCH_CHECK();
  }
  return HB_OK;
}

// prep + factorisation (+ M) + alpha / per-task nll
int run_factor(hb_handle_t h, const Plan& p, const Params& P, const void* raw,
               uint64_t warp_mask, cudaStream_t st) {
  {
    Section sec(h, 0, st);
    k_prep<<<1, 64, 0, st>>>((const double*)raw, warp_mask, p.d, P.mean_id,
                             (double*)h->theta.p, P.bad, p.T);
    HB_LAUNCH_CHECK();
    if (p.T == 0 || p.nblk_max == 0) return HB_OK;
    int rc;
    switch (P.kernel_id) {
      case 0: rc = run_factor_k<0>(h, p, P, st); break;
      case 1: rc = run_factor_k<1>(h, p, P, st); break;
      default: rc = run_factor_k<2>(h, p, P, st); break;
    }
    if (rc) return rc;
  }
  Section sec(h, 1, st);
  k_alpha<<<dim3(p.nblk_max, p.T), NTHREADS, 0, st>>>(P);
  HB_LAUNCH_CHECK();
  return HB_OK;
}


int factorize_impl(hb_handle_t h, int kernel_id, int mean_id, int T,
                   const int64_t* offs, int d, const void* X, const void* y,
                   const void* raw, uint64_t warp_mask, int with_trtri,
                   void* chol_out, void* alpha_out, void* nll_out,
                   int32_t* info_out, cudaStream_t st, Plan** plan_out,
                   Params* P_out) {
  int rc = check_common(h, kernel_id, mean_id, d);
  if (rc) return rc;
  if (T < 0 || !offs || !raw || (T > 0 && (!X || !y)))
    return fail(h, HB_ERR_BAD_ARG, "null arg");
  if ((rc = set_attrs(h))) return rc;
  Plan* p;
  if ((rc = get_plan(h, T, offs, d, st, &p))) return rc;
  if ((rc = ensure_ws(h, *p, false))) return rc;
  Params P = make_params(h, *p, kernel_id, mean_id, X, y, with_trtri);
  *plan_out = p;
  *P_out = P;
  if ((rc = run_factor(h, *p, P, raw, warp_mask, st))) return rc;
  if (T == 0 || p->nblk_max == 0) return HB_OK;
  if (chol_out) {
    const int nt = p->nblk_max * (p->nblk_max + 1) / 2;
    k_unpack_chol<<<dim3(nt, T), NTHREADS, 0, st>>>(P, (double*)chol_out);
    HB_LAUNCH_CHECK();
  }
  if (alpha_out && with_trtri) {
    k_unpad_vec<<<dim3(std::max(1, (p->nblk_max * 64 + 255) / 256), T), 256, 0,
                  st>>>(P, P.alpha, (double*)alpha_out);
    HB_LAUNCH_CHECK();
  }
  if (nll_out) {
    k_copy_scalars<<<(T + 255) / 256, 256, 0, st>>>(P.nll_task, (double*)nll_out, T);
    HB_LAUNCH_CHECK();
  }
  if (info_out) {
    k_copy_info<<<(T + 255) / 256, 256, 0, st>>>(P.info, info_out, T);
    HB_LAUNCH_CHECK();
  }
  return HB_OK;
}

template <int KID>
void launch_kstar(const PredParams& Q, size_t smem, cudaStream_t st) {
  k_kstar<KID><<<dim3(Q.nblk, Q.nqc), NTHREADS, smem, st>>>(Q);
}

}  // namespace

extern "C" {

const char* hb_version(void) { return "hyperbo_b200 0.1 (sm_100a, fp64 DMMA)"; }

int hb_create(hb_handle_t* out, int device, int dtype) {
  if (!out) return HB_ERR_BAD_ARG;
  *out = nullptr;
  if (dtype != HB_F64 && dtype != HB_F32) return HB_ERR_BAD_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return HB_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= ndev) return HB_ERR_BAD_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return HB_ERR_CUDA;
  hb_handle_t h = new hb_handle_s();
  h->device = device;
  h->dtype = dtype;
  *out = h;
  return HB_OK;
}

int hb_destroy(hb_handle_t h) {
  if (!h) return HB_ERR_BAD_ARG;
  Buf* all[] = {&h->theta, &h->Lt,    &h->Mt,  &h->Wt, &h->zz,  &h->z,    &h->alpha,
                &h->logdet, &h->asum, &h->nll_task, &h->gpart, &h->gtask,
                &h->info,  &h->bad,   &h->sums,  &h->kst,  &h->mupart,
                &h->vpart, &h->pcache};
  for (auto* b : all)
    if (b->p) cudaFree(b->p);
  for (auto& p : h->plans)
    if (p.tasks_d) cudaFree(p.tasks_d);
  delete h;
  return HB_OK;
}

const char* hb_last_error(hb_handle_t h) { return h ? h->err.c_str() : "null handle"; }
int64_t hb_launch_count(hb_handle_t h) { return h ? h->launches : 0; }
int64_t hb_workspace_bytes(hb_handle_t h) { return h ? (int64_t)total_ws(h) : 0; }

#ifdef HB_STAMPS
int hb_debug_stamps(hb_handle_t h, long long* host_out, int64_t n) {
  cudaDeviceSynchronize();
  return cudaMemcpy(host_out, h->stamps.p, n * 8, cudaMemcpyDeviceToHost);
}
#endif

int hb_profile_enable(hb_handle_t h, int enable) {
  if (!h) return HB_ERR_BAD_ARG;
  for (auto& v : h->prof_ev) {
    for (auto& e : v) {
      cudaEventDestroy(e.first);
      cudaEventDestroy(e.second);
    }
    v.clear();
  }
  h->profiling = enable != 0;
  return HB_OK;
}

int hb_profile_read(hb_handle_t h, double* ms_out, int64_t* count_out) {
  if (!h || !ms_out || !count_out) return HB_ERR_BAD_ARG;
  for (int s = 0; s < HB_PROFILE_SECTIONS; ++s) {
    double tot = 0.0;
    for (auto& e : h->prof_ev[s]) {
      HB_CUDA(cudaEventSynchronize(e.second));
      float ms = 0.f;
      HB_CUDA(cudaEventElapsedTime(&ms, e.first, e.second));
      tot += ms;
    }
    ms_out[s] = tot;
    count_out[s] = (int64_t)h->prof_ev[s].size();
  }
  return HB_OK;
}

int hb_kernel_matrix(hb_handle_t h, int kernel_id, const void* X1, int64_t n1,
                     const void* X2, int64_t n2, int d, const void* raw,
                     uint64_t warp_mask, int diag_only, int add_noise,
                     double jitter, void* out, void* stream) {
  int rc = check_common(h, kernel_id, 1, d);
  if (rc) return rc;
  if (!X1 || !raw || !out || n1 < 0) return fail(h, HB_ERR_BAD_ARG, "null arg");
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = ensure(h, h->theta, TH_SIZE * 8))) return rc;
  if ((rc = ensure(h, h->bad, 4))) return rc;
  k_prep<<<1, 64, 0, st>>>((const double*)raw, warp_mask, d, 1,
                           (double*)h->theta.p, (unsigned*)h->bad.p, 0);
  HB_LAUNCH_CHECK();
  const bool self = (X2 == nullptr);
  if (self) { X2 = X1; n2 = n1; }
  if (n1 == 0 || n2 == 0) return HB_OK;
  if (self && diag_only) {
    k_fill<<<(unsigned)((n1 + 255) / 256), 256, 0, st>>>(
        (double*)out, n1, (const double*)h->theta.p, TH_SV, 0.0);
    HB_LAUNCH_CHECK();
    return HB_OK;
  }
  dim3 grid((unsigned)((n2 + 63) / 64), (unsigned)((n1 + 63) / 64));
  const size_t smem = 2 * 64 * xstride(d) * 8;
  const int an = (self && add_noise) ? 1 : 0;
#define HB_KM(K)                                                              \
  k_kernel_matrix<K><<<grid, NTHREADS, smem, st>>>(                           \
      (const double*)X1, n1, (const double*)X2, n2, d,                        \
      (const double*)h->theta.p, an, jitter, (double*)out)
  switch (kernel_id) {
    case 0: HB_KM(0); break;
    case 1: HB_KM(1); break;
    default: HB_KM(2); break;
  }
#undef HB_KM
  HB_LAUNCH_CHECK();
  return HB_OK;
}

int hb_factorize_batched(hb_handle_t h, int kernel_id, int mean_id, int T,
                         const int64_t* offs, int d, const void* X,
                         const void* y, const void* raw, uint64_t warp_mask,
                         void* chol_out, void* alpha_out, void* nll_out,
                         int32_t* info_out, void* stream) {
  Plan* p;
  Params P;
  // alpha needs M = L^{-1}; a pure factorisation (alpha_out == NULL) skips it
  return factorize_impl(h, kernel_id, mean_id, T, offs, d, X, y, raw, warp_mask,
                        alpha_out ? 1 : 0, chol_out, alpha_out, nll_out,
                        info_out, (cudaStream_t)stream, &p, &P);
}

int hb_nll_grad_batched(hb_handle_t h, int kernel_id, int mean_id, int T,
                        const int64_t* offs, int d, const void* X, const void* y,
                        const void* raw, uint64_t warp_mask, void* sums_out,
                        void* nll_task_out, int32_t* info_out, void* stream) {
  int rc = check_common(h, kernel_id, mean_id, d);
  if (rc) return rc;
  if (T < 0 || !offs || !raw || !sums_out || (T > 0 && (!X || !y)))
    return fail(h, HB_ERR_BAD_ARG, "null arg");
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = set_attrs(h))) return rc;
  Plan* p;
  if ((rc = get_plan(h, T, offs, d, st, &p))) return rc;
  if ((rc = ensure_ws(h, *p, true))) return rc;
  Params P = make_params(h, *p, kernel_id, mean_id, X, y, 1, true);
  if ((rc = run_factor(h, *p, P, raw, warp_mask, st))) return rc;
  if (T > 0 && p->nblk_max > 0) {
    const int nt = p->nblk_max * (p->nblk_max + 1) / 2;
    const size_t smem = lauum_smem_bytes(d);
    const int lpt = 1;
    dim3 grid((unsigned)nt * T);
    {
      Section sec(h, 2, st);
      switch (kernel_id) {
        case 0: k_lauum_grad<0><<<grid, NTHREADS, smem, st>>>(P, nt, lpt); break;
        case 1: k_lauum_grad<1><<<grid, NTHREADS, smem, st>>>(P, nt, lpt); break;
        default: k_lauum_grad<2><<<grid, NTHREADS, smem, st>>>(P, nt, lpt); break;
      }
      HB_LAUNCH_CHECK();
    }
    Section sec(h, 3, st);
    k_reduce_task<<<T, 64, 0, st>>>(P);
    HB_LAUNCH_CHECK();
  }
  {
    Section sec(h, 3, st);
    k_reduce_final<<<1, 1024, 0, st>>>(P, (double*)sums_out, (double*)nll_task_out);
    HB_LAUNCH_CHECK();
  }
  if (info_out && T > 0) {
    k_copy_info<<<(T + 255) / 256, 256, 0, st>>>(P.info, info_out, T);
    HB_LAUNCH_CHECK();
  }
  return HB_OK;
}

int hb_adam_step(hb_handle_t h, int P_, void* raw, void* m, void* v,
                 void* accepted, const void* sums, void* scalars_io, double lr,
                 double b1, double b2, double eps, int tie_lengthscale,
                 void* stream) {
  if (!h) return HB_ERR_BAD_ARG;
  if (h->dtype != HB_F64) return fail(h, HB_ERR_UNSUPPORTED, "dtype");
  if (P_ < 1 || P_ > 3 + MAX_DIM || !raw || !m || !v || !accepted || !sums ||
      !scalars_io)
    return fail(h, HB_ERR_BAD_ARG, "adam args");
  k_adam<<<1, 64, 0, (cudaStream_t)stream>>>(P_, (double*)raw, (double*)m,
                                            (double*)v, (double*)accepted,
                                            (const double*)sums,
                                            (double*)scalars_io, lr, b1, b2, eps,
                                            tie_lengthscale);
  HB_LAUNCH_CHECK();
  return HB_OK;
}

int64_t hb_predictor_bytes(hb_handle_t h, int64_t n) {
  (void)h;
  if (n < 0) return -1;
  const int64_t nblk = (n + TB - 1) / TB;
  return (nblk * (nblk + 1) / 2 * TILE_ELEMS + nblk * TB) * 8 + 256;
}

int hb_build_predictor(hb_handle_t h, int kernel_id, int mean_id, int64_t n,
                       int d, const void* X, const void* y, const void* raw,
                       uint64_t warp_mask, void* cache, void* chol_out,
                       void* kinvy_out, void* nll_out, int32_t* info_out,
                       void* stream) {
  if (!h) return HB_ERR_BAD_ARG;
  if (n < 1 || !cache) return fail(h, HB_ERR_BAD_ARG, "predictor args");
  const int64_t offs[2] = {0, n};
  cudaStream_t st = (cudaStream_t)stream;
  Plan* p;
  Params P;
  int rc = factorize_impl(h, kernel_id, mean_id, 1, offs, d, X, y, raw,
                          warp_mask, 1, chol_out, kinvy_out, nll_out, info_out,
                          st, &p, &P);
  if (rc) return rc;
  // cache = [M tiles][64-padded alpha]
  const size_t mt_bytes = (size_t)p->total_tiles * TILE_ELEMS * 8;
  const size_t al_bytes = (size_t)p->total_blocks * TB * 8;
  HB_CUDA(cudaMemcpyAsync(cache, P.Mt, mt_bytes, cudaMemcpyDeviceToDevice, st));
  HB_CUDA(cudaMemcpyAsync((char*)cache + mt_bytes, P.alpha, al_bytes,
                          cudaMemcpyDeviceToDevice, st));
  return HB_OK;
}

int hb_predict(hb_handle_t h, int kernel_id, int mean_id, int64_t n, int d,
               const void* X, const void* cache, const void* raw,
               uint64_t warp_mask, int64_t nq, const void* Xq,
               double noise_add_flag, double var_scale, int acq_id,
               double acq_param, void* mu_out, void* var_out, void* acq_out,
               void* stream) {
  int rc = check_common(h, kernel_id, mean_id, d);
  if (rc) return rc;
  if (n < 0 || nq < 0 || !raw || (nq > 0 && !Xq) || (n > 0 && (!X || !cache)))
    return fail(h, HB_ERR_BAD_ARG, "predict args");
  if (acq_id < 0 || acq_id > 3) return fail(h, HB_ERR_BAD_ARG, "acq_id");
  if (acq_out && acq_id == HB_ACQ_NONE) return fail(h, HB_ERR_BAD_ARG, "acq_id");
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = set_attrs(h))) return rc;
  if ((rc = ensure(h, h->theta, TH_SIZE * 8))) return rc;
  if ((rc = ensure(h, h->bad, 4))) return rc;
  k_prep<<<1, 64, 0, st>>>((const double*)raw, warp_mask, d, mean_id,
                           (double*)h->theta.p, (unsigned*)h->bad.p, 0);
  HB_LAUNCH_CHECK();
  if (nq == 0) return HB_OK;
  const int nblk = (int)((n + TB - 1) / TB);
  const int64_t QPASS = 16384;  // queries per pass (bounds the K* scratch)
  const int nqc_max = (int)((std::min(nq, QPASS) + 63) / 64);
  if (nblk > 0) {
    if ((rc = ensure(h, h->kst, (size_t)nqc_max * nblk * TILE_ELEMS * 8))) return rc;
    if ((rc = ensure(h, h->mupart, (size_t)nblk * nqc_max * 64 * 8))) return rc;
    if ((rc = ensure(h, h->vpart, (size_t)nblk * nqc_max * 64 * 8))) return rc;
  }
  const size_t mt_elems = (size_t)nblk * (nblk + 1) / 2 * TILE_ELEMS;
  for (int64_t q0 = 0; q0 < nq; q0 += QPASS) {
    PredParams Q;
    Q.X = (const double*)X;
    Q.Xq = (const double*)Xq + q0 * d;
    Q.theta = (const double*)h->theta.p;
    Q.Mt = (const double*)cache;
    Q.alpha = (const double*)cache + mt_elems;
    Q.kst = (double*)h->kst.p;
    Q.mupart = (double*)h->mupart.p;
    Q.vpart = (double*)h->vpart.p;
    Q.n = (int)n;
    Q.nblk = nblk;
    Q.d = d;
    Q.nq = std::min(QPASS, nq - q0);
    Q.nqc = (int)((Q.nq + 63) / 64);
    if (nblk > 0) {
      const size_t smem = (TILE_ELEMS + 8 * 64 + 64 + 2 * 64 * xstride(d)) * 8;
      switch (kernel_id) {
        case 0: launch_kstar<0>(Q, smem, st); break;
        case 1: launch_kstar<1>(Q, smem, st); break;
        default: launch_kstar<2>(Q, smem, st); break;
      }
      HB_LAUNCH_CHECK();
      k_predict_gemm<<<dim3(nblk, Q.nqc), NTHREADS, PREDICT_SMEM_BYTES, st>>>(Q);
      HB_LAUNCH_CHECK();
    }
    k_predict_final<<<(unsigned)((Q.nq + 255) / 256), 256, 0, st>>>(
        Q, noise_add_flag, var_scale, acq_id, acq_param,
        mu_out ? (double*)mu_out + q0 : nullptr,
        var_out ? (double*)var_out + q0 : nullptr,
        acq_out ? (double*)acq_out + q0 : nullptr);
    HB_LAUNCH_CHECK();
  }
  return HB_OK;
}

int hb_acquisition(hb_handle_t h, int acq_id, double acq_param, int64_t nq,
                   const void* mu, const void* var, void* out, void* stream) {
  if (!h) return HB_ERR_BAD_ARG;
  if (h->dtype != HB_F64) return fail(h, HB_ERR_UNSUPPORTED, "dtype");
  if (acq_id < 1 || acq_id > 3 || nq < 0 || (nq > 0 && (!mu || !var || !out)))
    return fail(h, HB_ERR_BAD_ARG, "acquisition args");
  if (nq == 0) return HB_OK;
  k_acq<<<(unsigned)((nq + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      acq_id, acq_param, nq, (const double*)mu, (const double*)var, (double*)out);
  HB_LAUNCH_CHECK();
  return HB_OK;
}

}  // extern "C"
