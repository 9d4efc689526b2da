// hyperbo_b200 C ABI (include/hyperbo_b200.h): host-side planning, workspace
// arena and kernel launches.  Everything is enqueued on the caller's stream.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/hyperbo_b200.h"
#include "hb_common.cuh"

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
      sums, kst, mupart, vpart, pcache, stamps, pre;
  bool attr_set = false;
  int pre_override = -1;       // HB_PRE env: force the k_step pre roles off / on
  long long pre_cta_limit = 0; // pre roles on when T * (nblk_max + 1) <= this
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
                      &h->vpart, &h->pcache, &h->pre};
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

int check_common(hb_handle_t h, int kernel_id, int mean_id, int d) {
  if (!h) return HB_ERR_BAD_ARG;
  if (kernel_id < 0 || kernel_id > 2) return fail(h, HB_ERR_BAD_ARG, "kernel_id");
  if (mean_id < 0 || mean_id > 1) return fail(h, HB_ERR_BAD_ARG, "mean_id");
  if (d < 1) return fail(h, HB_ERR_BAD_ARG, "d < 1");
  if (d > MAX_DIM) return fail(h, HB_ERR_UNSUPPORTED, "d > HB_MAX_DIM");
  return HB_OK;
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

}  // namespace

// ------------------------------------------------------------------------
// The kernels and their launch code exist once per engine precision: the
// .inc files are compiled twice with Real = double (hb::f64, DMMA tile
// products) and Real = float (hb::f32, 3xTF32 tile products).
#define HB_F64 1
#define HB_MIN_CTAS 2
namespace hb { namespace f64 {
using Real = double;
using Real2 = double2;
__device__ __forceinline__ Real2 make_real2(Real a, Real b) { return make_double2(a, b); }
#include "hb_device.inc"
#include "hb_kernels.inc"
#include "hb_host.inc"
} }  // namespace hb::f64
#undef HB_F64
#undef HB_MIN_CTAS
#define HB_F64 0
#define HB_MIN_CTAS 3
namespace hb { namespace f32 {
using Real = float;
using Real2 = float2;
__device__ __forceinline__ Real2 make_real2(Real a, Real b) { return make_float2(a, b); }
#include "hb_device.inc"
#include "hb_kernels.inc"
#include "hb_host.inc"
} }  // namespace hb::f32
#undef HB_F64

#define HB_DISPATCH(fn, ...)                                          \
  do {                                                                \
    if (!h) return HB_ERR_BAD_ARG;                                    \
    return h->dtype == HB_F64 ? hb::f64::fn(__VA_ARGS__)              \
                              : hb::f32::fn(__VA_ARGS__);             \
  } while (0)

extern "C" {

const char* hb_version(void) {
  return "hyperbo_b200 0.2 (sm_100a; fp64 DMMA and fp32 3xTF32 engines)";
}

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
  if (const char* e = getenv("HB_PRE")) h->pre_override = atoi(e) ? 1 : 0;
  {
    cudaDeviceProp prop;
    // CTA slots of one wave (2 resident CTAs per SM in fp64, 3 in fp32)
    h->pre_cta_limit = cudaGetDeviceProperties(&prop, device) == cudaSuccess
                           ? 2LL * prop.multiProcessorCount * (dtype == HB_F64 ? 2 : 3)
                           : 592;
    if (const char* e = getenv("HB_PRE_LIMIT")) h->pre_cta_limit = atoll(e);
  }
  *out = h;
  return HB_OK;
}

int hb_destroy(hb_handle_t h) {
  if (!h) return HB_ERR_BAD_ARG;
  Buf* all[] = {&h->theta, &h->Lt,    &h->Mt,  &h->Wt, &h->zz,  &h->z,    &h->alpha,
                &h->logdet, &h->asum, &h->nll_task, &h->gpart, &h->gtask,
                &h->info,  &h->bad,   &h->sums,  &h->kst,  &h->mupart,
                &h->vpart, &h->pcache, &h->pre};
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
  HB_DISPATCH(kernel_matrix_impl, h, kernel_id, X1, n1, X2, n2, d, raw,
              warp_mask, diag_only, add_noise, jitter, out, stream);
}

int hb_factorize_batched(hb_handle_t h, int kernel_id, int mean_id, int T,
                         const int64_t* offs, int d, const void* X,
                         const void* y, const void* raw, uint64_t warp_mask,
                         void* chol_out, void* alpha_out, void* nll_out,
                         int32_t* info_out, void* stream) {
  Plan* p = nullptr;
  // alpha needs M = L^{-1}; a pure factorisation (alpha_out == NULL) skips it
  if (!h) return HB_ERR_BAD_ARG;
  if (h->dtype == HB_F64) {
    hb::f64::Params P;
    return hb::f64::factorize_impl(h, kernel_id, mean_id, T, offs, d, X, y, raw,
                                   warp_mask, alpha_out ? 1 : 0, chol_out,
                                   alpha_out, nll_out, info_out,
                                   (cudaStream_t)stream, &p, &P);
  }
  hb::f32::Params P;
  return hb::f32::factorize_impl(h, kernel_id, mean_id, T, offs, d, X, y, raw,
                                 warp_mask, alpha_out ? 1 : 0, chol_out,
                                 alpha_out, nll_out, info_out,
                                 (cudaStream_t)stream, &p, &P);
}

int hb_nll_grad_batched(hb_handle_t h, int kernel_id, int mean_id, int T,
                        const int64_t* offs, int d, const void* X, const void* y,
                        const void* raw, uint64_t warp_mask, void* sums_out,
                        void* nll_task_out, int32_t* info_out, void* stream) {
  HB_DISPATCH(nll_grad_batched_impl, h, kernel_id, mean_id, T, offs, d, X, y,
              raw, warp_mask, nullptr, hb::JITTER, sums_out, nll_task_out,
              info_out, stream);
}

int hb_nll_grad_weighted(hb_handle_t h, int kernel_id, int mean_id, int T,
                         const int64_t* offs, int d, const void* X,
                         const void* y, const void* raw, uint64_t warp_mask,
                         const void* task_weight, double jitter, void* sums_out,
                         void* nll_task_out, int32_t* info_out, void* stream) {
  HB_DISPATCH(nll_grad_batched_impl, h, kernel_id, mean_id, T, offs, d, X, y,
              raw, warp_mask, task_weight, jitter, sums_out, nll_task_out,
              info_out, stream);
}

int hb_adam_step(hb_handle_t h, int P_, void* raw, void* m, void* v,
                 void* accepted, const void* sums, void* scalars_io, double lr,
                 double b1, double b2, double eps, int tie_lengthscale,
                 void* stream) {
  HB_DISPATCH(adam_step_impl, h, P_, raw, m, v, accepted, sums, scalars_io, lr,
              b1, b2, eps, tie_lengthscale, stream);
}

int64_t hb_predictor_bytes(hb_handle_t h, int64_t n) {
  if (n < 0) return -1;
  const int64_t es = (h && h->dtype == HB_F32) ? 4 : 8;
  const int64_t nblk = (n + TB - 1) / TB;
  return (nblk * (nblk + 1) / 2 * TILE_ELEMS + nblk * TB) * es + 256;
}

int hb_build_predictor(hb_handle_t h, int kernel_id, int mean_id, int64_t n,
                       int d, const void* X, const void* y, const void* raw,
                       uint64_t warp_mask, void* cache, void* chol_out,
                       void* kinvy_out, void* nll_out, int32_t* info_out,
                       void* stream) {
  HB_DISPATCH(build_predictor_impl, h, kernel_id, mean_id, n, d, X, y, raw,
              warp_mask, cache, chol_out, kinvy_out, nll_out, info_out, stream);
}

int hb_predict(hb_handle_t h, int kernel_id, int mean_id, int64_t n, int d,
               const void* X, const void* cache, const void* raw,
               uint64_t warp_mask, int64_t nq, const void* Xq,
               double noise_add_flag, double var_scale, int acq_id,
               double acq_param, void* mu_out, void* var_out, void* acq_out,
               void* stream) {
  HB_DISPATCH(predict_impl, h, kernel_id, mean_id, n, d, X, cache, raw,
              warp_mask, nq, Xq, noise_add_flag, var_scale, acq_id, acq_param,
              mu_out, var_out, acq_out, stream);
}

int hb_acquisition(hb_handle_t h, int acq_id, double acq_param, int64_t nq,
                   const void* mu, const void* var, void* out, void* stream) {
  HB_DISPATCH(acquisition_impl, h, acq_id, acq_param, nq, mu, var, out, stream);
}

}  // extern "C"
