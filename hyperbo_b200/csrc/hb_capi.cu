// hyperbo_b200 C ABI (include/hyperbo_b200.h): the extern "C" entry points.
// Host-side planning / workspace code lives in hb_internal.cuh and hb_host.inc,
// the kernels in hb_kernels.inc; both are compiled once per engine precision
// (hb_f64.cu, hb_f32.cu).  Everything is enqueued on the caller's stream.
#include "hb_internal.cuh"

using namespace hb;
using namespace hb::host;

#define HB_IMPL_PROTOTYPES                                                      \
  int kernel_matrix_impl(hb_handle_t, int, const void*, int64_t, const void*,   \
                         int64_t, int, const void*, uint64_t, int, int, double, \
                         void*, void*);                                         \
  int factorize_batched_impl(hb_handle_t, int, int, int, const int64_t*, int,   \
                             const void*, const void*, const void*, uint64_t,   \
                             void*, void*, void*, int32_t*, void*);             \
  int nll_grad_batched_impl(hb_handle_t, int, int, int, const int64_t*, int,    \
                            const void*, const void*, const void*, uint64_t,    \
                            const void*, double, void*, void*, int32_t*, void*);\
  int adam_step_impl(hb_handle_t, int, void*, void*, void*, void*, const void*, \
                     void*, double, double, double, double, int, void*);        \
  int build_predictor_impl(hb_handle_t, int, int, int64_t, int, const void*,    \
                           const void*, const void*, uint64_t, void*, void*,    \
                           void*, void*, int32_t*, void*);                      \
  int predict_impl(hb_handle_t, int, int, int64_t, int, const void*,            \
                   const void*, const void*, uint64_t, int64_t, const void*,    \
                   double, double, int, double, void*, void*, void*, void*);    \
  int acquisition_impl(hb_handle_t, int, double, int64_t, const void*,          \
                       const void*, void*, void*);
namespace hb {
#define HB_COMM_PROTOTYPES                                                      \
  int64_t debug_items_impl(hb_handle_t, int, const int64_t*, int, int,          \
                           int32_t*, int64_t);                                  \
  int nll_grad_multi_impl(hb_handle_t, int, int, int, int, const int64_t*, int, \
                          const void*, const void*, const void*, uint64_t,      \
                          void*, void*, void*);                                 \
  int build_predictors_multi_impl(hb_handle_t, int, int, int, int64_t, int,     \
                                  const void*, const void*, const void*,        \
                                  uint64_t, void*, int64_t, void*, int32_t*,    \
                                  void*);                                       \
  int subsample_impl(hb_handle_t, int, int, const void*, const void*,           \
                     const void*, int64_t, const void*, const void*, void*,     \
                     void*, uint64_t, const void*, int64_t, void*);             \
  int bo_init_impl(hb_handle_t, int, int, int64_t, int64_t, int, const void*,   \
                   const void*, const void*, uint64_t, void*, void*);           \
  int bo_step_impl(hb_handle_t, int, int, int64_t, int64_t, int, void*, void*,  \
                   const void*, uint64_t, void*, int64_t, const void*,          \
                   const void*, double, double, int, double, int, int32_t*,     \
                   void*);                                                      \
  int fused_timeout_impl();                                                     \
  int predict_cov_impl(hb_handle_t, int, int, int64_t, int, const void*,        \
                       const void*, const void*, uint64_t, int64_t, const void*,\
                       double, double, void*, void*, void*);                    \
  int euclid_grad_impl(hb_handle_t, int, int, int, const int64_t*, int,         \
                       const void*, int, const void*, const void*, const void*, \
                       uint64_t, double, double, const void*, void*, void*);    \
  int nll_grad_mrhs_impl(hb_handle_t, int, int, int, const int64_t*, int,       \
                         const void*, int, const void*, const void*,            \
                         const int32_t*, const void*, uint64_t, const void*,    \
                         double, void*, int32_t*, void*);                       \
  int allreduce_impl(hb_handle_t, void*, int, void*);                           \
  int allreduce_adam_impl(hb_handle_t, int, void*, void*, void*, void*, void*,  \
                          void*, double, double, double, double, int, void*);
namespace f64 { HB_IMPL_PROTOTYPES HB_COMM_PROTOTYPES }
namespace f32 { HB_IMPL_PROTOTYPES HB_COMM_PROTOTYPES }
}  // namespace hb
#undef HB_IMPL_PROTOTYPES

// Every compute entry point: serialise on the handle (its plan cache, workspace
// and error string are shared state) and run on the handle's device whatever
// device is current in the calling thread.
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev); else prev = -1;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define HB_DISPATCH(fn, ...)                                          \
  do {                                                                \
    if (!h) return HB_ERR_BAD_ARG;                                    \
    std::lock_guard<std::recursive_mutex> lock_(h->mu);               \
    DeviceGuard guard_(h->device);                                    \
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
  if (const char* e = getenv("HB_FUSED")) h->fused = std::max(0, std::min(2, atoi(e)));
  if (const char* e = getenv("HB_FUSED_GRID")) h->fused_grid = atoi(e);
  if (const char* e = getenv("HB_FUSED_SKEW")) h->fused_skew = atof(e);
  if (const char* e = getenv("HB_FUSED_GROUPS")) h->fused_groups = atoi(e);
  if (const char* e = getenv("HB_FUSED_FASTPATH")) h->fused_fastpath = atoi(e) ? 1 : 0;
  if (const char* e = getenv("HB_FUSED_VT_DIAG")) h->fused_vt_diag = atof(e);
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

static void comm_release(hb_handle_t h) {
  for (int r = 0; r < (int)h->xbuf_peers.size(); ++r)
    if (h->xbuf_peers[r] && r != h->comm_rank) cudaIpcCloseMemHandle(h->xbuf_peers[r]);
  h->xbuf_peers.clear();
  if (h->xbuf_peers_d) cudaFree(h->xbuf_peers_d);
  if (h->comm_step_d) cudaFree(h->comm_step_d);
  h->xbuf_peers_d = nullptr;
  h->comm_step_d = nullptr;
  h->comm_world = 0;
  h->comm_rank = -1;
}

int hb_destroy(hb_handle_t h) {
  if (!h) return HB_ERR_BAD_ARG;
  comm_release(h);
  if (h->xbuf_own) cudaFree(h->xbuf_own);
  Buf* all[] = {&h->theta, &h->Lt,    &h->Mt,  &h->Wt, &h->zz,  &h->z,    &h->alpha,
                &h->logdet, &h->asum, &h->nll_task, &h->gpart, &h->gtask,
                &h->info,  &h->bad,   &h->sums,  &h->kst,  &h->mupart,
                &h->vpart, &h->pcache, &h->pre, &h->sync, &h->apart,
                &h->mrz,   &h->mra,    &h->zeros, &h->vt};
  for (auto* b : all)
    if (b->p) cudaFree(b->p);
  for (auto& p : h->plans) {
    if (p.tasks_d) cudaFree(p.tasks_d);
    for (void* it : p.items_d)
      if (it) cudaFree(it);
  }
  delete h;
  return HB_OK;
}

const char* hb_last_error(hb_handle_t h) { return h ? h->err.c_str() : "null handle"; }
int64_t hb_launch_count(hb_handle_t h) { return h ? h->launches : 0; }
int64_t hb_workspace_bytes(hb_handle_t h) { return h ? (int64_t)total_ws(h) : 0; }
int64_t hb_generation(hb_handle_t h) { return h ? (int64_t)h->generation : -1; }
// debug: copy an internal workspace buffer to the host (tests / triage only)
int64_t hb_debug_read(hb_handle_t h, int which, void* host_out, int64_t max_bytes) {
  if (!h) return -1;
  hb::host::Buf* bufs[] = {&h->Lt, &h->Mt, &h->Wt, &h->z, &h->alpha, &h->apart,
                           &h->gpart, &h->gtask, &h->logdet, &h->nll_task, &h->sync};
  if (which < 0 || which >= (int)(sizeof(bufs) / sizeof(bufs[0]))) return -1;
  cudaDeviceSynchronize();
  const int64_t nb = std::min<int64_t>(max_bytes, (int64_t)bufs[which]->cap);
  if (host_out && nb > 0 &&
      cudaMemcpy(host_out, bufs[which]->p, nb, cudaMemcpyDeviceToHost) != cudaSuccess)
    return -1;
  return (int64_t)bufs[which]->cap;
}
// tuning / test knobs of a handle (the HB_* environment variables set the
// defaults at hb_create): "fused" 0|1|2, "fastpath" 0|1, "groups", "skew", "grid"
int hb_set_option(hb_handle_t h, const char* name, double value) {
  if (!h || !name) return HB_ERR_BAD_ARG;
  const std::string n(name);
  if (n == "fused") h->fused = std::max(0, std::min(2, (int)value));
  else if (n == "fastpath") h->fused_fastpath = value != 0.0;
  else if (n == "groups") h->fused_groups = (int)value;
  else if (n == "skew") h->fused_skew = value;
  else if (n == "grid") h->fused_grid = (int)value;
  else return fail(h, HB_ERR_BAD_ARG, "unknown option");
  for (auto& p : h->plans)  // item lists depend on these
    for (int& k : p.nitems) k = -1;
  ++h->generation;
  return HB_OK;
}
int64_t hb_debug_items(hb_handle_t h, int T, const int64_t* offs, int d, int variant,
                       int32_t* out, int64_t max_items) {
  if (!h) return -1;
  return h->dtype == HB_F64 ? hb::f64::debug_items_impl(h, T, offs, d, variant, out, max_items)
                            : hb::f32::debug_items_impl(h, T, offs, d, variant, out, max_items);
}
int hb_debug_fused_timeout(hb_handle_t h) {
  if (!h) return -1;
  return h->dtype == HB_F64 ? hb::f64::fused_timeout_impl() : hb::f32::fused_timeout_impl();
}

#ifdef HB_STAMPS
extern "C++" {
namespace hb { namespace f64 { int dbg_potrf64_impl(int, int, long long*); }
               namespace f32 { int dbg_potrf64_impl(int, int, long long*); } }
}
int hb_debug_potrf64(int f32, int grid, int reps, long long* host_out) {
  return f32 ? hb::f32::dbg_potrf64_impl(grid, reps, host_out)
             : hb::f64::dbg_potrf64_impl(grid, reps, host_out);
}
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

// ---- peer-memory all-reduce (SURVEY 8b hb_allreduce, 8e) -------------------
int hb_comm_export(hb_handle_t h, void* ipc_handle_out) {
  if (!h || !ipc_handle_out) return HB_ERR_BAD_ARG;
  HB_CUDA(cudaSetDevice(h->device));
  if (!h->xbuf_own) {
    HB_CUDA(cudaMalloc(&h->xbuf_own, HB_XBUF_BYTES));
    HB_CUDA(cudaMemset(h->xbuf_own, 0, HB_XBUF_BYTES));
    HB_CUDA(cudaDeviceSynchronize());
  }
  cudaIpcMemHandle_t hd;
  HB_CUDA(cudaIpcGetMemHandle(&hd, h->xbuf_own));
  static_assert(sizeof(cudaIpcMemHandle_t) == HB_IPC_HANDLE_BYTES, "ipc handle size");
  std::memcpy(ipc_handle_out, &hd, sizeof(hd));
  return HB_OK;
}

int hb_comm_import(hb_handle_t h, int rank, int world, const void* ipc_handles) {
  if (!h || !ipc_handles || world < 1 || world > HB_XCHG_MAX || rank < 0 || rank >= world)
    return fail(h, HB_ERR_BAD_ARG, "comm args");
  if (!h->xbuf_own) return fail(h, HB_ERR_BAD_ARG, "hb_comm_export first");
  HB_CUDA(cudaSetDevice(h->device));
  comm_release(h);
  h->comm_rank = rank;
  h->xbuf_peers.assign(world, nullptr);
  for (int r = 0; r < world; ++r) {
    if (r == rank) {
      h->xbuf_peers[r] = h->xbuf_own;
      continue;
    }
    cudaIpcMemHandle_t hd;
    std::memcpy(&hd, (const char*)ipc_handles + (size_t)r * sizeof(hd), sizeof(hd));
    HB_CUDA(cudaIpcOpenMemHandle(&h->xbuf_peers[r], hd, cudaIpcMemLazyEnablePeerAccess));
  }
  HB_CUDA(cudaMemset(h->xbuf_own, 0, HB_XBUF_BYTES));
  HB_CUDA(cudaMalloc((void**)&h->xbuf_peers_d, sizeof(void*) * world));
  HB_CUDA(cudaMemcpy(h->xbuf_peers_d, h->xbuf_peers.data(), sizeof(void*) * world,
                     cudaMemcpyHostToDevice));
  HB_CUDA(cudaMalloc((void**)&h->comm_step_d, sizeof(unsigned)));
  HB_CUDA(cudaMemset(h->comm_step_d, 0, sizeof(unsigned)));
  HB_CUDA(cudaDeviceSynchronize());
  h->comm_world = world;
  ++h->generation;
  return HB_OK;
}

int hb_allreduce(hb_handle_t h, void* buf, int count, void* stream) {
  HB_DISPATCH(allreduce_impl, h, buf, count, stream);
}

int hb_allreduce_adam_step(hb_handle_t h, int P_, void* raw, void* m, void* v,
                           void* accepted, void* sums, void* scalars_io, double lr,
                           double b1, double b2, double eps, int tie_lengthscale,
                           void* stream) {
  HB_DISPATCH(allreduce_adam_impl, h, P_, raw, m, v, accepted, sums, scalars_io, lr,
              b1, b2, eps, tie_lengthscale, stream);
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
  HB_DISPATCH(factorize_batched_impl, h, kernel_id, mean_id, T, offs, d, X, y,
              raw, warp_mask, chol_out, alpha_out, nll_out, info_out, stream);
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

int hb_nll_grad_mrhs(hb_handle_t h, int kernel_id, int mean_id, int T,
                     const int64_t* offs, int d, const void* X, int R,
                     const void* B, const void* col_weight, const int32_t* col_mean,
                     const void* raw, uint64_t warp_mask, const void* task_weight,
                     double jitter, void* sums_out, int32_t* info_out, void* stream) {
  HB_DISPATCH(nll_grad_mrhs_impl, h, kernel_id, mean_id, T, offs, d, X, R, B,
              col_weight, col_mean, raw, warp_mask, task_weight, jitter, sums_out,
              info_out, stream);
}

int hb_euclid_grad(hb_handle_t h, int kernel_id, int mean_id, int T,
                   const int64_t* offs, int d, const void* X, int R, const void* Yc,
                   const void* mu0, const void* raw, uint64_t warp_mask,
                   double mean_weight, double cov_weight, const void* task_weight,
                   void* sums_out, void* stream) {
  HB_DISPATCH(euclid_grad_impl, h, kernel_id, mean_id, T, offs, d, X, R, Yc, mu0, raw,
              warp_mask, mean_weight, cov_weight, task_weight, sums_out, stream);
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

int hb_predict_cov(hb_handle_t h, int kernel_id, int mean_id, int64_t n, int d,
                   const void* X, const void* cache, const void* raw,
                   uint64_t warp_mask, int64_t nq, const void* Xq,
                   double noise_add_flag, double var_scale, void* mu_out,
                   void* cov_out, void* stream) {
  HB_DISPATCH(predict_cov_impl, h, kernel_id, mean_id, n, d, X, cache, raw, warp_mask,
              nq, Xq, noise_add_flag, var_scale, mu_out, cov_out, stream);
}

int hb_nll_grad_multi(hb_handle_t h, int kernel_id, int mean_id, int S, int T,
                      const int64_t* offs, int d, const void* X, const void* y,
                      const void* raw_sets, uint64_t warp_mask, void* sums_out,
                      void* nll_task_out, void* stream) {
  HB_DISPATCH(nll_grad_multi_impl, h, kernel_id, mean_id, S, T, offs, d, X, y, raw_sets,
              warp_mask, sums_out, nll_task_out, stream);
}

int hb_build_predictors_multi(hb_handle_t h, int kernel_id, int mean_id, int S, int64_t n,
                              int d, const void* X, const void* y, const void* raw_sets,
                              uint64_t warp_mask, void* caches,
                              int64_t cache_stride_bytes, void* nll_out,
                              int32_t* info_out, void* stream) {
  HB_DISPATCH(build_predictors_multi_impl, h, kernel_id, mean_id, S, n, d, X, y, raw_sets,
              warp_mask, caches, cache_stride_bytes, nll_out, info_out, stream);
}

int hb_subsample(hb_handle_t h, int T, int d, const void* offs_src_dev,
                 const void* offs_dst_dev, const void* task_ids_dev, int64_t max_rows,
                 const void* Xs, const void* ys, void* Xd, void* yd, uint64_t seed,
                 const void* step_scalars_dev, int64_t step, void* stream) {
  HB_DISPATCH(subsample_impl, h, T, d, offs_src_dev, offs_dst_dev, task_ids_dev, max_rows,
              Xs, ys, Xd, yd, seed, step_scalars_dev, step, stream);
}

// the permutation hb_subsample uses, for host-side checks (tests)
uint32_t hb_subsample_perm(uint32_t i, uint32_t n, uint64_t seed, uint64_t step,
                           int64_t task_id) {
  const unsigned long long key = hb::hb_mix64(
      hb::hb_mix64(seed ^ 0x5851f42d4c957f2dULL) +
      hb::hb_mix64(step) * 0x2545f4914f6cdd1dULL + (unsigned long long)task_id);
  return hb::hb_feistel_perm(i, n, key);
}

int64_t hb_bo_cache_bytes(hb_handle_t h, int64_t n_cap) {
  if (n_cap < 1) return -1;
  const int64_t es = (h && h->dtype == HB_F32) ? 4 : 8;
  const int64_t cb = (n_cap + TB - 1) / TB;
  return (cb * (cb + 1) / 2 * TILE_ELEMS + cb * TB + 64) * es + 256;
}

int hb_bo_init(hb_handle_t h, int kernel_id, int mean_id, int64_t n0, int64_t n_cap,
               int d, const void* X, const void* y, const void* raw,
               uint64_t warp_mask, void* cache, void* stream) {
  HB_DISPATCH(bo_init_impl, h, kernel_id, mean_id, n0, n_cap, d, X, y, raw, warp_mask,
              cache, stream);
}

int hb_bo_step(hb_handle_t h, int kernel_id, int mean_id, int64_t n, int64_t n_cap,
               int d, void* X, void* y, const void* raw, uint64_t warp_mask,
               void* cache, int64_t nq, const void* Xq, const void* yq,
               double noise_add_flag, double var_scale, int acq_id, double acq_param,
               int target_is_ymax, int32_t* sel_out, void* stream) {
  HB_DISPATCH(bo_step_impl, h, kernel_id, mean_id, n, n_cap, d, X, y, raw, warp_mask,
              cache, nq, Xq, yq, noise_add_flag, var_scale, acq_id, acq_param,
              target_is_ymax, sel_out, stream);
}

int hb_acquisition(hb_handle_t h, int acq_id, double acq_param, int64_t nq,
                   const void* mu, const void* var, void* out, void* stream) {
  HB_DISPATCH(acquisition_impl, h, acq_id, acq_param, nq, mu, var, out, stream);
}

}  // extern "C"
