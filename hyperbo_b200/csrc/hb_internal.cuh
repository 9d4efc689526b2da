// hyperbo_b200 host-side internals shared by the C-ABI translation units:
// the handle (workspace arena, plan cache), error helpers and per-section
// event profiling.  The kernels and their launch code are compiled once per
// engine precision (hb_f64.cu / hb_f32.cu, in parallel); hb_capi.cu holds the
// extern "C" entry points and dispatches on the handle's dtype.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/hyperbo_b200.h"
#include "hb_common.cuh"

namespace hb {
namespace host {

struct Buf {
  void* p = nullptr;
  size_t cap = 0;
};

struct Plan {
  std::vector<int64_t> offs;
  int T = 0, d = 0;   // T = virtual tasks = S * (tasks of the batch)
  int S = 1;          // hyper-parameter sets (second batch axis)
  std::vector<TaskDesc> tasks;
  TaskDesc* tasks_d = nullptr;
  size_t tasks_cap = 0;
  int nblk_max = 0;
  long long total_tiles = 0, total_blocks = 0, sum_n = 0, chol_elems = 0;
  uint64_t stamp = 0;
  bool uploaded = false;
  // work-item lists of the persistent kernel (hb_fused.inc), built lazily per
  // variant: 0 = factor only, 1 = + triangular inverse, 2 = + K~^{-1}/gradient
  void* items_d[3] = {nullptr, nullptr, nullptr};
  size_t items_cap[3] = {0, 0, 0};
  int nitems[3] = {-1, -1, -1};
  int ngroups[3] = {1, 1, 1};
  size_t gofs_off[3] = {0, 0, 0};
};

constexpr int NPLAN = 4;

}  // namespace host
}  // namespace hb

struct hb_handle_s {
  std::recursive_mutex mu;  // entry points serialise on the handle
  int device = 0;
  int dtype = HB_F64;
  std::string err;
  int64_t launches = 0;
  uint64_t clock = 0;
  hb::host::Plan plans[hb::host::NPLAN];
  hb::host::Buf theta, Lt, Mt, Wt, zz, z, alpha, logdet, asum, nll_task, gpart, gtask, info, bad,
      sums, kst, mupart, vpart, pcache, stamps, pre, sync, apart, mrz, mra, zeros, vt;
  bool attr_set = false;
  bool bo_attr_set = false;
  bool mrhs_attr_set = false;
  bool cov_attr_set = false;
  int sm_count = 0;
  int fused = 1;               // HB_FUSED env: 0 = launch-per-column path, 2 = always persistent
  int fused_grid = 0;          // HB_FUSED_GRID env: CTAs of the persistent kernel
  double fused_skew = 0.0;     // HB_FUSED_SKEW env: task skew of the item order
  int fused_fastpath = 1;      // HB_FUSED_FASTPATH env: 0 = acquire-poll every dependency
  int fused_groups = 0;        // HB_FUSED_GROUPS env: work queues (0 = automatic)
  double fused_vt_diag = 0.4;  // HB_FUSED_VT_DIAG env: queue position of DIAG(j+1)
  int fused_per_sm[2][hb::MAX_DIM + 1] = {};  // cached occupancy per (mapping, d)
  uint64_t generation = 0;     // bumped whenever a workspace buffer or plan moves
  int pre_override = -1;       // HB_PRE env: force the k_step pre roles off / on
  long long pre_cta_limit = 0; // pre roles on when T * (nblk_max + 1) <= this
  int smem_d = -1;
  // peer-memory all-reduce (hb_comm_*): own exchange buffer, the ranks' buffers
  // mapped with CUDA IPC, device array of their addresses, device step counter
  int comm_rank = -1, comm_world = 0;
  void* xbuf_own = nullptr;
  std::vector<void*> xbuf_peers;     // [world]; entry `rank` == xbuf_own
  void** xbuf_peers_d = nullptr;
  unsigned* comm_step_d = nullptr;
  bool profiling = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_ev[HB_PROFILE_SECTIONS];
};

namespace hb {
namespace host {

#define HB_CUDA(call)                                                         \
  do {                                                                        \
    cudaError_t e_ = (call);                                                  \
    if (e_ != cudaSuccess) {                                                  \
      h->err = std::string(#call) + ": " + cudaGetErrorString(e_);            \
      return HB_ERR_CUDA;                                                     \
    }                                                                         \
  } while (0)

inline int fail(hb_handle_t h, int code, const char* msg) {
  if (h) h->err = msg;
  return code;
}

inline int ensure(hb_handle_t h, Buf& b, size_t bytes) {
  if (bytes <= b.cap) return HB_OK;
  ++h->generation;  // captured CUDA graphs hold the old pointer
  if (b.p) HB_CUDA(cudaFree(b.p));
  b.p = nullptr;
  b.cap = 0;
  size_t want = bytes + bytes / 8 + 256;
  HB_CUDA(cudaMalloc(&b.p, want));
  b.cap = want;
  return HB_OK;
}

inline size_t total_ws(hb_handle_t h) {
  const Buf* all[] = {&h->theta, &h->Lt,    &h->Mt,  &h->Wt, &h->zz,  &h->z,    &h->alpha,
                      &h->logdet, &h->asum, &h->nll_task, &h->gpart, &h->gtask,
                      &h->info,  &h->bad,   &h->sums,  &h->kst,  &h->mupart,
                      &h->vpart, &h->pcache, &h->pre, &h->sync, &h->apart,
                      &h->mrz,   &h->mra,    &h->zeros, &h->vt};
  size_t s = 0;
  for (auto* b : all) s += b->cap;
  for (auto& p : h->plans) {
    s += p.tasks_cap;
    for (size_t c : p.items_cap) s += c;
  }
  return s;
}

// find or build the plan for (T, offs, d); uploads descriptors when new
// S > 1: S hyper-parameter sets x the same Tb = T tasks: virtual task s * Tb + t
// reads the data rows of task t and the parameter set s, and owns its own
// workspace tiles
inline int get_plan(hb_handle_t h, int Tb, const int64_t* offs, int d, cudaStream_t st,
             Plan** out, int S = 1) {
  ++h->clock;
  const int T = Tb * S;
  Plan* lru = &h->plans[0];
  for (auto& p : h->plans) {
    if (p.uploaded && p.T == T && p.S == S && p.d == d && (int)p.offs.size() == Tb + 1 &&
        std::memcmp(p.offs.data(), offs, sizeof(int64_t) * (Tb + 1)) == 0) {
      p.stamp = h->clock;
      *out = &p;
      return HB_OK;
    }
    if (p.stamp < lru->stamp) lru = &p;
  }
  Plan& p = *lru;
  if (p.uploaded) ++h->generation;  // an evicted plan may be baked into a graph
  p.uploaded = false;
  for (int& n : p.nitems) n = -1;
  p.T = T;
  p.S = S;
  p.d = d;
  p.offs.assign(offs, offs + Tb + 1);
  p.tasks.resize(T);
  p.nblk_max = 0;
  long long tiles = 0, blocks = 0, chol = 0;
  for (int t = 0; t < T; ++t) {
    const int tb = t % std::max(Tb, 1);
    const int64_t n = offs[tb + 1] - offs[tb];
    if (n < 0 || n > (1 << 20)) return fail(h, HB_ERR_BAD_ARG, "bad offs");
    TaskDesc& td = p.tasks[t];
    td.n = (int)n;
    td.nblk = (int)((n + TB - 1) / TB);
    td.xoff = offs[tb];
    td.theta_idx = Tb > 0 ? t / Tb : 0;
    td.pad_ = 0;
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
  p.sum_n = (offs[Tb] - offs[0]) * S;
  p.chol_elems = chol;
  const size_t bytes = sizeof(TaskDesc) * (size_t)std::max(T, 1);
  if (bytes > p.tasks_cap) {
    ++h->generation;
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

inline int check_common(hb_handle_t h, int kernel_id, int mean_id, int d) {
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

}  // namespace host
}  // namespace hb
