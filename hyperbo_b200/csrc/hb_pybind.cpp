// Thin pybind11 layer over the C ABI (include/hyperbo_b200.h): forwards raw
// device pointers (as integers), sizes and the CUDA stream.  No torch types, no
// arithmetic -- the product is the C-ABI library this links against.
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/hyperbo_b200.h"

namespace py = pybind11;
using ptr_t = std::uintptr_t;

namespace {

inline void* P(ptr_t p) { return reinterpret_cast<void*>(p); }

struct Handle {
  hb_handle_t h = nullptr;
  Handle(int device, int dtype) {
    const int rc = hb_create(&h, device, dtype);
    if (rc != HB_OK)
      throw std::runtime_error("hb_create failed with status " +
                               std::to_string(rc) +
                               (rc == HB_ERR_NO_DEVICE ? " (no CUDA device)" : ""));
  }
  ~Handle() {
    if (h) hb_destroy(h);
  }
  void check(int rc, const char* what) const {
    if (rc == HB_OK) return;
    const std::string msg = std::string(what) + ": status " + std::to_string(rc) +
                            " (" + hb_last_error(h) + ")";
    if (rc == HB_ERR_UNSUPPORTED) throw py::value_error(msg);  // -> mapped in Python
    throw std::runtime_error(msg);
  }
};

}  // namespace

PYBIND11_MODULE(_C, m) {
  m.def("subsample_perm", [](uint32_t i, uint32_t n, uint64_t seed, uint64_t step,
                             int64_t task) { return hb_subsample_perm(i, n, seed, step, task); });
  m.doc() = "hyperbo_b200 C-ABI bindings";
  m.def("version", [] { return std::string(hb_version()); });
  m.attr("MAX_DIM") = HB_MAX_DIM;
  m.attr("TILE") = HB_TILE;

  py::class_<Handle>(m, "Handle")
      .def(py::init<int, int>(), py::arg("device"), py::arg("dtype"))
      .def("launch_count", [](Handle& s) { return hb_launch_count(s.h); })
      .def("workspace_bytes", [](Handle& s) { return hb_workspace_bytes(s.h); })
      .def("profile_enable",
           [](Handle& s, bool on) {
             s.check(hb_profile_enable(s.h, on ? 1 : 0), "hb_profile_enable");
           })
      .def("profile_read",
           [](Handle& s) {
             std::vector<double> ms(HB_PROFILE_SECTIONS);
             std::vector<int64_t> cnt(HB_PROFILE_SECTIONS);
             s.check(hb_profile_read(s.h, ms.data(), cnt.data()),
                     "hb_profile_read");
             return std::make_pair(ms, cnt);
           })
      .def("predictor_bytes",
           [](Handle& s, int64_t n) { return hb_predictor_bytes(s.h, n); })
      .def("kernel_matrix",
           [](Handle& s, int kernel_id, ptr_t X1, int64_t n1, ptr_t X2, int64_t n2,
              int d, ptr_t raw, uint64_t mask, int diag_only, int add_noise,
              double jitter, ptr_t out, ptr_t stream) {
             py::gil_scoped_release rel;
             s.check(hb_kernel_matrix(s.h, kernel_id, P(X1), n1, P(X2), n2, d,
                                      P(raw), mask, diag_only, add_noise, jitter,
                                      P(out), P(stream)),
                     "hb_kernel_matrix");
           })
      .def("factorize_batched",
           [](Handle& s, int kernel_id, int mean_id, std::vector<int64_t> offs,
              int d, ptr_t X, ptr_t y, ptr_t raw, uint64_t mask, ptr_t chol,
              ptr_t alpha, ptr_t nll, ptr_t info, ptr_t stream) {
             py::gil_scoped_release rel;
             s.check(hb_factorize_batched(
                         s.h, kernel_id, mean_id, (int)offs.size() - 1,
                         offs.data(), d, P(X), P(y), P(raw), mask, P(chol),
                         P(alpha), P(nll), (int32_t*)P(info), P(stream)),
                     "hb_factorize_batched");
           })
      .def("nll_grad_batched",
           [](Handle& s, int kernel_id, int mean_id, std::vector<int64_t> offs,
              int d, ptr_t X, ptr_t y, ptr_t raw, uint64_t mask, ptr_t sums,
              ptr_t nll_task, ptr_t info, ptr_t stream) {
             py::gil_scoped_release rel;
             s.check(hb_nll_grad_batched(
                         s.h, kernel_id, mean_id, (int)offs.size() - 1,
                         offs.data(), d, P(X), P(y), P(raw), mask, P(sums),
                         P(nll_task), (int32_t*)P(info), P(stream)),
                     "hb_nll_grad_batched");
           })
      .def("nll_grad_weighted",
           [](Handle& s, int kernel_id, int mean_id, std::vector<int64_t> offs,
              int d, ptr_t X, ptr_t y, ptr_t raw, uint64_t mask, ptr_t weight,
              double jitter, ptr_t sums, ptr_t nll_task, ptr_t info,
              ptr_t stream) {
             py::gil_scoped_release rel;
             s.check(hb_nll_grad_weighted(
                         s.h, kernel_id, mean_id, (int)offs.size() - 1,
                         offs.data(), d, P(X), P(y), P(raw), mask, P(weight),
                         jitter, P(sums), P(nll_task), (int32_t*)P(info),
                         P(stream)),
                     "hb_nll_grad_weighted");
           })
      .def("nll_grad_mrhs",
           [](Handle& s, int kernel_id, int mean_id, std::vector<int64_t> offs,
              int d, ptr_t X, int R, ptr_t B, ptr_t col_w, ptr_t col_mean, ptr_t raw,
              uint64_t mask, ptr_t weight, double jitter, ptr_t sums, ptr_t info,
              ptr_t stream) {
             py::gil_scoped_release rel;
             s.check(hb_nll_grad_mrhs(
                         s.h, kernel_id, mean_id, (int)offs.size() - 1,
                         offs.data(), d, P(X), R, P(B), P(col_w),
                         (const int32_t*)P(col_mean), P(raw), mask, P(weight),
                         jitter, P(sums), (int32_t*)P(info), P(stream)),
                     "hb_nll_grad_mrhs");
           })
      .def("euclid_grad",
           [](Handle& s, int kernel_id, int mean_id, std::vector<int64_t> offs,
              int d, ptr_t X, int R, ptr_t Yc, ptr_t mu0, ptr_t raw, uint64_t mask,
              double mean_weight, double cov_weight, ptr_t weight, ptr_t sums,
              ptr_t stream) {
             py::gil_scoped_release rel;
             s.check(hb_euclid_grad(s.h, kernel_id, mean_id, (int)offs.size() - 1,
                                    offs.data(), d, P(X), R, P(Yc), P(mu0), P(raw),
                                    mask, mean_weight, cov_weight, P(weight), P(sums),
                                    P(stream)),
                     "hb_euclid_grad");
           })
      .def("adam_step",
           [](Handle& s, int np, ptr_t raw, ptr_t mm, ptr_t vv, ptr_t accepted,
              ptr_t sums, ptr_t scal, double lr, double b1, double b2, double eps,
              int tie_ls, ptr_t stream) {
             py::gil_scoped_release rel;
             s.check(hb_adam_step(s.h, np, P(raw), P(mm), P(vv), P(accepted),
                                  P(sums), P(scal), lr, b1, b2, eps, tie_ls,
                                  P(stream)),
                     "hb_adam_step");
           })
      .def("generation", [](Handle& s) { return hb_generation(s.h); })
      .def("debug_fused_timeout", [](Handle& s) { return hb_debug_fused_timeout(s.h); })
      .def("nll_grad_multi",
           [](Handle& s, int kernel_id, int mean_id, int S, std::vector<int64_t> offs,
              int d, ptr_t X, ptr_t y, ptr_t raws, uint64_t mask, ptr_t sums,
              ptr_t nll_task, ptr_t stream) {
             py::gil_scoped_release rel;
             s.check(hb_nll_grad_multi(s.h, kernel_id, mean_id, S, (int)offs.size() - 1,
                                       offs.data(), d, P(X), P(y), P(raws), mask, P(sums),
                                       P(nll_task), P(stream)), "hb_nll_grad_multi");
           })
      .def("build_predictors_multi",
           [](Handle& s, int kernel_id, int mean_id, int S, int64_t n, int d, ptr_t X,
              ptr_t y, ptr_t raws, uint64_t mask, ptr_t caches, int64_t stride, ptr_t nll,
              ptr_t info, ptr_t stream) {
             py::gil_scoped_release rel;
             s.check(hb_build_predictors_multi(s.h, kernel_id, mean_id, S, n, d, P(X), P(y),
                                               P(raws), mask, P(caches), stride, P(nll),
                                               (int32_t*)P(info), P(stream)),
                     "hb_build_predictors_multi");
           })
      .def("subsample",
           [](Handle& s, int T, int d, ptr_t offs_src, ptr_t offs_dst, ptr_t ids,
              int64_t max_rows, ptr_t Xs, ptr_t ys, ptr_t Xd, ptr_t yd, uint64_t seed,
              ptr_t scal, int64_t step, ptr_t stream) {
             py::gil_scoped_release rel;
             s.check(hb_subsample(s.h, T, d, P(offs_src), P(offs_dst), P(ids), max_rows,
                                  P(Xs), P(ys), P(Xd), P(yd), seed, P(scal), step,
                                  P(stream)), "hb_subsample");
           })
      .def("bo_cache_bytes", [](Handle& s, int64_t n_cap) { return hb_bo_cache_bytes(s.h, n_cap); })
      .def("bo_init",
           [](Handle& s, int kernel_id, int mean_id, int64_t n0, int64_t n_cap, int d,
              ptr_t X, ptr_t y, ptr_t raw, uint64_t mask, ptr_t cache, ptr_t stream) {
             py::gil_scoped_release rel;
             s.check(hb_bo_init(s.h, kernel_id, mean_id, n0, n_cap, d, P(X), P(y), P(raw),
                                mask, P(cache), P(stream)), "hb_bo_init");
           })
      .def("bo_step",
           [](Handle& s, int kernel_id, int mean_id, int64_t n, int64_t n_cap, int d,
              ptr_t X, ptr_t y, ptr_t raw, uint64_t mask, ptr_t cache, int64_t nq,
              ptr_t Xq, ptr_t yq, double noise_flag, double var_scale, int acq_id,
              double acq_param, int target_is_ymax, ptr_t sel, ptr_t stream) {
             py::gil_scoped_release rel;
             s.check(hb_bo_step(s.h, kernel_id, mean_id, n, n_cap, d, P(X), P(y), P(raw),
                                mask, P(cache), nq, P(Xq), P(yq), noise_flag, var_scale,
                                acq_id, acq_param, target_is_ymax, (int32_t*)P(sel),
                                P(stream)), "hb_bo_step");
           })
      .def("debug_items",
           [](Handle& s, std::vector<int64_t> offs, int d, int variant) {
             int64_t n = hb_debug_items(s.h, (int)offs.size() - 1, offs.data(), d, variant,
                                        nullptr, 0);
             std::vector<int32_t> out((size_t)std::max<int64_t>(n, 0) * 4);
             if (n > 0)
               hb_debug_items(s.h, (int)offs.size() - 1, offs.data(), d, variant, out.data(), n);
             return out;
           })
      .def("set_option", [](Handle& s, const std::string& name, double v) {
             s.check(hb_set_option(s.h, name.c_str(), v), "hb_set_option");
           })
      .def("debug_read",
           [](Handle& s, int which, ptr_t host_out, int64_t max_bytes) {
             return hb_debug_read(s.h, which, P(host_out), max_bytes);
           })
      .def("comm_export",
           [](Handle& s) {
             char buf[HB_IPC_HANDLE_BYTES];
             s.check(hb_comm_export(s.h, buf), "hb_comm_export");
             return py::bytes(buf, HB_IPC_HANDLE_BYTES);
           })
      .def("comm_import",
           [](Handle& s, int rank, int world, const std::string& handles) {
             if ((int)handles.size() != world * HB_IPC_HANDLE_BYTES)
               throw std::invalid_argument("comm_import: handles size");
             s.check(hb_comm_import(s.h, rank, world, handles.data()), "hb_comm_import");
           })
      .def("allreduce",
           [](Handle& s, ptr_t buf, int count, ptr_t stream) {
             py::gil_scoped_release rel;
             s.check(hb_allreduce(s.h, P(buf), count, P(stream)), "hb_allreduce");
           })
      .def("allreduce_adam_step",
           [](Handle& s, int np, ptr_t raw, ptr_t mm, ptr_t vv, ptr_t accepted,
              ptr_t sums, ptr_t scal, double lr, double b1, double b2, double eps,
              int tie_ls, ptr_t stream) {
             py::gil_scoped_release rel;
             s.check(hb_allreduce_adam_step(s.h, np, P(raw), P(mm), P(vv), P(accepted),
                                            P(sums), P(scal), lr, b1, b2, eps, tie_ls,
                                            P(stream)),
                     "hb_allreduce_adam_step");
           })
      .def("build_predictor",
           [](Handle& s, int kernel_id, int mean_id, int64_t n, int d, ptr_t X,
              ptr_t y, ptr_t raw, uint64_t mask, ptr_t cache, ptr_t chol,
              ptr_t kinvy, ptr_t nll, ptr_t info, ptr_t stream) {
             py::gil_scoped_release rel;
             s.check(hb_build_predictor(s.h, kernel_id, mean_id, n, d, P(X), P(y),
                                        P(raw), mask, P(cache), P(chol), P(kinvy),
                                        P(nll), (int32_t*)P(info), P(stream)),
                     "hb_build_predictor");
           })
      .def("predict",
           [](Handle& s, int kernel_id, int mean_id, int64_t n, int d, ptr_t X,
              ptr_t cache, ptr_t raw, uint64_t mask, int64_t nq, ptr_t Xq,
              double noise_flag, double var_scale, int acq_id, double acq_param,
              ptr_t mu, ptr_t var, ptr_t acq, ptr_t stream) {
             py::gil_scoped_release rel;
             s.check(hb_predict(s.h, kernel_id, mean_id, n, d, P(X), P(cache),
                                P(raw), mask, nq, P(Xq), noise_flag, var_scale,
                                acq_id, acq_param, P(mu), P(var), P(acq),
                                P(stream)),
                     "hb_predict");
           })
      .def("predict_cov",
           [](Handle& s, int kernel_id, int mean_id, int64_t n, int d, ptr_t X,
              ptr_t cache, ptr_t raw, uint64_t mask, int64_t nq, ptr_t Xq,
              double noise_flag, double var_scale, ptr_t mu, ptr_t cov, ptr_t stream) {
             py::gil_scoped_release rel;
             s.check(hb_predict_cov(s.h, kernel_id, mean_id, n, d, P(X), P(cache),
                                    P(raw), mask, nq, P(Xq), noise_flag, var_scale,
                                    P(mu), P(cov), P(stream)),
                     "hb_predict_cov");
           })
      .def("acquisition",
           [](Handle& s, int acq_id, double param, int64_t nq, ptr_t mu, ptr_t var,
              ptr_t out, ptr_t stream) {
             py::gil_scoped_release rel;
             s.check(hb_acquisition(s.h, acq_id, param, nq, P(mu), P(var), P(out),
                                    P(stream)),
                     "hb_acquisition");
           });
}
