"""Parity of the CUDA path (through the C ABI) against the CPU oracle.

Tolerances (fp64 engine vs fp64 oracle): nll 1e-10 relative,
gradients 1e-8 relative, chol/alpha 1e-9, predictions / acquisition 1e-6
relative (the north-star figure).  Measured errors are ~1e-13.
"""
import numpy as np
import pytest
import torch

from oracle import hyperbo_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu

WF = O.DEFAULT_WARP_FUNC
TOL_NLL, TOL_GRAD, TOL_FACT, TOL_PRED = 1e-10, 1e-8, 1e-9, 1e-6


@pytest.fixture(scope="module")
def eng():
  from hyperbo_b200.engine import Engine
  return Engine.get()


def _ids(cov, mean):
  from hyperbo_b200.engine import KERNEL_IDS, MEAN_IDS
  return KERNEL_IDS[cov], MEAN_IDS[mean]


def _pack(eng, ds):
  return eng.pack([(k, v[0], v[1]) for k, v in ds.items()])


@pytest.mark.parametrize("name", H.golden_cases())
def test_golden_fixtures(eng, name):
  g = H.load_golden(name)
  d, ns = g["d"], g["ns"]
  kid, mid = _ids(g["cov"], g["mean"])
  mask = H.default_mask(d)
  ds = _pack(eng, g["dataset"])
  chols, alpha, nll, info = eng.factorize(kid, mid, ds, g["raw"], mask)
  assert info.tolist() == [0] * len(ns)
  assert H.rel(nll.cpu().numpy(), g["nll_task"]) < TOL_NLL
  assert H.rel(chols[0].cpu().numpy(), g["chol0"]) < TOL_FACT
  assert np.all(np.triu(chols[0].cpu().numpy(), 1) == 0.0)
  assert H.rel(alpha[:ns[0]].cpu().numpy(), g["alpha0"]) < TOL_FACT
  sums = eng.nll_grad(kid, mid, ds, g["raw"], mask).cpu().numpy()
  T = len(ns)
  assert sums[-1] == T
  assert abs(sums[0] / T - g["mean_nll"]) < TOL_NLL * abs(g["mean_nll"])
  assert H.rel(sums[1:-1] / T, g["grad"]) < TOL_GRAD
  # predict + acquisition on task 0 with GP.predict conventions
  cache, chol, kinvy, _, _ = eng.build_predictor(kid, mid, g["x0"], g["y0"],
                                                 g["raw"], mask)
  scale = T / (T - 1.0) if T > 1 else 1.0
  target = float(np.max(g["y0"]))
  for acq_id, key, param in ((1, "ei", target), (2, "pi", target + 0.1),
                             (3, "ucb", 3.0)):
    mu, var, acq = eng.predict(kid, mid, eng.tensor(g["x0"]), cache, g["raw"],
                               mask, g["xq"], noise_flag=1.0, var_scale=scale,
                               acq_id=acq_id, acq_param=param)
    assert H.rel(mu.cpu().numpy().ravel(), g["mu"]) < TOL_PRED
    assert H.rel(var.cpu().numpy().ravel(), g["var"]) < TOL_PRED
    assert H.rel(acq.cpu().numpy().ravel(), g[key]) < TOL_PRED
  assert H.rel(kinvy.cpu().numpy().ravel(), g["alpha0"]) < TOL_FACT


@pytest.mark.parametrize("cov", O.KERNELS)
@pytest.mark.parametrize("ns,d", [([1], 1), ([63, 64, 65], 2), ([200, 5, 129], 5),
                                  ([512, 300], 8), ([96], 32),
                                  # d > 12: X blocks staged in a ring stage
                                  ([130, 70, 300], 16), ([100, 65], 30)])
def test_nll_grad_vs_oracle(eng, cov, ns, d):
  ds_np = {t: O.make_task(7 * len(ns) + t, n, d, cov) for t, n in enumerate(ns)}
  model = O.init_raw_params(d)
  model["lengthscale"] = np.random.default_rng(d).normal(0, 0.3, d)
  kid, mid = _ids(cov, "constant")
  sums = eng.nll_grad(kid, mid, _pack(eng, ds_np), H.raw_vec(model, d),
                      H.default_mask(d)).cpu().numpy()
  v_ref, g_ref = O.nll_value_and_grad("constant", cov, model, ds_np, WF)
  T = len(ns)
  assert abs(sums[0] / T - v_ref) < TOL_NLL * abs(v_ref)
  assert H.rel(sums[1:-1] / T, H.grad_vec(g_ref, d)) < TOL_GRAD


def test_identity_warp_and_zero_mean(eng):
  d = 3
  ds_np = {t: O.make_task(t, 50 + 20 * t, d) for t in range(3)}
  model = {"lengthscale": np.array([0.7, 1.1, 0.4]), "signal_variance": 1.3,
           "noise_variance": 0.05}
  kid, mid = _ids("squared_exponential", "zero")
  sums = eng.nll_grad(kid, mid, _pack(eng, ds_np), H.raw_vec(model, d),
                      0).cpu().numpy()
  v_ref, g_ref = O.nll_value_and_grad("zero", "squared_exponential", model,
                                      ds_np, None)
  assert abs(sums[0] / 3 - v_ref) < TOL_NLL * abs(v_ref)
  want = H.grad_vec(g_ref, d)
  assert sums[1] == 0.0  # no constant parameter with mean.zero
  assert H.rel(sums[2:-1] / 3, want[1:]) < TOL_GRAD


def test_duplicate_points_matern_gradient(eng):
  # r = 0 off the diagonal: the reference's _safe_sqrt yields a zero gradient
  x, y = O.make_task(3, 80, 2, "matern32")
  x[10] = x[20]
  model = O.init_raw_params(2)
  for cov in ("matern32", "matern52"):
    kid, mid = _ids(cov, "constant")
    sums = eng.nll_grad(kid, mid, _pack(eng, {0: (x, y)}), H.raw_vec(model, 2),
                        H.default_mask(2)).cpu().numpy()
    v_ref, g_ref = O.nll_value_and_grad("constant", cov, model, {0: (x, y)}, WF)
    assert np.all(np.isfinite(sums))
    assert abs(sums[0] - v_ref) < TOL_NLL * abs(v_ref)
    assert H.rel(sums[1:-1], H.grad_vec(g_ref, 2)) < TOL_GRAD


@pytest.mark.parametrize("cov", O.KERNELS)
def test_kernel_matrix(eng, cov):
  rng = np.random.default_rng(1)
  x1, x2 = rng.normal(size=(150, 3)), rng.normal(size=(70, 3))
  model = {"lengthscale": np.array([0.3, -0.2, 0.9]), "signal_variance": 0.4,
           "noise_variance": -3.0}
  raw, mask = H.raw_vec(model, 3), H.default_mask(3)
  kid, _ = _ids(cov, "constant")
  k11 = eng.kernel_matrix(kid, x1, None, raw, mask).cpu().numpy()
  k12 = eng.kernel_matrix(kid, x1, x2, raw, mask).cpu().numpy()
  kd = eng.kernel_matrix(kid, x1, None, raw, mask, diag=True).cpu().numpy()
  kn = eng.kernel_matrix(kid, x1, None, raw, mask, add_noise=True).cpu().numpy()
  assert H.rel(k11, O.cov_matrix(cov, model, x1, warp_func=WF)) < 1e-13
  assert H.rel(k12, O.cov_matrix(cov, model, x1, x2, warp_func=WF)) < 1e-13
  assert kd.shape == (150,) and H.rel(kd, O.cov_matrix(cov, model, x1, warp_func=WF, diag=True)) < 1e-14
  _, kref = O.compute_delta_y_and_cov("constant", cov, dict(model, constant=0.0),
                                      x1, np.zeros((150, 1)), WF)
  assert H.rel(kn, kref) < 1e-13
  # kernel_test.py:77-152: symmetric, PSD
  assert np.array_equal(k11, k11.T)
  assert np.linalg.eigvalsh(k11).min() > -1e-10


def test_empty_tasks_are_skipped(eng):
  ds_np = {0: O.make_task(0, 30, 2), 1: (np.zeros((0, 2)), np.zeros((0, 1))),
           2: O.make_task(2, 70, 2)}
  model = O.init_raw_params(2)
  kid, mid = _ids("squared_exponential", "constant")
  sums = eng.nll_grad(kid, mid, _pack(eng, ds_np), H.raw_vec(model, 2),
                      H.default_mask(2)).cpu().numpy()
  v_ref, _ = O.nll_value_and_grad("constant", "squared_exponential", model,
                                  ds_np, WF)
  assert sums[-1] == 2 and abs(sums[0] / 2 - v_ref) < TOL_NLL * abs(v_ref)
  # C ABI level: n_t = 0 inside offs is allowed
  from hyperbo_b200.engine import PackedDataset
  ds = _pack(eng, {0: ds_np[0], 2: ds_np[2]})
  ds0 = PackedDataset([0, 1, 2], ds.x, ds.y, [0, 30, 30, 100])
  s2 = eng.nll_grad(kid, mid, ds0, H.raw_vec(model, 2), H.default_mask(2)).cpu().numpy()
  assert s2[-1] == 2 and np.allclose(s2, sums, rtol=1e-14)
  # no tasks at all -> zeros (objectives.py:192-193)
  z = eng.nll_grad(kid, mid, _pack(eng, {}), H.raw_vec(O.init_raw_params(1), 1), 0).cpu().numpy()
  assert np.all(z == 0.0)


def test_non_pd_sets_info_and_nan_without_raising(eng):
  # task 0: 40 identical points with the noise cancelling the jitter -> K~ is
  # the rank-1 all-ones matrix, breakdown at column 2.  Task 1: a well-spaced
  # grid that stays PD without noise at this short lengthscale.
  x = np.full((40, 1), 0.5)
  y = np.ones((40, 1))
  x1 = np.linspace(0.0, 1.0, 20)[:, None]
  y1 = np.sin(6 * x1)
  model = {"constant": 0.0, "lengthscale": np.array([0.05]),
           "signal_variance": 1.0, "noise_variance": -1e-6}
  kid, mid = _ids("squared_exponential", "constant")
  ds = _pack(eng, {0: (x, y), 1: (x1, y1)})
  _, _, nll, info = eng.factorize(kid, mid, ds, H.raw_vec(model, 1), 0,
                                  want_chol=False, want_alpha=False)
  assert info[0].item() == 2 and np.isnan(nll[0].item())
  assert info[1].item() == 0 and np.isfinite(nll[1].item())
  ref = O.nll_sub_dataset("constant", "squared_exponential", model, x1, y1, None)
  assert abs(nll[1].item() - ref) < 1e-9 * abs(ref)
  sums = eng.nll_grad(kid, mid, ds, H.raw_vec(model, 1), 0).cpu().numpy()
  assert np.isnan(sums[0])  # the mean loss is NaN -> the host loop stops


def test_many_small_ragged_tasks(eng):
  # T > 256 exercises the grouped launch order of k_step; sizes straddle the
  # 64-point tile edge (1 and 2 block columns), incl. n = 1 and n = 64
  rng = np.random.default_rng(11)
  ns = [int(v) for v in rng.integers(1, 130, 300)] + [1, 64, 65, 128]
  d = 2
  ds_np = {t: O.make_task(t, n, d, "matern52", surrogate=True)
           for t, n in enumerate(ns)}
  model = O.init_raw_params(d)
  kid, mid = _ids("matern52", "constant")
  sums, nll_task = eng.nll_grad(kid, mid, _pack(eng, ds_np), H.raw_vec(model, d),
                                H.default_mask(d), want_task_nll=True)
  sums = sums.cpu().numpy()
  v_ref, g_ref = O.nll_value_and_grad("constant", "matern52", model, ds_np, WF)
  T = len(ns)
  assert sums[-1] == T
  assert abs(sums[0] / T - v_ref) < TOL_NLL * abs(v_ref)
  assert H.rel(sums[1:-1] / T, H.grad_vec(g_ref, d)) < TOL_GRAD
  for t in (0, 150, 299, 300, 303):
    ref = O.nll_sub_dataset("constant", "matern52", model, *ds_np[t], warp_func=WF)
    assert abs(nll_task[t].item() - ref) < 1e-9 * max(1.0, abs(ref))


def test_large_n_blocked_path(eng):
  # 17 block columns, ragged second task
  ns, d = [1040, 777], 6
  ds_np = {t: O.make_task(t, n, d, "matern52", surrogate=True)
           for t, n in enumerate(ns)}
  model = O.init_raw_params(d)
  kid, mid = _ids("matern52", "constant")
  sums = eng.nll_grad(kid, mid, _pack(eng, ds_np), H.raw_vec(model, d),
                      H.default_mask(d)).cpu().numpy()
  v_ref, g_ref = O.nll_value_and_grad("constant", "matern52", model, ds_np, WF)
  assert abs(sums[0] / 2 - v_ref) < TOL_NLL * abs(v_ref)
  assert H.rel(sums[1:-1] / 2, H.grad_vec(g_ref, d)) < TOL_GRAD


def test_full_size_properties_256x512x8(eng):
  """BASELINE config 2 at full size, through size-independent properties:
  task-permutation invariance, shard linearity (what the multi-GPU all-reduce
  relies on), and a spot check of 3 tasks against the oracle."""
  T, n, d = 256, 512, 8
  rng = np.random.default_rng(0)
  x = rng.random((T, n, d))
  y = 5.0 + rng.standard_normal((T, n, 1))
  model = O.init_raw_params(d)
  raw, mask = H.raw_vec(model, d), H.default_mask(d)
  kid, mid = _ids("squared_exponential", "constant")
  full = {t: (x[t], y[t]) for t in range(T)}
  s_full, nll_task = eng.nll_grad(kid, mid, _pack(eng, full), raw, mask,
                                  want_task_nll=True)
  s_full, nll_task = s_full.cpu().numpy(), nll_task.cpu().numpy()
  assert s_full[-1] == T and np.all(np.isfinite(s_full))
  perm = rng.permutation(T)
  s_perm = eng.nll_grad(kid, mid, _pack(eng, {i: full[t] for i, t in enumerate(perm)}),
                        raw, mask).cpu().numpy()
  assert H.rel(s_perm, s_full) < 1e-12
  parts = [eng.nll_grad(kid, mid, _pack(eng, {t: full[t] for t in range(T) if t % 4 == r}),
                        raw, mask).cpu().numpy() for r in range(4)]
  assert H.rel(sum(parts), s_full) < 1e-12
  for t in (0, 100, 255):
    ref = O.nll_sub_dataset("constant", "squared_exponential", model, x[t], y[t], WF)
    assert abs(nll_task[t] - ref) < TOL_NLL * abs(ref)
  # determinism: same inputs -> bitwise identical sums
  again = eng.nll_grad(kid, mid, _pack(eng, full), raw, mask).cpu().numpy()
  assert np.array_equal(again, s_full)


def test_adam_loop_matches_oracle(eng):
  from hyperbo_b200.gp_utils.gp import AdamTrainer
  d = 3
  ds_np = {t: O.make_task(t, 40 + 30 * t, d, "matern52") for t in range(4)}
  model = O.init_raw_params(d)
  kid, mid = _ids("matern52", "constant")
  tr = AdamTrainer(eng, kid, mid, H.raw_vec(model, d), H.default_mask(d), d, 1e-2)
  ds = _pack(eng, ds_np)
  losses = []
  for i in range(8):
    tr.step(ds, use_graph=(i >= 2))  # eager, then CUDA-graph replay
    losses.append(tr.loss())
  ref_model, ref_losses = O.infer_parameters_adam(
      "constant", "matern52", model, ds_np, WF, 1e-2, 8, 10**6)
  assert H.rel(losses, ref_losses) < 1e-9
  # raw holds the params after 8 updates == oracle's final accepted params
  assert H.rel(tr.raw.cpu().numpy(), H.raw_vec(ref_model, d)) < 1e-8
  assert not tr.stopped and tr.scal[3].item() == 8


def test_adam_stops_on_non_finite_loss(eng):
  from hyperbo_b200.gp_utils.gp import AdamTrainer
  x = np.full((30, 1), 0.5)
  y = np.ones((30, 1))
  raw = np.array([0.0, 1.0, -1e-6, 1.0])
  kid, mid = _ids("squared_exponential", "constant")
  tr = AdamTrainer(eng, kid, mid, raw, 0, 1, 1e-2)
  ds = _pack(eng, {0: (x, y)})
  tr.step(ds)
  assert np.isnan(tr.loss()) and tr.stopped
  tr.step(ds)
  assert np.array_equal(tr.raw.cpu().numpy(), raw) and tr.scal[3].item() == 0


def test_tied_scalar_lengthscale_adam(eng):
  from hyperbo_b200.gp_utils.gp import AdamTrainer
  d = 3
  ds_np = {t: O.make_task(t, 50, d) for t in range(2)}
  model = {"constant": 5.1, "lengthscale": 0.2, "signal_variance": 0.0,
           "noise_variance": -4.0}
  kid, mid = _ids("squared_exponential", "constant")
  tr = AdamTrainer(eng, kid, mid, H.raw_vec(model, d), H.default_mask(d), d, 1e-2,
                   tie_lengthscale=True)
  ds = _pack(eng, ds_np)
  for _ in range(4):
    tr.step(ds)
  ref_model, _ = O.infer_parameters_adam("constant", "squared_exponential", model,
                                         ds_np, WF, 1e-2, 4, 10**6)
  raw = tr.raw.cpu().numpy()
  assert raw[3] == raw[4] == raw[5]
  assert abs(raw[3] - float(ref_model["lengthscale"])) < 1e-9


# ---- CUDA path vs the 60-digit mpmath known answers (generated without the
# oracle: tests/golden/make_mpmath_kat.py)
@pytest.mark.parametrize("case", H.load_kat(), ids=lambda c: "kat%d_%s_%s_%s" % (
    c["id"], c["cov"], c["mean"], "warp" if c["warped"] else "raw"))
def test_cuda_matches_mpmath_known_answers(eng, case):
  c = case
  d, T = c["d"], len(c["ns"])
  kid, mid = _ids(c["cov"], c["mean"])
  mask = H.default_mask(d) if c["warped"] else 0
  ds = _pack(eng, c["dataset"])
  _, alpha, nll, info = eng.factorize(kid, mid, ds, c["raw"], mask)
  assert info.tolist() == [0] * T
  assert H.rel(nll.cpu().numpy(), c["nll_task"]) < TOL_NLL
  assert H.rel(alpha[:c["ns"][0]].cpu().numpy(), c["alpha0"]) < TOL_FACT
  sums = eng.nll_grad(kid, mid, ds, c["raw"], mask).cpu().numpy()
  assert abs(sums[0] / T - c["mean_nll"]) < TOL_NLL * abs(c["mean_nll"])
  assert H.rel(sums[1:-1] / T, c["grad"]) < TOL_GRAD
  x0, y0 = c["dataset"][0]
  cache, _, kinvy, _, _ = eng.build_predictor(kid, mid, x0, y0, c["raw"], mask)
  scale = T / (T - 1.0) if T > 1 else 1.0
  target = float(np.max(y0))
  for acq_id, key, param in ((1, "ei", target), (2, "pi", target + 0.1),
                             (3, "ucb", 3.0)):
    mu, var, acq = eng.predict(kid, mid, eng.tensor(x0), cache, c["raw"], mask,
                               c["xq"], noise_flag=1.0, var_scale=scale,
                               acq_id=acq_id, acq_param=param)
    assert H.rel(mu.cpu().numpy().ravel(), c["mu"]) < TOL_PRED
    assert H.rel(var.cpu().numpy().ravel(), c["var"]) < TOL_PRED
    assert H.rel(acq.cpu().numpy().ravel(), c[key]) < TOL_PRED
  # full posterior covariance (hb_predict_cov) against the same known answers
  mu, cov = eng.predict_cov(kid, mid, eng.tensor(x0), cache, c["raw"], mask, c["xq"],
                            noise_flag=1.0, var_scale=scale)
  assert H.rel(mu.cpu().numpy().ravel(), c["mu"]) < TOL_PRED
  assert np.abs(cov.cpu().numpy() - c["cov_full"]).max() < \
      TOL_PRED * np.abs(c["cov_full"]).max()
