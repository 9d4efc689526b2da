"""Pins the CPU oracle (not gpu): golden fixtures, the reference's own test
identities, closed-form known answers, finite differences, and
agreement of the two independent restatements."""
import math

import numpy as np
import pytest

from oracle import hyperbo_oracle as O
from oracle import hyperbo_oracle_torch as OT
from tests import helpers as H

WF = O.DEFAULT_WARP_FUNC


@pytest.mark.parametrize("name", H.golden_cases())
def test_oracle_matches_golden(name):
  g = H.load_golden(name)
  model = H.model_from_raw(g["raw"], g["d"], g["mean"])
  val, grad = O.nll_value_and_grad(g["mean"], g["cov"], model, g["dataset"], WF)
  assert abs(val - g["mean_nll"]) <= 1e-12 * abs(g["mean_nll"])
  assert H.rel(H.grad_vec(grad, g["d"]), g["grad"]) < 1e-10
  mu, var = O.gp_predict(g["mean"], g["cov"], model, g["dataset"], g["xq"], 0, WF)
  assert H.rel(mu.ravel(), g["mu"]) < 1e-12
  assert H.rel(var.ravel(), g["var"]) < 1e-10


@pytest.mark.parametrize("cov", O.KERNELS)
def test_two_oracles_agree_incl_duplicate_points(cov):
  ds = O.make_dataset(3, 40, 3, cov)
  x, y = ds[0]
  x[5] = x[7]  # r = 0 off the diagonal: exercises _safe_sqrt (linalg.py:175-197)
  ds[0] = (x, y)
  model = O.init_raw_params(3)
  model["lengthscale"] = np.array([0.1, -0.3, 0.5])
  v, g = O.nll_value_and_grad("constant", cov, model, ds, WF)
  vt, gt = OT.value_and_grad("constant", cov, model, ds)
  assert abs(v - vt) < 1e-12 * abs(v)
  for k in g:
    assert H.rel(g[k], gt[k]) < 1e-11, k


def test_scalar_lengthscale_gradient_is_summed():
  model = {"constant": 5.1, "lengthscale": 0.2, "signal_variance": 0.0,
           "noise_variance": -4.0}
  ds = O.make_dataset(2, 30, 4)
  v, g = O.nll_value_and_grad("constant", "matern52", model, ds, WF)
  vt, gt = OT.value_and_grad("constant", "matern52", model, ds)
  assert abs(v - vt) < 1e-12 * abs(v)
  assert abs(float(g["lengthscale"]) - float(gt["lengthscale"])) < 1e-10


@pytest.mark.parametrize("cov", O.KERNELS)
def test_finite_differences(cov):
  ds = O.make_dataset(2, 25, 2, cov)
  model = O.init_raw_params(2)
  _, g = O.nll_value_and_grad("constant", cov, model, ds, WF)
  h = 1e-6
  for key in ("constant", "signal_variance", "noise_variance"):
    mp, mm = dict(model), dict(model)
    mp[key] = model[key] + h
    mm[key] = model[key] - h
    fd = (O.neg_log_marginal_likelihood("constant", cov, mp, ds, WF) -
          O.neg_log_marginal_likelihood("constant", cov, mm, ds, WF)) / (2 * h)
    assert abs(fd - g[key]) < 1e-6 * max(1.0, abs(fd))
  for k in range(2):
    mp, mm = dict(model), dict(model)
    e = np.zeros(2)
    e[k] = h
    mp["lengthscale"] = model["lengthscale"] + e
    mm["lengthscale"] = model["lengthscale"] - e
    fd = (O.neg_log_marginal_likelihood("constant", cov, mp, ds, WF) -
          O.neg_log_marginal_likelihood("constant", cov, mm, ds, WF)) / (2 * h)
    assert abs(fd - g["lengthscale"][k]) < 1e-6 * max(1.0, abs(fd))


def test_known_answer_n1_and_n2():
  # n = 1: nll = .5 (y-c)^2 / v + .5 log v + .5 log 2pi, v = sf2 + sn2 + 1e-6
  model = {"constant": 0.3, "lengthscale": 0.7, "signal_variance": 1.3,
           "noise_variance": 0.2}
  x, y = np.array([[0.4]]), np.array([[1.1]])
  v = 1.3 + 0.2 + 1e-6
  want = 0.5 * (1.1 - 0.3)**2 / v + 0.5 * math.log(v) + 0.5 * math.log(2 * math.pi)
  got = O.nll_sub_dataset("constant", "squared_exponential", model, x, y)
  assert abs(got - want) < 1e-14
  # n = 2 closed form
  x = np.array([[0.0], [0.5]])
  y = np.array([[1.0], [-0.5]])
  k = 1.3 * math.exp(-0.5 * (0.5 / 0.7)**2)
  a = v
  det = a * a - k * k
  r = y.ravel() - 0.3
  quad = (a * r[0]**2 - 2 * k * r[0] * r[1] + a * r[1]**2) / det
  want = 0.5 * quad + 0.5 * math.log(det) + math.log(2 * math.pi)
  got = O.nll_sub_dataset("constant", "squared_exponential", model, x, y)
  assert abs(got - want) < 1e-13


@pytest.mark.parametrize("cov", O.KERNELS)
def test_gram_shape_symmetry_psd(cov):  # kernel_test.py:77-152
  rng = np.random.default_rng(0)
  x1, x2 = rng.normal(size=(30, 3)), rng.normal(size=(20, 3))
  model = {"lengthscale": np.array([0.5, 1.0, 2.0]), "signal_variance": 1.7}
  k12 = O.cov_matrix(cov, model, x1, x2)
  assert k12.shape == (30, 20)
  k11 = O.cov_matrix(cov, model, x1)
  assert np.allclose(k11, k11.T, atol=1e-12)
  assert np.linalg.eigvalsh(k11).min() > -1e-10
  assert np.allclose(np.diag(k11), 1.7)
  assert O.cov_matrix(cov, model, x1, diag=True).shape == (30,)
  # diag is ignored when vx2 is given (kernel.py:54-58)
  assert O.cov_matrix(cov, model, x1, x2, diag=True).shape == (30, 20)


def test_predict_identities():  # gp_test.py:150-207
  x, y = O.make_task(3, 20, 1)
  xq = np.random.default_rng(1).normal(size=(10, 1))
  model = dict(O.GROUND_TRUTH)
  mu, var = O.predict("constant", "squared_exponential", model, x, y, xq)
  mu_m, var_m = O.gp_predict("constant", "squared_exponential", model,
                             {0: (x, y)}, xq, 0, with_noise=True)
  assert mu.shape == (10, 1) and var.shape == (10, 1)
  assert np.allclose(mu, mu_m, atol=1e-12)
  assert np.allclose(var + model["noise_variance"], var_m, atol=1e-12)
  mu2, cov = O.predict("constant", "squared_exponential", model, x, y, xq,
                       full_cov=True)
  assert cov.shape == (10, 10)
  assert np.allclose(np.diag(cov), var.ravel(), atol=1e-9)
  # prior branch with no observations (gp.py:275-282)
  mu0, var0 = O.predict("constant", "squared_exponential", model, None, None, xq)
  assert np.allclose(mu0, 5.0) and np.allclose(var0, 1.0)


def test_svd_nll_matches_cholesky_nll():  # objectives_test.py:298-301
  ds = O.make_dataset(3, 20, 2)
  model = O.init_raw_params(2)
  a = O.neg_log_marginal_likelihood("constant", "squared_exponential", model, ds,
                                    WF, use_cholesky=True)
  b = O.neg_log_marginal_likelihood("constant", "squared_exponential", model, ds,
                                    WF, use_cholesky=False)
  assert abs(a / b - 1) < 1e-8


def test_unbiased_inflation_and_noise_without_jitter():  # gp.py:607-619
  ds = O.make_dataset(3, 15, 2)
  xq = np.random.default_rng(2).random((5, 2))
  model = O.init_raw_params(2)
  _, v1 = O.gp_predict("constant", "squared_exponential", model, ds, xq, 0, WF,
                       with_noise=False, unbiased=False)
  _, v2 = O.gp_predict("constant", "squared_exponential", model, ds, xq, 0, WF,
                       with_noise=True, unbiased=True)
  nv = float(O.default_softplus(model["noise_variance"]))
  assert np.allclose(v2, (v1 + nv) * 1.5, rtol=1e-12)


def test_training_decreases_nll_and_adam_matches_torch():  # gp_test.py:148
  ds = O.make_dataset(4, 30, 1)
  model = O.init_raw_params(1)
  init = O.neg_log_marginal_likelihood("constant", "squared_exponential", model,
                                       ds, WF)
  out, losses = O.infer_parameters_adam("constant", "squared_exponential", model,
                                        ds, WF, 1e-2, 5, 100)
  final = O.neg_log_marginal_likelihood("constant", "squared_exponential", out,
                                        ds, WF)
  assert final < init and len(losses) == 5
  x = np.stack([ds[t][0] for t in range(4)])
  y = np.stack([ds[t][1] for t in range(4)])
  tr = OT.AdamTrainer("constant", "squared_exponential", model, x, y, lr=1e-2)
  tl = [tr.step() for _ in range(5)]
  assert H.rel(tl, losses) < 1e-10


def test_skips_empty_and_aligned_tasks():  # objectives.py:181-185
  ds = O.make_dataset(2, 10, 1)
  full = O.neg_log_marginal_likelihood("constant", "squared_exponential",
                                       O.init_raw_params(1), ds, WF)
  ds[7] = (np.zeros((0, 1)), np.zeros((0, 1)))
  ds[8] = (ds[0][0], ds[0][1], "aligned-tag")
  again = O.neg_log_marginal_likelihood("constant", "squared_exponential",
                                        O.init_raw_params(1), ds, WF)
  assert full == again


def test_acquisition_shapes_and_values():  # acfun_test.py:43-72
  ds = O.make_dataset(2, 12, 2)
  xq = np.random.default_rng(3).random((9, 2))
  for name in ("ei", "pi", "pi2", "pi3", "ucb", "ucb2", "ucb4"):
    out = O.acquisition(name, "constant", "matern32", O.init_raw_params(2), ds,
                        0, xq, WF)
    assert out.shape == (9, 1) and np.all(np.isfinite(out))
  # EI >= 0 and EI(mu=target, std) = std * pdf(0)
  assert np.isclose(O.expected_improvement_sub(1.0, 2.0, 1.0),
                    2.0 / math.sqrt(2 * math.pi))


# ---- divergence objectives on aligned data -------------
@pytest.mark.parametrize("name", H.golden_cases(kl=True))
def test_kl_oracle_matches_golden(name):
  import functools
  g = H.load_golden_kl(name)
  model = H.model_from_raw(g["raw"], g["d"], g["mean"])
  val, grad = O.kl_value_and_grad(g["mean"], g["cov"], model, g["dataset"], WF)
  assert abs(val - g["kl"]) <= 1e-12 * abs(g["kl"])
  assert H.rel(H.grad_vec(grad, g["d"]), g["kl_grad"]) < 1e-10
  # direct restatement of objectives.py:29-101 == closed form
  v0 = O.multivariate_normal_divergence(g["mean"], g["cov"], model,
                                        g["dataset"], WF)
  assert abs(v0 - val) <= 1e-11 * abs(val)
  # op-by-op torch restatement incl. the whitened KL and the Euclidean distance
  vt, gt = OT.divergence_value_and_grad(g["mean"], g["cov"], model, g["dataset"])
  assert abs(vt - val) <= 1e-11 * abs(val)
  for k in grad:
    assert H.rel(grad[k], gt[k]) < 1e-10, k
  vf, _ = OT.divergence_value_and_grad(g["mean"], g["cov"], model, g["dataset"],
                                       eps=1e-6, partial=False)
  assert abs(vf - g["kl_full"]) <= 1e-8 * abs(g["kl_full"])
  ve, _ = OT.divergence_value_and_grad(g["mean"], g["cov"], model, g["dataset"],
                                       euc=True)
  assert abs(ve - g["euc"]) <= 1e-12 * abs(g["euc"])


def test_kl_is_a_weighted_sum_of_task_nlls():
  """The identity the engine's KL program rests on (objectives.py of the
  package): with B = [Yc/sqrt(m) | mu0], partial KL = 2 sum_q nll0(B_q) +
  2 nll_mean(mu0) - 2 m nll0(0) - n log 2pi, all with jitter = eps."""
  rng = np.random.default_rng(3)
  n, m, d = 45, 5, 2
  x = rng.random((n, d))
  y = rng.standard_normal((n, m)) + 2.0 * x[:, :1]
  model = O.init_raw_params(d)
  model["constant"] = 0.7
  for eps in (0.0, 1e-6):
    def task_nll(mean_name, yy):
      chol, kinvy, r = O.solve_gp_linear_system(mean_name, "matern32", model, x,
                                                yy[:, None], WF, eps=eps)
      return float(0.5 * (r.T @ kinvy).item() + np.sum(np.log(np.diag(chol))) +
                   0.5 * n * math.log(2 * math.pi))
    mu0 = y.mean(axis=1)
    yc = (y - mu0[:, None]) / math.sqrt(m)
    total = 2 * sum(task_nll("zero", yc[:, q]) for q in range(m)) + \
        2 * task_nll("constant", mu0) - 2 * m * task_nll("zero", np.zeros(n)) - \
        n * math.log(2 * math.pi)
    import functools
    ref = O.multivariate_normal_divergence(
        "constant", "matern32", model, {0: (x, y, 1)}, WF,
        functools.partial(O.kl_multivariate_normal, eps=eps))
    if eps > 0:  # cov0 + eps I adds eps tr(K1^-1)
      _, cov1 = O.compute_delta_y_and_cov("constant", "matern32", model, x,
                                          y[:, :1], WF, eps=eps)
      total += eps * np.trace(np.linalg.inv(cov1))
    assert abs(total - ref) < 1e-10 * abs(ref), (eps, total, ref)


def test_kl_gradient_finite_differences():
  ds = {0: (np.random.default_rng(0).random((30, 2)),
            np.random.default_rng(1).standard_normal((30, 4)), 1)}
  model = O.init_raw_params(2)
  model["lengthscale"] = np.array([0.2, -0.1])
  _, g = O.kl_value_and_grad("constant", "matern52", model, ds, WF)
  h = 1e-6
  for key in ("constant", "signal_variance", "noise_variance"):
    mp, mm = dict(model), dict(model)
    mp[key], mm[key] = model[key] + h, model[key] - h
    fd = (O.multivariate_normal_divergence("constant", "matern52", mp, ds, WF) -
          O.multivariate_normal_divergence("constant", "matern52", mm, ds, WF)
          ) / (2 * h)
    assert abs(fd - g[key]) < 1e-5 * max(1.0, abs(fd)), key
  for k in range(2):
    mp, mm = dict(model), dict(model)
    e = np.zeros(2); e[k] = h
    mp["lengthscale"], mm["lengthscale"] = model["lengthscale"] + e, \
        model["lengthscale"] - e
    fd = (O.multivariate_normal_divergence("constant", "matern52", mp, ds, WF) -
          O.multivariate_normal_divergence("constant", "matern52", mm, ds, WF)
          ) / (2 * h)
    assert abs(fd - g["lengthscale"][k]) < 1e-5 * max(1.0, abs(fd)), k


# ---- further pure-maths pins (SURVEY.md 8c "extra known-answer checks") -------
@pytest.mark.parametrize("cov", O.KERNELS)
def test_permutation_invariance_and_diagonal(cov):
  x, y = O.make_task(3, 37, 3, cov)
  model = O.init_raw_params(3)
  model["lengthscale"] = np.array([0.3, -0.2, 0.1])
  v = O.nll_sub_dataset("constant", cov, model, x, y, WF)
  perm = np.random.default_rng(0).permutation(37)
  vp = O.nll_sub_dataset("constant", cov, model, x[perm], y[perm], WF)
  assert abs(v - vp) < 1e-11 * abs(v)
  k = O.cov_matrix(cov, model, x, warp_func=WF)
  (sv,) = O.retrieve_params(model, ["signal_variance"], WF)
  assert np.allclose(np.diag(k), float(sv), rtol=0, atol=1e-15)  # k(x,x) = sigma_f^2
  assert np.allclose(O.cov_matrix(cov, model, x, warp_func=WF, diag=True),
                     float(sv))
  # diag is ignored when vx2 is given (kernel.py:54-58)
  assert O.cov_matrix(cov, model, x, x[:5], warp_func=WF, diag=True).shape == (37, 5)


def test_infinite_lengthscale_gives_rank_one_gram():
  x = np.random.default_rng(1).random((20, 2))
  model = {"constant": 0.0, "lengthscale": 1e9, "signal_variance": 2.5,
           "noise_variance": 0.1}
  for cov in O.KERNELS:
    k = O.cov_matrix(cov, model, x)
    assert np.allclose(k, 2.5, atol=1e-7)
    s = np.linalg.svd(k, compute_uv=False)
    assert s[1] < 1e-6 * s[0]


def test_task_sum_linearity_across_shards():
  """What the multi-GPU reduction relies on: [sum nll, sum grad, count] over
  disjoint task shards add up to the single-process sums."""
  d = 2
  ds = O.make_dataset(7, 15, d, "matern52")
  model = O.init_raw_params(d)
  def sums(keys):
    tot, g = 0.0, np.zeros(3 + d)
    for k in keys:
      v, gr = O.nll_and_grad_sub_dataset("constant", "matern52", model, *ds[k],
                                         warp_func=WF)
      tot += v
      g += H.grad_vec(gr, d)
    return tot, g
  full = sums(range(7))
  parts = [sums([k for k in range(7) if k % 3 == r]) for r in range(3)]
  assert abs(sum(p[0] for p in parts) - full[0]) < 1e-12 * abs(full[0])
  assert H.rel(sum(p[1] for p in parts), full[1]) < 1e-12
  v, gr = O.nll_value_and_grad("constant", "matern52", model, ds, WF)
  assert abs(v - full[0] / 7) < 1e-13 * abs(v)
  assert H.rel(H.grad_vec(gr, d), full[1] / 7) < 1e-12


# ---- independent pin: 60-digit mpmath evaluation of the reference formulas
# (tests/golden/make_mpmath_kat.py does not import oracle/)
KAT = H.load_kat()


@pytest.mark.parametrize("case", KAT, ids=lambda c: "kat%d_%s_%s_%s" % (
    c["id"], c["cov"], c["mean"], "warp" if c["warped"] else "raw"))
def test_oracle_matches_mpmath_known_answers(case):
  c = case
  wf = O.DEFAULT_WARP_FUNC if c["warped"] else None
  model = H.model_from_raw(c["raw"], c["d"], c["mean"])
  ds = c["dataset"]
  val, grad = O.nll_value_and_grad(c["mean"], c["cov"], model, ds, wf)
  assert abs(val - c["mean_nll"]) <= 1e-12 * abs(c["mean_nll"])
  for t, (x, y) in ds.items():
    v = O.nll_sub_dataset(c["mean"], c["cov"], model, x, y, wf)
    assert abs(v - c["nll_task"][t]) <= 1e-12 * abs(c["nll_task"][t])
  g = H.grad_vec(grad, c["d"])
  if c["mean"] == "zero":
    g[0] = 0.0
  assert H.rel(g, c["grad"]) < 1e-10
  mu, var = O.gp_predict(c["mean"], c["cov"], model, ds, c["xq"], 0, wf)
  assert H.rel(np.ravel(mu), c["mu"]) < 1e-11
  assert H.rel(np.ravel(var), c["var"]) < 1e-9
  for name in ("ei", "pi", "ucb"):
    a = O.acquisition(name, c["mean"], c["cov"], model, ds, 0, c["xq"], wf)
    assert H.rel(np.ravel(a), c[name]) < 1e-8, name
  _, cov_full = O.gp_predict(c["mean"], c["cov"], model, ds, c["xq"], 0, wf,
                             full_cov=True)
  assert np.abs(cov_full - c["cov_full"]).max() < 1e-9 * np.abs(c["cov_full"]).max()
  _, kinvy, _ = O.solve_gp_linear_system(c["mean"], c["cov"], model, ds[0][0],
                                         ds[0][1], wf)
  assert H.rel(np.ravel(kinvy), c["alpha0"]) < 1e-9


# ---- divergence objectives on aligned data against 60-digit known answers -------
# (tests/golden/make_mpmath_kat_div.py does not import oracle/)
KAT_DIV = H.load_kat_div()


@pytest.mark.parametrize("case", KAT_DIV, ids=lambda c: "katdiv%d_%s_%s_%s" % (
    c["id"], c["cov"], c["mean"], "warp" if c["warped"] else "raw"))
def test_oracle_divergences_match_mpmath_known_answers(case):
  c = case
  wf = O.DEFAULT_WARP_FUNC if c["warped"] else None
  model = H.model_from_raw(c["raw"], c["d"], c["mean"])
  kl = O.multivariate_normal_divergence(c["mean"], c["cov"], model, c["dataset"], wf)
  assert abs(kl - c["kl"]) <= 1e-11 * abs(c["kl"])
  euc = O.multivariate_normal_divergence(
      c["mean"], c["cov"], model, c["dataset"], wf,
      distance=O.euclidean_multivariate_normal)
  assert abs(euc - c["euc"]) <= 1e-12 * abs(c["euc"])
  v, g = O.kl_value_and_grad(c["mean"], c["cov"], model, c["dataset"], wf)
  gv = H.grad_vec(g, c["d"])
  if c["mean"] == "zero":
    gv[0] = 0.0
  assert abs(v - c["kl"]) <= 1e-11 * abs(c["kl"])
  assert H.rel(gv, c["kl_grad"]) < 1e-9
