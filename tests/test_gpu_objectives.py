"""The reference's other objectives on the engine:
multivariate_normal_divergence (kl / ekl / regkl, euc), the add / mul
combinations, the SVD NLL branch and GP.stats() -- parity with the CPU oracle
and the committed fixtures, through the public API."""
import functools

import numpy as np
import pytest
import torch

from hyperbo_b200.basics import definitions as defs
from hyperbo_b200.gp_utils import gp, kernel, mean, objectives, utils
from oracle import hyperbo_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu
WF, WFO = utils.DEFAULT_WARP_FUNC, O.DEFAULT_WARP_FUNC
COVS = {"squared_exponential": kernel.squared_exponential,
        "matern32": kernel.matern32, "matern52": kernel.matern52}
MEANS = {"constant": mean.constant, "zero": mean.zero}


def _case(name):
  g = H.load_golden_kl(name)
  model = H.model_from_raw(g["raw"], g["d"], g["mean"])
  params = defs.GPParams(model=dict(model))
  dataset = {k: defs.SubDataset(*v) for k, v in g["dataset"].items()}
  return g, model, params, dataset, MEANS[g["mean"]], COVS[g["cov"]]


@pytest.mark.parametrize("name", H.golden_cases(kl=True))
def test_divergence_values_match_golden(name):
  g, model, params, dataset, mf, cf = _case(name)
  kl = utils.kl_multivariate_normal
  v = objectives.multivariate_normal_divergence(mf, cf, params, dataset, WF)
  assert abs(float(v) - g["kl"]) < 1e-10 * abs(g["kl"])  # fp64 engine vs oracle
  v = objectives.ekl(mf, cf, params, dataset, WF,
                     distance=functools.partial(kl, eps=1e-6))
  assert abs(float(v) - g["kl_eps"]) < 1e-10 * abs(g["kl_eps"])
  v = objectives.kl(mf, cf, params, dataset, WF,
                    distance=functools.partial(kl, eps=1e-6, partial=False))
  assert abs(float(v) - g["kl_full"]) < 1e-7 * abs(g["kl_full"])
  v = objectives.euc(mf, cf, params, dataset, WF)
  assert abs(float(v) - g["euc"]) < 1e-10 * abs(g["euc"])
  v = objectives.kl(mf, cf, params, dataset, WF,
                    distance=functools.partial(kl, weight=0.25))
  assert abs(float(v) - 0.25 * g["kl"]) < 1e-10 * abs(g["kl"])


@pytest.mark.parametrize("name", H.golden_cases(kl=True))
def test_kl_value_and_grad_match_golden(name):
  g, model, params, dataset, mf, cf = _case(name)
  val, grads = objectives.value_and_grad(objectives.kl, mf, cf, params, dataset,
                                         WF)
  assert abs(float(val) - g["kl"]) < 1e-10 * abs(g["kl"])
  assert set(grads) == set(model)
  assert H.rel(H.grad_vec(grads, g["d"]), g["kl_grad"]) < 1e-8


@pytest.mark.parametrize("cov", sorted(COVS))
def test_nll_plus_regkl_value_and_grad(cov):
  """objectives.nll_regkl(c) = nll + c * regkl (objectives.py:239-247) on a
  dataset mixing ordinary and aligned sub-datasets."""
  g, model, params, dataset, mf, cf = _case("kl_m52_const_d3")
  cf = COVS[cov]
  ds_np = g["dataset"]
  for c, objective in ((0.1, objectives.nll_regkl01), (10.0, objectives.nll_regkl10),
                       (2.5, objectives.add(objectives.nll,
                                            objectives.mul(2.5, objectives.kl)))):
    val, grads = objectives.value_and_grad(objective, mf, cf, params, dataset, WF)
    v_nll, g_nll = O.nll_value_and_grad("constant", cov, model, ds_np, WFO)
    v_kl, g_kl = O.kl_value_and_grad("constant", cov, model, ds_np, WFO)
    ref = v_nll + c * v_kl
    assert abs(float(val) - ref) < 1e-10 * abs(ref)
    # the callable itself evaluates the same number
    assert abs(float(objective(mf, cf, params, dataset, WF)) - ref) < 1e-10 * abs(ref)
    for k in g_nll:
      want = np.asarray(g_nll[k]) + c * np.asarray(g_kl[k])
      assert H.rel(grads[k], want) < 1e-8, (k, c)


def test_scalar_lengthscale_and_no_warp():
  g, model, params, dataset, mf, cf = _case("kl_se_zero_d2")
  model = {"lengthscale": 0.7, "signal_variance": 1.3, "noise_variance": 0.05}
  params = defs.GPParams(model=dict(model))
  val, grads = objectives.value_and_grad(objectives.ekl, mf, cf, params, dataset,
                                         None)
  v_ref, g_ref = O.kl_value_and_grad("zero", g["cov"], model, g["dataset"], None)
  assert abs(float(val) - v_ref) < 1e-10 * abs(v_ref)
  for k in g_ref:
    assert H.rel(grads[k], g_ref[k]) < 1e-8, k


def test_divergence_edge_cases():
  g, model, params, dataset, mf, cf = _case("kl_m52_const_d3")
  # no aligned sub-dataset -> 0 (objectives.py:99-100)
  plain = {k: s for k, s in dataset.items() if s.aligned is None}
  assert float(objectives.kl(mf, cf, params, plain, WF)) == 0.0
  # mismatched rows raise like the reference (objectives.py:91-96)
  bad = dict(dataset)
  bad[0] = defs.SubDataset(dataset[0].x, dataset[0].y[:-1], 1)
  with pytest.raises(ValueError):
    objectives.kl(mf, cf, params, bad, WF)
  # value-only distances have no gradient program
  with pytest.raises(NotImplementedError):
    objectives.value_and_grad(
        functools.partial(objectives.kl, distance=functools.partial(
            utils.kl_multivariate_normal, eps=1e-6)), mf, cf, params, dataset, WF)


@pytest.mark.parametrize("method", ["adam", "lbfgs"])
def test_training_on_ekl_decreases_it(method):
  """gp_test.py:48-148 with objective = ekl (the paper's second objective)."""
  g, model, params, dataset, mf, cf = _case("kl_m52_const_d3")
  params.config = {"method": method, "learning_rate": 1e-2, "beta": 0.9,
                   "max_training_step": 30 if method == "adam" else 5,
                   "batch_size": 1000, "objective": objectives.ekl, "alpha": 1.0}
  before = float(objectives.ekl(mf, cf, params, dataset, WF))
  model_ = gp.GP(dataset=dataset, mean_func=mf, cov_func=cf, params=params,
                 warp_func=WF)
  seen = []
  out = model_.train(callback=(lambda i, p, l: seen.append(l))
                     if method == "adam" else None)
  after = float(objectives.ekl(mf, cf, out, dataset, WF))
  assert after < before
  if method == "adam":
    assert len(seen) == 30 and abs(seen[0] - before) < 1e-9 * abs(before)
    # the Adam trajectory equals the oracle's (optax.adam restated) on the
    # oracle's closed-form KL gradient
    opt = O.Adam(1e-2)
    m = dict(model)
    for _ in range(30):
      _, gr = O.kl_value_and_grad(g["mean"], g["cov"], m, g["dataset"], WFO)
      m = opt.update(m, gr)
    for k in m:
      assert H.rel(out.model[k], m[k]) < 1e-7, k


def test_gp_stats():  # gp.py:511-533
  g, model, params, dataset, mf, cf = _case("kl_m52_const_d3")
  model_ = gp.GP(dataset=dataset, mean_func=mf, cov_func=cf, params=params,
                 warp_func=WF)
  nll, ekl, ekl_partial, euc, key2nll = model_.stats(verbose=False)
  v_nll, k2n = O.neg_log_marginal_likelihood(
      g["mean"], g["cov"], model, g["dataset"], WFO, return_key2nll=True,
      use_cholesky=False)
  assert abs(nll - v_nll) < 1e-8 * abs(v_nll)
  assert set(key2nll) == set(k2n)
  assert abs(ekl - g["kl_full"]) < 1e-7 * abs(g["kl_full"])
  assert abs(ekl_partial - g["kl_eps"]) < 1e-10 * abs(g["kl_eps"])
  assert abs(euc - g["euc"]) < 1e-10 * abs(g["euc"])


def test_weighted_call_and_jitter():
  """hb_nll_grad_weighted: sums are the w-weighted per-task values; `jitter`
  replaces the 1e-6 of linalg.py:42 (checked against the oracle by moving the
  difference into an un-warped noise variance)."""
  from hyperbo_b200.engine import Engine
  eng = Engine.get()
  d, ns = 3, [70, 20, 129]
  ds_np = {t: O.make_task(t, n, d, "matern32") for t, n in enumerate(ns)}
  ds = eng.pack([(t, x, y) for t, (x, y) in ds_np.items()])
  model = {"constant": 0.4, "signal_variance": 1.2, "noise_variance": 0.03,
           "lengthscale": np.array([0.6, 0.9, 1.4])}
  raw = H.raw_vec(model, d)
  w = np.array([0.5, -2.0, 3.0])
  jit = 1e-3
  sums = eng.nll_grad(1, 1, ds, raw, 0, weights=w, jitter=jit).cpu().numpy()
  shifted = dict(model)
  shifted["noise_variance"] = model["noise_variance"] + jit - O.JITTER
  val, grad = 0.0, np.zeros(3 + d)
  for t in range(3):
    v, g = O.nll_and_grad_sub_dataset("constant", "matern32", shifted,
                                      *ds_np[t], warp_func=None)
    val += w[t] * v
    grad += w[t] * H.grad_vec(g, d)
  assert abs(sums[0] - val) < 1e-10 * abs(val)
  assert H.rel(sums[1:-1], grad) < 1e-8
  assert sums[-1] == 3.0  # the task count stays unweighted
  # weights = None, default jitter == hb_nll_grad_batched bit for bit
  a = eng.nll_grad(1, 1, ds, raw, 0).cpu().numpy()
  b = eng.nll_grad(1, 1, ds, raw, 0, weights=np.ones(3), jitter=O.JITTER).cpu().numpy()
  assert np.array_equal(a, b)


@pytest.mark.parametrize("cov", sorted(COVS))
def test_sample_mean_cov_regularizer(cov):
  """Port of objectives_test.py:67-196 (kl distance, lbfgs; the engine kernels):
  10 GP samples on shared inputs form ONE aligned sub-dataset; training on the
  divergence decreases it; the SVD NLL agrees with the Cholesky NLL (evaluated
  with exclude_aligned=False on the (n, 10) y, objectives.py:153-155)."""
  n = 20
  vx = np.random.default_rng(0).normal(size=(n, 2))
  truth = defs.GPParams(model={"constant": 5., "lengthscale": 1.,
                               "signal_variance": 1.0, "noise_variance": 0.01})
  cf = COVS[cov]
  vy = gp.sample_from_gp(1, mean.constant, cf, truth, vx, num_samples=10)
  assert vy.shape == (n, 10)
  dataset = [(vx, vy, "all_data")]
  distance = utils.kl_multivariate_normal
  init = defs.GPParams(
      model={"constant": 5.1, "lengthscale": 0., "signal_variance": 0.,
             "noise_variance": -4.},
      config={"method": "lbfgs", "max_training_step": 2, "logging_interval": 1,
              "objective": functools.partial(
                  objectives.multivariate_normal_divergence, distance=distance),
              "batch_size": 100, "learning_rate": 0.001})
  model = gp.GP(dataset=dataset, mean_func=mean.constant, cov_func=cf,
                params=init, warp_func=WF)

  def reg(p, wf=None):
    return float(objectives.multivariate_normal_divergence(
        mean_func=model.mean_func, cov_func=model.cov_func, params=p,
        dataset=model.dataset, warp_func=wf, distance=distance))

  def nll_func(p, wf=None, use_cholesky=True):
    return float(objectives.neg_log_marginal_likelihood(
        mean_func=model.mean_func, cov_func=model.cov_func, params=p,
        dataset=model.dataset, warp_func=wf, use_cholesky=use_cholesky,
        exclude_aligned=False))

  assert np.isfinite(reg(truth)) and np.isfinite(nll_func(truth))
  init_reg = reg(init, WF)
  init_nll = nll_func(init, WF)
  assert abs(nll_func(init, WF, use_cholesky=False) / init_nll - 1.) < 5e-3
  # the multi-column value equals the oracle's literal restatement
  ds_np = {0: (vx, vy.cpu().numpy(), 0)}
  want = O.neg_log_marginal_likelihood("constant", cov, dict(init.model), ds_np,
                                       WFO, exclude_aligned=False)
  assert abs(init_nll - want) < 1e-9 * abs(want)
  inferred = model.train()
  inferred_reg = reg(inferred, WF)
  inferred_nll = nll_func(inferred, WF)
  assert abs(nll_func(inferred, WF, use_cholesky=False) / inferred_nll - 1.) < 5e-3
  assert init_reg > inferred_reg


def test_hgp_stats_average_over_samples():  # gp.py:634-664
  g, model, params, dataset, mf, cf = _case("kl_m52_const_d3")
  other = dict(model)
  other["signal_variance"] = model["signal_variance"] + 0.3
  single = [gp.GP(dataset=dataset, mean_func=mf, cov_func=cf,
                  params=defs.GPParams(model=dict(m)), warp_func=WF
                  ).stats(verbose=False) for m in (model, other)]
  hgp = gp.HGP(dataset=dataset, mean_func=mf, cov_func=cf,
               params=defs.GPParams(model=dict(model),
                                    samples=[dict(model), dict(other)]),
               warp_func=WF)
  nll, ekl, ekl_partial, euc, key2nll = hgp.stats(verbose=False)
  for got, i in ((nll, 0), (ekl, 1), (ekl_partial, 2), (euc, 3)):
    want = 0.5 * (single[0][i] + single[1][i])
    assert abs(got - want) < 1e-12 * abs(want)
  for k in key2nll:
    assert abs(key2nll[k] - 0.5 * (single[0][4][k] + single[1][4][k])) < 1e-9
