"""L-BFGS / BFGS drivers (host logic, no GPU) + their use through
gp.infer_parameters and the BO drivers on the GPU."""
import math

import numpy as np
import pytest
import torch

from hyperbo_b200.basics import bfgs, lbfgs
from hyperbo_b200.basics import definitions as defs
from oracle import hyperbo_oracle as O
from tests import helpers as H


def rosen(x):
  v = 100 * (x[1] - x[0]**2)**2 + (1 - x[0])**2
  g = np.array([-400 * x[0] * (x[1] - x[0]**2) - 2 * (1 - x[0]),
                200 * (x[1] - x[0]**2)])
  return float(v), g


def test_lbfgs_rosenbrock_and_resume():
  v, x, state = lbfgs.lbfgs(rosen, np.array([-1.2, 1.0]), steps=200, tol=1e-16)
  assert v < 1e-14 and np.allclose(x, 1.0, atol=1e-6)
  # resume from a saved state after a few steps (lbfgs.py:285-286)
  v1, x1, st = lbfgs.lbfgs(rosen, np.array([-1.2, 1.0]), steps=5)
  v2, x2, _ = lbfgs.lbfgs(rosen, x1, steps=200, tol=1e-16, state=st)
  assert v2 < v1 and v2 < 1e-12


def test_lbfgs_two_loop_matches_bfgs_inverse_on_quadratic():
  rng = np.random.default_rng(0)
  a = rng.normal(size=(5, 5))
  hess = a @ a.T + 5 * np.eye(5)
  s = [rng.normal(size=5) for _ in range(3)]
  y = [hess @ si for si in s]
  g = rng.normal(size=5)
  d = lbfgs.descent_direction(g, s, y)
  # explicit BFGS inverse-Hessian recursion from H0 = gamma I
  hk = (s[-1] @ y[-1]) / (y[-1] @ y[-1]) * np.eye(5)
  for si, yi in zip(s, y):
    rho = 1.0 / (yi @ si)
    v = np.eye(5) - rho * np.outer(si, yi)
    hk = v @ hk @ v.T + rho * np.outer(si, si)
  assert np.allclose(d, -hk @ g, rtol=1e-10)


def test_linesearch_conditions_and_nan_handling():
  f = lambda x: (float(x @ x), 2 * x)
  x = np.array([1.0, -2.0])
  v, g = f(x)
  new_v, step = lbfgs.backtracking_linesearch(f, v, x, g, -g, alpha=1.0)
  assert new_v < v and step > 0
  # ascent direction: no progress is reported
  assert lbfgs.backtracking_linesearch(f, v, x, g, g) == (v, 0.0)
  # objective that is NaN everywhere but the start: return where we started
  nanf = lambda z: (float("nan"), np.full_like(z, np.nan))
  assert lbfgs.backtracking_linesearch(nanf, v, x, g, -g, max_steps=5) == (v, 0.0)
  # converged at start
  v0, x0, st = lbfgs.lbfgs(f, np.zeros(2))
  assert v0 == 0.0 and st is None


def test_bfgs_driver():
  x, v = bfgs.bfgs(rosen, np.array([-1.2, 1.0]), max_training_step=200)
  assert v < 1e-10 and np.allclose(x, 1.0, atol=1e-4)


# ------------------------------------------------------------------- GPU ---
def _oracle_val_and_grad(model_keys_d, ds_np, cov, scalar_ls):
  d = model_keys_d

  def vg(v):
    model = {"constant": v[0], "signal_variance": v[1], "noise_variance": v[2],
             "lengthscale": float(v[3]) if scalar_ls else np.array(v[3:])}
    val, g = O.nll_value_and_grad("constant", cov, model, ds_np,
                                  O.DEFAULT_WARP_FUNC)
    ls = [float(g["lengthscale"])] if scalar_ls else list(g["lengthscale"])
    return val, np.array([g["constant"], g["signal_variance"],
                          g["noise_variance"]] + ls)

  return vg


@pytest.mark.gpu
@pytest.mark.parametrize("scalar_ls", [False, True])
def test_infer_parameters_lbfgs_matches_oracle_driven_run(scalar_ls):
  from hyperbo_b200.gp_utils import gp, kernel, mean, utils
  d = 3
  ds_np = O.make_dataset(4, 45, d, "matern52")
  model = O.init_raw_params(d)
  if scalar_ls:
    model["lengthscale"] = 0.0
  cfg = {"method": "lbfgs", "max_training_step": 6, "batch_size": 1000}
  seen = []
  out = gp.infer_parameters(
      mean.constant, kernel.matern52,
      defs.GPParams(model=dict(model), config=cfg),
      {k: defs.SubDataset(*v) for k, v in ds_np.items()},
      utils.DEFAULT_WARP_FUNC, callback=lambda i, m, l: seen.append(l))
  v0 = np.array([model["constant"], model["signal_variance"],
                 model["noise_variance"]] +
                ([0.0] if scalar_ls else [0.0] * d))
  ref_losses = []
  _, x_ref, _ = lbfgs.lbfgs(
      _oracle_val_and_grad(d, ds_np, "matern52", scalar_ls), v0, steps=6,
      callback=lambda step, model_params, loss: ref_losses.append(loss))
  assert len(seen) == len(ref_losses) and H.rel(seen, ref_losses) < 1e-8
  got = np.concatenate([[out.model["constant"], out.model["signal_variance"],
                         out.model["noise_variance"]],
                        np.atleast_1d(out.model["lengthscale"])])
  assert H.rel(got, x_ref) < 1e-6
  assert seen[-1] < seen[0]
  assert np.ndim(out.model["lengthscale"]) == (0 if scalar_ls else 1)


@pytest.mark.gpu
def test_infer_parameters_bfgs_decreases_nll():
  from hyperbo_b200.gp_utils import gp, kernel, mean, objectives, utils
  ds_np = O.make_dataset(3, 40, 2)
  dataset = {k: defs.SubDataset(*v) for k, v in ds_np.items()}
  p0 = defs.GPParams(model=dict(O.init_raw_params(2)),
                     config={"method": "bfgs", "max_training_step": 5,
                             "batch_size": 1000, "tol": 1e-8})
  before = float(objectives.nll(mean.constant, kernel.squared_exponential, p0,
                                dataset, utils.DEFAULT_WARP_FUNC))
  out = gp.infer_parameters(mean.constant, kernel.squared_exponential, p0,
                            dataset, utils.DEFAULT_WARP_FUNC)
  after = float(objectives.nll(mean.constant, kernel.squared_exponential, out,
                               dataset, utils.DEFAULT_WARP_FUNC))
  assert after < before


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["expected_improvement",
                                  "probability_of_improvement", "ucb",
                                  "random_search"])
def test_run_bayesopt_synthetic(name):  # bayesopt_test.py:45-103
  from hyperbo_b200.bo_utils import bayesopt, const, data
  from hyperbo_b200.gp_utils import kernel, mean
  params = defs.GPParams(
      model={"constant": 5., "lengthscale": 1., "signal_variance": 1.0,
             "noise_variance": 0.01},
      config={"method": "adam", "learning_rate": 1e-5, "beta": 0.9,
              "max_training_step": 1})
  dataset, key, queried = data.random(
      key=0, mean_func=mean.constant, cov_func=kernel.squared_exponential,
      params=params, dim=5, n_observed=0, n_queries=30, n_func_historical=2,
      m_points_historical=10)
  assert len(dataset) == 3
  for i in range(2):
    assert dataset[i].x.shape == (10, 5) and dataset[i].y.shape == (10, 1)
    assert dataset[i].aligned is None
  obs, best, out_params = bayesopt.run_bayesopt(
      dataset=dataset, sub_dataset_key=key, queried_sub_dataset=queried,
      mean_func=mean.constant, cov_func=kernel.squared_exponential,
      init_params=params, ac_func=const.ACFUN[name], iters=3,
      init_random_key=0)
  assert obs[0].shape == (3, 5) and obs[1].shape == (3, 1)
  assert best[0].shape == (5,)
  assert float(best[1]) == float(queried.y.max())


@pytest.mark.gpu
def test_simulated_bo_picks_the_oracle_argmax():
  """One BO iteration = fused predict + EI over all candidates + arg-max; the
  selected candidate must be the oracle's arg-max."""
  from hyperbo_b200.bo_utils import acfun, bayesopt
  from hyperbo_b200.gp_utils import gp, kernel, mean, utils
  d = 3
  ds_np = O.make_dataset(3, 60, d, "matern32")
  m = O.init_raw_params(d)
  model = gp.GP({k: defs.SubDataset(*v) for k, v in ds_np.items()},
                mean.constant, kernel.matern32, defs.GPParams(model=dict(m)),
                utils.DEFAULT_WARP_FUNC)
  xq, yq = O.make_task(50, 400, d, "matern32")
  out = bayesopt.simulated_bayesopt(model, 1, defs.SubDataset(xq, yq),
                                    acfun.ei, iters=1)
  ei_ref = O.acquisition("ei", "constant", "matern32", m, ds_np, 1, xq,
                         O.DEFAULT_WARP_FUNC)
  assert out.x.shape == (61, d)
  assert np.allclose(out.x[-1].cpu().numpy(), xq[int(np.argmax(ei_ref))])


def test_lbfgs_memo_and_speculative_line_search_keep_the_trajectory():
  """The evaluator behind the driver (basics/lbfgs.py::_Evaluator): the re-evaluation
  of an accepted point is served from the memo, and with a multi-point objective
  the line search needs about half the engine round trips -- same iterates."""
  calls = {"single": 0, "multi": 0}

  def fn(x):
    calls["single"] += 1
    return rosen(x)

  def multi(xs):
    calls["multi"] += 1
    return [rosen(x) for x in xs]

  x0 = np.array([-1.2, 1.0])
  ref_calls = {"n": 0}

  def fn_ref(x):
    ref_calls["n"] += 1
    return rosen(x)

  # reference trajectory: the plain driver without memo (private entry)
  v_ref, x_ref, _ = lbfgs._lbfgs(fn_ref, x0, 10, 50, 60, 1.0, 1e-16, 0.5, None, None)
  st_a, st_b = {}, {}
  v_a, x_a, _ = lbfgs.lbfgs(fn, x0, steps=60, tol=1e-16, stats=st_a)
  n_single = calls["single"]
  v_b, x_b, _ = lbfgs.lbfgs(fn, x0, steps=60, tol=1e-16, multi_fn=multi, stats=st_b)
  assert np.array_equal(x_a, x_ref) and v_a == v_ref     # memo: identical iterates
  assert np.array_equal(x_b, x_ref) and v_b == v_ref     # speculation: identical too
  assert st_a["calls"] == n_single < ref_calls["n"]      # accepted points not re-run
  assert st_b["calls"] < st_a["calls"]                   # fewer round trips
  assert st_b["points"] >= st_a["points"]                # (some speculated points unused)
