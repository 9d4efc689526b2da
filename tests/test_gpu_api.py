"""The reference-facing Python API on the GPU: ports of the reference's own
tests (gp_test.py, kernel_test.py, acfun_test.py) plus oracle parity through
the public entry points."""
import numpy as np
import pytest
import torch

from hyperbo_b200.basics import definitions as defs
from hyperbo_b200.basics import linalg
from hyperbo_b200.bo_utils import acfun, const
from hyperbo_b200.gp_utils import gp, kernel, mean, objectives, utils
from oracle import hyperbo_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu
WF, WFO = utils.DEFAULT_WARP_FUNC, O.DEFAULT_WARP_FUNC
COVS = {"squared_exponential": kernel.squared_exponential,
        "matern32": kernel.matern32, "matern52": kernel.matern52}


def _np(t):
  return t.detach().cpu().numpy()


@pytest.mark.parametrize("cov", sorted(COVS))
def test_infer_parameters_decreases_nll(cov):  # gp_test.py:48-148
  n = 100
  vx = np.random.default_rng(0).normal(size=(n, 1))
  truth = defs.GPParams(model=dict(O.GROUND_TRUTH))
  dataset = []
  for i in range(10):
    vy = gp.sample_from_gp(i, mean.constant, COVS[cov], truth, vx)
    assert vy.shape == (n, 1)
    dataset.append((vx, vy))
  init = defs.GPParams(
      model={"constant": 5.1, "lengthscale": 0., "signal_variance": 0.,
             "noise_variance": -4.},
      config={"method": "adam", "learning_rate": 1e-5, "beta": 0.9,
              "max_training_step": 1, "logging_interval": 1, "batch_size": 100})
  model = gp.GP(dataset=dataset, mean_func=mean.constant, cov_func=COVS[cov],
                params=init, warp_func=WF)
  model.initialize_params(0)
  dict_ds = {i: defs.SubDataset(*dataset[i]) for i in range(10)}

  def nll_func(p):
    return float(objectives.neg_log_marginal_likelihood(
        mean_func=mean.constant, cov_func=COVS[cov], params=p, dataset=dict_ds,
        warp_func=WF))

  init_nll = nll_func(model.params)
  inferred = model.train()
  assert init_nll > nll_func(inferred)
  assert inferred.cache == {}


@pytest.mark.parametrize("nx", [20, 0])
def test_predict(nx):  # gp_test.py:150-207
  rng = np.random.default_rng(0)
  nq = 10
  vx = rng.normal(size=(nx, 1))
  params = defs.GPParams(model=dict(O.GROUND_TRUTH))
  vy = _np(gp.sample_from_gp(1, mean.constant, kernel.squared_exponential,
                             params, vx)) if nx else np.zeros((0, 1))
  xq = rng.normal(size=(nq, 1))
  model = gp.GP(dataset=[(vx, vy)], mean_func=mean.constant,
                cov_func=kernel.squared_exponential, params=params)
  mu_model, var_model = model.predict(xq, full_cov=False, with_noise=True)
  mu, var = gp.predict(mean.constant, kernel.squared_exponential, params, vx, vy,
                       xq, full_cov=False)
  assert mu.shape == (nq, 1) and var.shape == (nq, 1)
  assert mu_model.shape == (nq, 1) and var_model.shape == (nq, 1)
  assert np.allclose(_np(mu), _np(mu_model), atol=1e-10)
  assert np.allclose(_np(var) + params.model["noise_variance"], _np(var_model),
                     atol=1e-10)
  if nx:
    assert model.params.cache[0].needs_update is False
  mu_, cov = gp.predict(mean.constant, kernel.squared_exponential, params, vx,
                        vy, xq, full_cov=True)
  mu_m, cov_m = model.predict(xq, full_cov=True, with_noise=True)
  assert cov.shape == (nq, nq) and cov_m.shape == (nq, nq)
  assert np.allclose(_np(mu), _np(mu_), atol=1e-10)
  assert np.allclose(np.diag(_np(cov)), _np(var).ravel(), atol=1e-8)
  assert np.allclose(np.diag(_np(cov_m)),
                     _np(var).ravel() + params.model["noise_variance"], atol=1e-8)
  off = ~np.eye(nq, dtype=bool)
  assert np.allclose(_np(cov)[off], _np(cov_m)[off], atol=1e-10)
  # oracle parity of the public predict
  mu_o, var_o = O.predict("constant", "squared_exponential", params.model,
                          vx if nx else None, vy, xq)
  assert H.rel(_np(mu), mu_o) < 1e-6 and H.rel(_np(var), var_o) < 1e-6


def test_update_dataset_and_prior_prediction():  # gp_test.py:209-277
  rng = np.random.default_rng(0)
  vx, vy = rng.normal(size=(20, 2)), rng.normal(size=(20, 1))
  params = defs.GPParams(model={"constant": 5., "lengthscale": 1.,
                                "signal_variance": 1., "noise_variance": .01})
  model = gp.GP(dataset=[(vx, vy)], mean_func=mean.constant,
                cov_func=kernel.squared_exponential, params=params)
  xq = rng.normal(size=(7, 2))
  mu0, _ = model.predict(xq)
  assert 0 in model.params.cache
  model.update_sub_dataset((vx[:3], vy[:3] + 1.0), 0, is_append=True)
  assert model.params.cache[0].needs_update
  mu1, _ = model.predict(xq)
  assert not model.params.cache[0].needs_update
  assert model.params.cache[0].chol.shape == (23, 23)
  assert not np.allclose(_np(mu0), _np(mu1))
  # unknown key -> prior (gp.py:584-593)
  mu_p, var_p = model.predict(xq, sub_dataset_key="nope", with_noise=False)
  assert np.allclose(_np(mu_p), 5.0) and np.allclose(_np(var_p), 1.0)
  # appended-to-new-key with one point
  model.update_sub_dataset((vx[0], vy[0]), "new", is_append=True)
  mu_n, var_n = model.predict(xq, sub_dataset_key="new")
  assert mu_n.shape == (7, 1) and np.all(_np(var_n) > 0)


@pytest.mark.parametrize("cov", sorted(COVS))
def test_kernel_call_signature(cov):  # kernel_test.py:37-89
  rng = np.random.default_rng(0)
  vx1, vx2 = rng.normal(size=(10, 2)), rng.normal(size=(20, 2))
  params = defs.GPParams(model={"lengthscale": np.array([1., 2.]),
                                "signal_variance": 1.5})
  f = COVS[cov]
  assert f(params, vx1, vx2).shape == (10, 20)
  k = _np(f(params, vx1))
  assert k.shape == (10, 10) and np.allclose(k, k.T)
  assert f(params, vx1, diag=True).shape == (10,)
  assert f(params, vx1, vx2, diag=True).shape == (10, 20)  # kernel.py:54-58
  assert H.rel(k, O.cov_matrix(cov, params.model, vx1)) < 1e-13
  kw = _np(f(params, vx1, warp_func=WF))
  assert H.rel(kw, O.cov_matrix(cov, params.model, vx1, warp_func=WFO)) < 1e-13
  with pytest.raises(ValueError):  # params_utils.py:90-94
    f(defs.GPParams(model={"lengthscale": 1.0}), vx1)


def test_solve_gp_linear_system_matches_oracle():
  x, y = O.make_task(5, 90, 3)
  model = O.init_raw_params(3)
  params = defs.GPParams(model=dict(model))
  chol, kinvy, dy = linalg.solve_gp_linear_system(
      mean.constant, kernel.squared_exponential, params, x, y, WF)
  c_ref, a_ref, dy_ref = O.solve_gp_linear_system(
      "constant", "squared_exponential", model, x, y, WFO)
  assert chol.shape == (90, 90) and kinvy.shape == (90, 1)
  assert H.rel(_np(chol), c_ref) < 1e-9 and H.rel(_np(kinvy), a_ref) < 1e-9
  assert H.rel(_np(dy), dy_ref) < 1e-14
  dy2, cov = linalg.compute_delta_y_and_cov(
      mean.constant, kernel.squared_exponential, params, x, y, WF)
  assert H.rel(_np(cov), c_ref @ c_ref.T) < 1e-12


def test_objective_api_and_value_and_grad():
  d = 4
  ds_np = O.make_dataset(5, 60, d, "matern32", ragged_seed=1, ragged_lo=40,
                         ragged_hi=90)
  ds_np[9] = (np.zeros((0, d)), np.zeros((0, 1)))
  ds_np[10] = (ds_np[0][0], ds_np[0][1], "aligned")  # skipped: objectives.py:182
  dataset = {k: defs.SubDataset(*v) for k, v in ds_np.items()}
  model = O.init_raw_params(d)
  params = defs.GPParams(model=dict(model))
  total, key2nll = objectives.neg_log_marginal_likelihood(
      mean.constant, kernel.matern32, params, dataset, WF, return_key2nll=True)
  t_ref, k_ref = O.neg_log_marginal_likelihood(
      "constant", "matern32", model, ds_np, WFO, return_key2nll=True)
  assert abs(float(total) - t_ref) < 1e-10 * abs(t_ref)
  assert set(key2nll) == set(k_ref)
  for k in k_ref:
    assert abs(float(key2nll[k]) - k_ref[k]) < 1e-10 * abs(k_ref[k])
  val, grads = objectives.nll_value_and_grad(mean.constant, kernel.matern32,
                                             params, dataset, WF)
  v_ref, g_ref = O.nll_value_and_grad("constant", "matern32", model, ds_np, WFO)
  assert abs(float(val) - v_ref) < 1e-10 * abs(v_ref)
  assert set(grads) == set(model)
  for k in g_ref:
    assert H.rel(grads[k], g_ref[k]) < 1e-8
  # the SVD branch (objectives.py:157-176) agrees with the Cholesky branch
  # (objectives_test.py:298-301 asserts 2 places; here to conditioning)
  v_svd = objectives.neg_log_marginal_likelihood(
      mean.constant, kernel.matern32, params, dataset, WF, use_cholesky=False)
  assert abs(float(v_svd) - v_ref) < 1e-8 * abs(v_ref)


def test_infer_parameters_matches_oracle_adam_loop():
  d = 2
  ds_np = O.make_dataset(6, 50, d)
  model = O.init_raw_params(d)
  cfg = {"method": "adam", "learning_rate": 1e-2, "max_training_step": 6,
         "batch_size": 1000}
  seen = []
  params = gp.infer_parameters(
      mean.constant, kernel.squared_exponential,
      defs.GPParams(model=dict(model), config=dict(cfg)),
      {k: defs.SubDataset(*v) for k, v in ds_np.items()}, WF,
      callback=lambda i, m, l: seen.append((i, l)))
  ref_model, ref_losses = O.infer_parameters_adam(
      "constant", "squared_exponential", model, ds_np, WFO, 1e-2, 6, 1000)
  assert [i for i, _ in seen] == list(range(6))
  assert H.rel([l for _, l in seen], ref_losses) < 1e-9
  for k in ref_model:
    assert H.rel(params.model[k], ref_model[k]) < 1e-8
  assert isinstance(params.model["constant"], float)
  assert np.asarray(params.model["lengthscale"]).shape == (d,)


def test_infer_parameters_subsampling_and_nan_at_step0():
  d = 1
  ds_np = O.make_dataset(3, 40, d)
  cfg = {"method": "adam", "learning_rate": 1e-3, "max_training_step": 3,
         "batch_size": 16}  # n >= batch_size -> per-step sub-sampling
  losses = []
  out = gp.infer_parameters(
      mean.constant, kernel.matern52,
      defs.GPParams(model=dict(O.init_raw_params(d)), config=cfg),
      {k: defs.SubDataset(*v) for k, v in ds_np.items()}, WF, key=3,
      callback=lambda i, m, l: losses.append(l))
  assert len(losses) == 3 and len(set(losses)) == 3 and np.all(np.isfinite(losses))
  bad = {0: defs.SubDataset(np.full((20, 1), 0.5), np.ones((20, 1)))}
  with pytest.raises(ValueError, match="NaN"):  # gp.py:135-137
    gp.infer_parameters(
        mean.constant, kernel.squared_exponential,
        defs.GPParams(model={"constant": 0., "lengthscale": np.array([1.0]),
                             "signal_variance": 1.0, "noise_variance": -1e-6},
                      config=dict(cfg, batch_size=100)), bad, None)


@pytest.mark.parametrize("name", sorted(const.ACFUN))
def test_acquisition_shape(name):  # acfun_test.py:43-72
  ds_np = O.make_dataset(3, 30, 2)
  model = gp.GP({k: defs.SubDataset(*v) for k, v in ds_np.items()},
                mean.constant, kernel.matern52,
                defs.GPParams(model=dict(O.init_raw_params(2))), WF)
  model.rng = 0
  xq = np.random.default_rng(0).random((25, 2))
  f = const.ACFUN[name]
  out = f(model=model, sub_dataset_key=1, x_queries=xq) if name != "random_search" \
      else f(model, xq)
  assert out.shape == (25, 1)


def test_acquisition_values_match_oracle():
  d = 4
  ds_np = O.make_dataset(4, 120, d, "matern52", ragged_seed=7, ragged_lo=90,
                         ragged_hi=140)
  m = O.init_raw_params(d)
  model = gp.GP({k: defs.SubDataset(*v) for k, v in ds_np.items()},
                mean.constant, kernel.matern52,
                defs.GPParams(model=dict(m)), WF)
  xq = np.random.default_rng(9).random((1000, d))
  for name, f in (("ei", acfun.ei), ("pi", acfun.pi), ("pi2", acfun.pi2),
                  ("pi3", acfun.pi3), ("ucb", acfun.ucb), ("ucb2", acfun.ucb2),
                  ("ucb4", acfun.ucb4)):
    got = _np(f(model=model, sub_dataset_key=2, x_queries=xq))
    want = O.acquisition(name, "constant", "matern52", m, ds_np, 2, xq, WFO)
    assert H.rel(got, want) < 1e-6, name
  # no observations for the key: prior + target 0.0 (acfun.py:145-148)
  got = _np(acfun.ei(model=model, sub_dataset_key="none", x_queries=xq))
  want = O.acquisition("ei", "constant", "matern52", m, ds_np, "none", xq, WFO)
  assert H.rel(got, want) < 1e-6
  # the *_sub functions on explicit vectors
  mu, var = model.predict(xq, sub_dataset_key=2)
  got = _np(acfun.expected_improvement_sub(mu, torch.sqrt(var), 0.3))
  want = O.expected_improvement_sub(_np(mu), np.sqrt(_np(var)), 0.3)
  assert H.rel(got, want) < 1e-9


def test_simulated_bo_loop_shape():  # bayesopt.py:137-193 caller pattern
  d = 2
  ds_np = O.make_dataset(3, 25, d)
  model = gp.GP({k: defs.SubDataset(*v) for k, v in ds_np.items()},
                mean.constant, kernel.squared_exponential,
                defs.GPParams(model=dict(O.init_raw_params(d))), WF)
  xq, yq = O.make_task(99, 200, d)
  for it in range(3):
    evals = acfun.ucb(model=model, sub_dataset_key=0, x_queries=xq)
    idx = int(evals.argmax())
    model.update_sub_dataset((xq[idx], yq[idx]), 0, is_append=True)
  assert model.dataset[0].x.shape == (28, d)


def test_adam_trainer_host_batches_double_buffered():
  """AdamTrainer.step_from_host uploads every step's batch on a copy stream
  into one of two device buffers (overlapping the previous step's kernels):
  with a DIFFERENT batch per step the losses must equal the same steps taken on
  device-resident copies of those batches."""
  from hyperbo_b200.engine import Engine, PackedDataset
  eng = Engine.get()
  d, n, T = 3, 70, 5
  rng = np.random.default_rng(11)
  offs = [n * t for t in range(T + 1)]
  batches = [(torch.from_numpy(rng.random((T * n, d))).pin_memory(),
              torch.from_numpy(rng.standard_normal(T * n)).pin_memory())
             for _ in range(5)]
  raw0 = np.array([0.1, 0.2, -2.0, 0.3, -0.2, 0.1])
  mask = H.default_mask(d)

  def run(from_host, use_graph):
    tr = gp.AdamTrainer(eng, 2, 1, raw0, mask, d, 1e-2)
    ds = PackedDataset(list(range(T)), batches[0][0].to(eng.device),
                       batches[0][1].to(eng.device), offs)
    losses = []
    for xh, yh in batches:
      if from_host:
        prev = tr.step_pipelined(ds, xh, yh, use_graph=use_graph)
        if prev is not None:
          losses.append(prev)
      else:
        dsk = PackedDataset(list(range(T)), xh.to(eng.device), yh.to(eng.device),
                            offs)
        tr.step(dsk)
        losses.append(tr.loss())
    if from_host:
      losses.append(tr.flush())
    return np.array(losses), _np(tr.raw)

  l_dev, r_dev = run(False, False)
  for use_graph in (False, True):
    l_host, r_host = run(True, use_graph)
    assert np.array_equal(l_dev, l_host), use_graph
    assert np.array_equal(r_dev, r_host), use_graph
  # and against the oracle for the first batch
  ds_np = {t: (batches[0][0].numpy()[n * t:n * (t + 1)],
               batches[0][1].numpy()[n * t:n * (t + 1), None]) for t in range(T)}
  v_ref = O.neg_log_marginal_likelihood(
      "constant", "matern52", H.model_from_raw(raw0, d, "constant"), ds_np, WFO)
  assert abs(l_dev[0] - v_ref) < 1e-10 * abs(v_ref)


@pytest.mark.parametrize("name", ["expected_improvement", "ucb",
                                  "probability_of_improvement"])
def test_acfun_over_hyperparameter_sets(name):
  """acfun_test.py:74-118 evaluates an acquisition for 100 (constant,
  lengthscale) sets with jax.vmap.  The engine cannot be traced by vmap; the
  same sweep is a host loop over update_model_params (each set = one factorise
  + one fused predict/acquisition launch sequence)."""
  rng = np.random.default_rng(0)
  nx, nq, dim, S = 20, 10, 5, 12
  vx = rng.normal(size=(nx, dim))
  truth = defs.GPParams(model={"constant": 5., "lengthscale": 0.1,
                               "signal_variance": 1.0, "noise_variance": 0.01})
  vy = gp.sample_from_gp(3, mean.constant, kernel.squared_exponential, truth, vx)
  xq = rng.normal(size=(nq, dim))
  sets = [np.hstack([rng.uniform(-10., 10.), rng.gamma(1., 1., (dim,)) + 0.05])
          for _ in range(S)]
  model = gp.GP(dataset=[(vx, vy)], mean_func=mean.constant,
                cov_func=kernel.squared_exponential,
                params=defs.GPParams(model=dict(truth.model)))
  f = const.ACFUN[name]
  evals = []
  for cl in sets:
    model.update_model_params({"constant": float(cl[0]), "lengthscale": cl[1:],
                               "signal_variance": 1.0, "noise_variance": 0.01})
    evals.append(f(model=model, sub_dataset_key=0, x_queries=xq))
  evals = torch.stack(evals)
  assert evals.shape == (S, nq, 1)
  ds_np = {0: (vx, _np(vy))}
  oname = {"expected_improvement": "ei", "ucb": "ucb",
           "probability_of_improvement": "pi"}[name]
  for s in (0, S - 1):
    m = {"constant": float(sets[s][0]), "lengthscale": sets[s][1:],
         "signal_variance": 1.0, "noise_variance": 0.01}
    want = O.acquisition(oname, "constant", "squared_exponential", m, ds_np, 0,
                         xq, None)
    assert np.max(np.abs(_np(evals[s]) - want)) < 1e-6 * (
        np.max(np.abs(want)) + 1e-3), (name, s)


def test_sample_from_gp_uses_the_engine_factor():
  """gp.py:198-239: mean + chol(K + (noise + 1e-6) I) z -- the factor now comes
  from hb_factorize_batched; same draw as the explicit-matrix Cholesky."""
  truth = defs.GPParams(model={"constant": 5.0, "lengthscale": 1.0,
                               "signal_variance": 1.0, "noise_variance": 0.01})
  vx = np.random.default_rng(5).random((150, 3))
  got = _np(gp.sample_from_gp(7, mean.constant, kernel.matern52, truth, vx, num_samples=4))
  _, cov = linalg.compute_delta_y_and_cov(mean.constant, kernel.matern52, truth, vx,
                                          torch.zeros((150, 1)), None, 1e-6)
  z = torch.randn((150, 4), generator=torch.Generator().manual_seed(7), dtype=torch.float64)
  want = 5.0 + _np(torch.linalg.cholesky(cov)) @ z.numpy()
  assert got.shape == (150, 4)
  assert np.max(np.abs(got - want)) < 1e-9 * np.max(np.abs(want))


@pytest.mark.parametrize("method", ["adam", "lbfgs"])
def test_train_writes_a_checkpoint(tmp_path, method):
  """GP.train(get_params_path=...) (gp.py:151-157,186-191): the final parameters,
  step and loss are pickled; load_params restores them."""
  from hyperbo_b200.basics import params_utils
  ds = {t: defs.SubDataset(*O.make_task(t, 30, 2, "matern32")) for t in range(3)}
  params = defs.GPParams(
      model=dict(O.init_raw_params(2)),
      config={"method": method, "learning_rate": 1e-2, "max_training_step": 5,
              "batch_size": 100, "alpha": 1.0})
  model = gp.GP(dataset=ds, mean_func=mean.constant, cov_func=kernel.matern32,
                params=params, warp_func=WF)
  path = str(tmp_path / f"{method}.pkl")
  out = model.train(get_params_path=lambda: path)
  loaded, (step, loss) = params_utils.load_params(path, include_state=True)
  assert step == 5 and np.isfinite(loss)
  assert set(loaded.model) == set(out.model)
  for k in out.model:
    assert np.allclose(np.asarray(loaded.model[k], dtype=np.float64),
                       _np(torch.as_tensor(out.model[k])) if not np.isscalar(out.model[k])
                       else out.model[k])
