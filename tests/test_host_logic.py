"""Host-side logic and the C-ABI boundary, no GPU needed."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from hyperbo_b200.basics import data_utils
from hyperbo_b200.basics import definitions as defs
from hyperbo_b200.basics import params_utils
from hyperbo_b200.bo_utils import const
from hyperbo_b200.gp_utils import gp, kernel, mean, objectives, utils

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_library_exports_every_declared_symbol():
  import __graft_entry__ as ge
  ge.build()
  hdr = open(os.path.join(ROOT, "include", "hyperbo_b200.h")).read()
  names = sorted(set(re.findall(r"\b(hb_[a-z_]+)\s*\(", hdr)))
  assert len(names) >= 12, names
  lib = ctypes.CDLL(os.path.join(ROOT, "hyperbo_b200", "libhyperbo_b200.so"))
  for n in names:
    assert hasattr(lib, n), f"{n} declared in include/ but not exported"
  lib.hb_version.restype = ctypes.c_char_p
  assert b"sm_100a" in lib.hb_version()


def test_cabi_rejects_misuse_without_touching_a_gpu():
  lib = ctypes.CDLL(os.path.join(ROOT, "hyperbo_b200", "libhyperbo_b200.so"))
  assert lib.hb_create(None, 0, 0) == 1  # HB_ERR_BAD_ARG
  h = ctypes.c_void_p()
  rc = lib.hb_create(ctypes.byref(h), 0, 7)
  assert rc == 1
  assert lib.hb_destroy(None) == 1
  lib.hb_predictor_bytes.restype = ctypes.c_int64
  lib.hb_predictor_bytes.argtypes = [ctypes.c_void_p, ctypes.c_int64]
  # 130 points -> 3 blocks -> 6 tiles of 32 KiB + 3*64 alpha + header pad
  assert lib.hb_predictor_bytes(None, 130) == (6 * 4096 + 192) * 8 + 256


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only check")
def test_no_cpu_fallback():
  from hyperbo_b200 import engine
  with pytest.raises(RuntimeError, match="CUDA"):
    engine.Engine.get()
  params = defs.GPParams(model={"lengthscale": 1.0, "signal_variance": 1.0})
  with pytest.raises(RuntimeError, match="CUDA"):
    kernel.squared_exponential(params, np.zeros((3, 2)))


def test_retrieve_params_and_missing_key():  # params_utils.py:90-111
  p = defs.GPParams(model={"lengthscale": torch.tensor(0.0), "constant": 2.0})
  ls, c = params_utils.retrieve_params(p, ["lengthscale", "constant"],
                                       utils.DEFAULT_WARP_FUNC)
  assert abs(float(ls) - (np.log(2.0) + 1e-10)) < 1e-15 and c == 2.0
  ls, = params_utils.retrieve_params(p, ["lengthscale"], None)
  assert float(ls) == 0.0
  with pytest.raises(ValueError, match="Expected parameters"):
    params_utils.retrieve_params(p, ["noise_variance"])


def test_pack_raw_layout_mask_and_unpack():
  model = {"constant": 5.1, "lengthscale": np.array([0.1, 0.2, 0.3]),
           "signal_variance": 0.5, "noise_variance": -4.0}
  raw, mask, scalar = params_utils.pack_raw(model, 3, True,
                                            utils.DEFAULT_WARP_FUNC)
  assert np.allclose(raw, [5.1, 0.5, -4.0, 0.1, 0.2, 0.3]) and not scalar
  assert mask == 0b111110  # constant identity, the rest softplus+eps
  _, mask0, _ = params_utils.pack_raw(model, 3, True, None)
  assert mask0 == 0
  back = params_utils.unpack_like(model, raw, 3, True)
  assert back["constant"] == 5.1 and np.allclose(back["lengthscale"],
                                                 model["lengthscale"])
  # scalar lengthscale: broadcast in, summed gradient out (kernel.py:80)
  m2 = dict(model, lengthscale=0.7)
  raw2, _, scalar2 = params_utils.pack_raw(m2, 3, True, None)
  assert scalar2 and np.allclose(raw2[3:], 0.7)
  g = params_utils.unpack_like(m2, np.array([1., 2., 3., 4., 5., 6.]), 3, True,
                               is_grad=True)
  assert g["lengthscale"] == 15.0
  with pytest.raises(ValueError):
    params_utils.pack_raw(dict(model, lengthscale=np.ones(2)), 3, True, None)


def test_unknown_warp_is_rejected():
  wf = dict(utils.DEFAULT_WARP_FUNC, lengthscale=lambda x: x * x)
  with pytest.raises(NotImplementedError, match="warp"):
    params_utils.pack_raw({"lengthscale": 1.0, "signal_variance": 1.0,
                           "noise_variance": 1.0}, 2, False, wf)


def test_registries_keep_reference_names():  # const.py:22-50
  assert set(const.KERNEL) == {"squared_exponential", "matern32", "matern52",
                               "dot_product", "dot_product_mlp"}
  assert set(const.MEAN) == {"constant", "linear", "linear_mlp", "zero"}
  assert set(const.ACFUN) == {"expected_improvement",
                              "probability_of_improvement", "ucb3",
                              "random_search", "ucb2", "ucb"}
  assert kernel.matern52.__name__ == "matern52"
  assert "mlp" in kernel.squared_exponential_mlp.__name__
  assert const.ACFUN["random_search"].__name__ in ("rand", "random_search")
  p = defs.GPParams(model={})
  for f in (kernel.dot_product, kernel.squared_exponential_mlp):
    with pytest.raises(NotImplementedError):
      f(p, np.zeros((2, 2)))
  with pytest.raises(NotImplementedError):
    mean.linear_mlp(p, np.zeros((2, 2)))
  # objectives are recognised, not traced (objectives.py:213-247)
  assert objectives.objective_terms(objectives.nll) == [(1.0, "nll", {})]
  assert objectives.objective_terms("ekl") == [(1.0, "kl", {})]
  assert objectives.objective_terms(objectives.nll_regkl01) == [
      (1.0, "nll", {}), (0.1, "kl", {})]
  assert objectives.objective_terms(objectives.regeuc) == [(1.0, "euc", {})]
  assert objectives.objective_terms(
      objectives.mul(3.0, objectives.nll_regkl10))[1] == (30.0, "kl", {})
  import functools
  assert utils.distance_spec(functools.partial(
      utils.kl_multivariate_normal, eps=1e-6, partial=False)) == (
          "kl", {"eps": 1e-6, "partial": False})
  with pytest.raises(NotImplementedError):
    objectives.objective_terms(lambda *a, **k: 0.0)


def test_sub_sample_dataset_iterator():  # data_utils.py:72-100
  ds = {
      "a": defs.SubDataset(torch.arange(20.).reshape(10, 2), torch.arange(10.).reshape(10, 1)),
      "b": defs.SubDataset(torch.zeros(3, 2), torch.zeros(3, 1), aligned="tag"),
  }
  it = data_utils.sub_sample_dataset_iterator(0, ds, 4)
  b1, b2 = next(it), next(it)
  assert b1["a"].x.shape == (4, 2) and b1["a"].y.shape == (4, 1)
  # rows stay paired
  assert torch.equal(b1["a"].x[:, 0] / 2, b1["a"].y[:, 0])
  assert not torch.equal(b1["a"].y, b2["a"].y)
  assert b1["b"].x.shape == (3, 2) and b1["b"].aligned == 1  # str tag -> index


def test_gp_dataset_and_cache_bookkeeping():  # gp_test.py:209-277
  x = torch.rand(5, 2, dtype=torch.float64)
  y = torch.rand(5, 1, dtype=torch.float64)
  params = defs.GPParams(model={"constant": 5., "lengthscale": 1.,
                                "signal_variance": 1., "noise_variance": .01})
  model = gp.GP(dataset=[(x, y), (x, y)], mean_func=mean.constant,
                cov_func=kernel.squared_exponential, params=params)
  assert list(model.dataset.keys()) == [0, 1] and model.input_dim == 2
  assert model.params.config["objective"] is objectives.neg_log_marginal_likelihood
  model.params.cache[0] = defs.GPCache(chol=None, kinvy=None, needs_update=False)
  model.update_sub_dataset((x[:2], y[:2]), sub_dataset_key=0, is_append=True)
  assert model.dataset[0].x.shape == (7, 2)
  assert model.params.cache[0].needs_update is True
  model.update_sub_dataset((x[:2], y[:2]), sub_dataset_key=0, is_append=False)
  assert model.dataset[0].x.shape == (2, 2)
  model.update_sub_dataset((x[:1], y[:1]), sub_dataset_key="new", is_append=True)
  assert model.dataset["new"].x.shape == (1, 2)
  # a single point given as 1-D arrays is appended as one row (bayesopt.py:187-190)
  model.update_sub_dataset((x[0], y[0]), sub_dataset_key="new", is_append=True)
  assert model.dataset["new"].x.shape == (2, 2)
  model.update_model_params(dict(params.model))
  assert model.params.cache == {}
  model.set_dataset({"k": (x, y)})
  assert list(model.dataset) == ["k"] and model.params.cache == {}
  model.initialize_params(0)
  assert np.allclose(model.params.model["lengthscale"], np.ones(2))  # gp.py:395-400


def test_infer_parameters_guards():
  params = defs.GPParams(
      model={"constant": 0., "lengthscale": 0., "signal_variance": 0.,
             "noise_variance": 0.},
      config={"method": "adam", "batch_size": 10, "max_training_step": 0,
              "learning_rate": 1e-3})
  x, y = torch.rand(4, 1), torch.rand(4, 1)
  # max_training_step <= 0 -> unchanged (gp.py:111-112); empty dataset too
  assert gp.infer_parameters(mean.constant, kernel.matern32, params,
                             {0: (x, y)}) is params
  assert gp.infer_parameters(mean.constant, kernel.matern32, params, {}) is params
  params.config["max_training_step"] = 1
  params.config["method"] = "nope"
  with pytest.raises(ValueError):
    gp.infer_parameters(mean.constant, kernel.matern32, params, {0: (x, y)})
  params.config["method"] = "adam"
  with pytest.raises(NotImplementedError):  # objectives are recognised, not traced
    gp.infer_parameters(mean.constant, kernel.matern32, params, {0: (x, y)},
                        objective=lambda *a, **k: 0.0)


def test_shard_tasks_round_robin():
  items = list(range(10))
  shards = [gp.shard_tasks(items, r, 4) for r in range(4)]
  assert sorted(sum(shards, [])) == items
  assert shards[1] == [1, 5, 9]


def test_api_surface_of_the_hot_path_modules():
  """Every public name of the reference's hot-path modules exists in the
  mirror, either implemented or raising NotImplementedError
  under the reference's own name.  (Names that only re-export jax -- vmap,
  jit, grad, custom_vjp, partial -- and the out-of-scope file I/O, data
  wrangling and method registries are excluded.)"""
  import importlib
  expected = {
      "gp_utils.gp": ["GP", "HGP", "infer_parameters", "predict", "sample_from_gp",
                      "GPCache", "SubDataset", "GPParams", "retrieve_params"],
      "gp_utils.objectives": ["neg_log_marginal_likelihood", "nll", "kl", "ekl",
                              "euc", "regkl", "regeuc", "add", "mul",
                              "multivariate_normal_divergence",
                              "multivariate_normal_euc_distance", "nll_regkl",
                              "nll_regeuc", "nll_regkl1", "nll_regeuc1",
                              "nll_regkl01", "nll_regeuc01", "nll_regkl10",
                              "nll_regeuc10", "retrieve_params"],
      "gp_utils.kernel": ["covariance_matrix", "squared_exponential", "matern32",
                          "matern52", "dot_product", "with_mlp_bases",
                          "with_kumar_bases", "squared_exponential_mlp",
                          "matern32_mlp", "matern52_mlp", "dot_product_mlp",
                          "squared_exponential_kumar", "matern32_kumar",
                          "matern52_kumar", "dot_product_kumar"],
      "gp_utils.mean": ["mean_vector", "zero", "constant", "linear", "linear_mlp"],
      "gp_utils.utils": ["EPS", "identity_warp", "softplus_warp", "squareplus_warp",
                         "DEFAULT_WARP_FUNC", "SubDataset",
                         "sub_sample_dataset_iterator", "partial_kl_mvn",
                         "kl_multivariate_normal",
                         "euclidean_multivariate_normal"],
      "basics.linalg": ["compute_delta_y_and_cov", "solve_gp_linear_system",
                        "solve_linear_system", "cholesky_cache",
                        "inverse_spdmatrix_vector_product", "svd_matrix_sqrt",
                        "safe_l2norm"],
      "basics.params_utils": ["retrieve_params"],
      "basics.data_utils": ["sub_sample_dataset_iterator", "log_dataset"],
      "basics.definitions": ["GPCache", "SubDataset", "GPParams"],
      "basics.lbfgs": ["lbfgs"],
      "basics.bfgs": ["bfgs"],
      "bo_utils.acfun": ["acfun_wrapper", "expected_improvement",
                         "probability_of_improvement", "ucb", "ucb2", "ucb3",
                         "ucb4", "ei", "pi", "rand", "random_search",
                         "expected_improvement_sub", "ucb_sub",
                         "probability_of_improvement_sub"],
      "bo_utils.bayesopt": ["retrain_model", "bayesopt", "simulated_bayesopt",
                            "run_bayesopt"],
      "bo_utils.const": ["KERNEL", "MEAN", "ACFUN", "ACFUN_SUB"],
  }
  for mod, names in expected.items():
    m = importlib.import_module("hyperbo_b200." + mod)
    missing = [n for n in names if not hasattr(m, n)]
    assert not missing, (mod, missing)


def test_explicit_matrix_helpers():  # linalg.py:29-33,112-197 (plain torch)
  from hyperbo_b200.basics import linalg
  rng = np.random.default_rng(0)
  a = rng.standard_normal((6, 6))
  spd = torch.from_numpy(a @ a.T + 6 * np.eye(6))
  b = torch.from_numpy(rng.standard_normal(6))
  chol, x = linalg.solve_linear_system(spd, b)
  assert torch.allclose(spd @ x, b, atol=1e-12)
  assert torch.allclose(chol @ chol.T, spd, atol=1e-12)
  x2 = linalg.inverse_spdmatrix_vector_product(spd, b[:, None], chol)
  assert torch.allclose(x2[:, 0], x)
  low = torch.from_numpy(a[:, :3] @ a[:, :3].T)  # rank 3
  f = linalg.svd_matrix_sqrt(low)
  assert f.shape == (6, 3) and torch.allclose(f @ f.T, low, atol=1e-10)
  assert abs(float(linalg.safe_l2norm(torch.tensor([3.0, 4.0]))) - 5.0) < 1e-15
  for f_ in (kernel.covariance_matrix, kernel.with_mlp_bases, mean.mean_vector):
    with pytest.raises(NotImplementedError):
      f_(lambda *a: 0.0)
  with pytest.raises(NotImplementedError):  # not differentiable by the engine
    utils.warp_kind({"lengthscale": utils.squareplus_warp}, "lengthscale")


def test_params_checkpoint_round_trip(tmp_path):
  """params_utils.save_params / load_params / log_params_loss
  (params_utils.py:64-87,193-207): callables are stored as strings, the cache is
  not part of a checkpoint, the state tuple (step, loss) comes back."""
  import numpy as np
  import torch
  from hyperbo_b200.basics import definitions as defs
  from hyperbo_b200.basics import params_utils
  from hyperbo_b200.gp_utils import objectives, utils
  params = defs.GPParams(
      model={"constant": 1.5, "lengthscale": torch.tensor([0.5, 0.7]),
             "signal_variance": np.float64(1.0), "noise_variance": -3.0},
      config={"method": "adam", "objective": objectives.nll, "learning_rate": 1e-3},
      cache={0: defs.GPCache(chol=torch.eye(2), kinvy=torch.ones(2, 1),
                             needs_update=False)})
  path = str(tmp_path / "ckpt" / "params.pkl")
  params_utils.log_params_loss(step=7, params=params, loss=1.25,
                               warp_func=utils.DEFAULT_WARP_FUNC,
                               params_save_file=path)
  loaded, state = params_utils.load_params(path, include_state=True)
  assert state == (7, 1.25)
  assert isinstance(loaded, defs.GPParams) and loaded.cache == {}
  assert loaded.model["constant"] == 1.5
  assert np.allclose(loaded.model["lengthscale"], [0.5, 0.7])
  assert isinstance(loaded.config["objective"], str)          # callable -> str
  as_dict = params_utils.load_params(path, use_gpparams=False)
  assert set(as_dict) == {"model", "config", "cache", "samples"}
  with __import__("pytest").raises(FileNotFoundError):
    params_utils.load_params(str(tmp_path / "missing.pkl"))
  params_utils.save_to_file(str(tmp_path / "nothing.pkl"), None)  # no state: no file
  assert not (tmp_path / "nothing.pkl").exists()


def test_log_dataset_reports_every_sub_dataset(caplog):
  """data_utils.log_dataset (data_utils.py:29-69): len, shape, mean / median / min /
  max per field; empty arrays as nan; non-array fields untouched."""
  import logging
  import numpy as np
  import torch
  from hyperbo_b200.basics import data_utils
  from hyperbo_b200.basics import definitions as defs
  ds = {0: defs.SubDataset(np.arange(6.0).reshape(3, 2), torch.ones(3, 1)),
        "empty": defs.SubDataset(np.zeros((0, 2)), np.zeros((0, 1)), aligned="a")}
  with caplog.at_level(logging.INFO):
    data_utils.log_dataset(ds)
  text = caplog.text
  assert "dataset len = 2." in text
  for word in ("shape", "mean", "median", "min", "max"):
    assert f"dataset {word}:" in text
  assert "(3, 2)" in text and "nan" in text and "'a'" in text


def test_round_robin_sharding_is_a_balanced_partition():
  """objectives._shard with a rotating start (several short launches must not all
  land on rank 0): for every (items, world, start) the ranks' shares are disjoint,
  cover everything and differ by at most one."""
  from hypothesis import given, settings, strategies as st
  from hyperbo_b200.gp_utils import objectives

  @settings(max_examples=200, deadline=None)
  @given(st.integers(0, 40), st.integers(1, 9), st.integers(0, 100))
  def check(n, world, start):
    items = list(range(n))
    shares = [objectives._shard(items, r, world, start) for r in range(world)]
    flat = sorted(x for s in shares for x in s)
    assert flat == items
    sizes = [len(s) for s in shares]
    assert max(sizes) - min(sizes) <= 1
    # continuing the rotation after this launch keeps the TOTAL balanced too
    nxt = [objectives._shard(items, r, world, start + n) for r in range(world)]
    tot = [len(a) + len(b) for a, b in zip(shares, nxt)]
    assert max(tot) - min(tot) <= 1

  check()


def test_lbfgs_evaluator_memo_is_bounded_and_exact():
  """basics/lbfgs._Evaluator: hits return the stored pair, the memo never exceeds
  `keep` entries, prefetch is a no-op without a multi-point objective."""
  import numpy as np
  from hyperbo_b200.basics import lbfgs
  n = {"c": 0}

  def fn(x):
    n["c"] += 1
    return float(x @ x), 2 * x

  ev = lbfgs._Evaluator(fn, keep=3)
  pts = [np.array([float(i), 1.0]) for i in range(5)]
  for p in pts:
    ev(p)
  assert n["c"] == 5 and len(ev.memo) == 3
  v, g = ev(pts[-1])                      # hit
  assert n["c"] == 5 and v == float(pts[-1] @ pts[-1])
  ev(pts[0])                              # evicted earlier: evaluated again
  assert n["c"] == 6
  ev.prefetch([np.array([9.0, 9.0]), np.array([8.0, 8.0])])
  assert n["c"] == 6 and ev.calls == 6 and ev.points == 6
