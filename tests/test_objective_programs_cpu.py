"""Host logic of the objective programs (hyperbo_b200/gp_utils/objectives.py)
on CPU: the decomposition of nll / kl / add / mul objectives into weighted
engine calls, round-robin task sharding and the one-all-reduce contract under
gloo.  The engine's arithmetic is supplied by tests/fake_engine.py (the
oracle); on a B200 the same host code drives the CUDA kernels and is checked by
tests/test_gpu_objectives.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hyperbo_b200.basics import definitions as defs
from hyperbo_b200.basics import params_utils
from hyperbo_b200.gp_utils import kernel, mean, objectives, utils
from oracle import hyperbo_oracle as O
from tests import fake_engine
from tests import helpers as H

WF, WFO = utils.DEFAULT_WARP_FUNC, O.DEFAULT_WARP_FUNC
COVS = {"squared_exponential": kernel.squared_exponential,
        "matern32": kernel.matern32, "matern52": kernel.matern52}
MEANS = {"constant": mean.constant, "zero": mean.zero}


def _case(name):
  g = H.load_golden_kl(name)
  model = H.model_from_raw(g["raw"], g["d"], g["mean"])
  dataset = {k: defs.SubDataset(*v) for k, v in g["dataset"].items()}
  return g, model, defs.GPParams(model=dict(model)), dataset


def test_fake_engine_is_the_oracle():
  d = 3
  ds_np = {t: O.make_task(t, n, d, "matern52") for t, n in enumerate([20, 31])}
  model = O.init_raw_params(d)
  eng = fake_engine.FakeEngine()
  ds = eng.pack([(t, x, y) for t, (x, y) in ds_np.items()])
  sums = eng.nll_grad(2, 1, ds, H.raw_vec(model, d), H.default_mask(d)).numpy()
  v, g = O.nll_value_and_grad("constant", "matern52", model, ds_np, WFO)
  assert abs(sums[0] / sums[-1] - v) < 1e-12 * abs(v)
  assert H.rel(sums[1:-1] / sums[-1], H.grad_vec(g, d)) < 1e-11


@pytest.mark.parametrize("name", H.golden_cases(kl=True))
def test_kl_program_matches_golden(monkeypatch, name):
  fake_engine.install(monkeypatch)
  g, model, params, dataset = _case(name)
  mf, cf = MEANS[g["mean"]], COVS[g["cov"]]
  val, grads = objectives.value_and_grad(objectives.kl, mf, cf, params, dataset,
                                         WF)
  assert abs(float(val) - g["kl"]) < 1e-11 * abs(g["kl"])
  assert H.rel(H.grad_vec(grads, g["d"]), g["kl_grad"]) < 1e-10
  # the callable, a weight, and eps > 0 (value-only trace term)
  import functools
  kl = utils.kl_multivariate_normal
  v = objectives.ekl(mf, cf, params, dataset, WF)
  assert abs(float(v) - g["kl"]) < 1e-11 * abs(g["kl"])
  v = objectives.kl(mf, cf, params, dataset, WF,
                    distance=functools.partial(kl, weight=3.0))
  assert abs(float(v) - 3.0 * g["kl"]) < 1e-11 * abs(g["kl"])
  v = objectives.kl(mf, cf, params, dataset, WF,
                    distance=functools.partial(kl, eps=1e-6))
  assert abs(float(v) - g["kl_eps"]) < 1e-10 * abs(g["kl_eps"])


def test_combined_objective_program(monkeypatch):
  eng = fake_engine.install(monkeypatch)
  g, model, params, dataset = _case("kl_m52_const_d3")
  objective = objectives.add(objectives.nll, objectives.mul(0.3, objectives.regkl))
  prog = objectives.compile_objective(objective, mean.constant, kernel.matern52,
                                      dataset)
  # nll launch + ONE multi-right-hand-side kl launch (one factorisation per
  # aligned sub-dataset)
  # (sub-datasets with the same number of columns share a launch)
  n_m = len({int(y.shape[1]) for _, _, y in objectives._aligned_subs(dataset)})
  assert len(prog.launches) == 1 + n_m and prog.has_exact_grad
  assert all(isinstance(l, objectives._LaunchMRHS) for l in prog.launches[1:])
  raw, mask, _ = params_utils.pack_raw(params.model, g["d"], True, WF)
  sums = prog.sums(raw, mask).numpy()
  assert eng.calls == 1 + n_m and sums[-1] == 1.0
  v_nll, g_nll = O.nll_value_and_grad("constant", "matern52", model,
                                      g["dataset"], WFO)
  v_kl, g_kl = O.kl_value_and_grad("constant", "matern52", model, g["dataset"],
                                   WFO)
  assert abs(sums[0] - (v_nll + 0.3 * v_kl)) < 1e-11 * abs(sums[0])
  want = H.grad_vec(g_nll, g["d"]) + 0.3 * H.grad_vec(g_kl, g["d"])
  assert H.rel(sums[1:-1], want) < 1e-10
  # the round-1 decomposition (m + 2 weighted tasks sharing x) gives the same sums
  monkeypatch.setattr(objectives, "KL_MULTI_RHS", False)
  prog_w = objectives.compile_objective(objective, mean.constant, kernel.matern52,
                                        dataset)
  # nll launch + zero-mean kl launch + model-mean kl launch
  assert len(prog_w.launches) == 3
  assert H.rel(prog_w.sums(raw, mask).numpy(), sums) < 1e-11
  # zero mean: the model-mean tasks ride in the zero-mean launch
  g2, model2, params2, dataset2 = _case("kl_se_zero_d2")
  prog2 = objectives.compile_objective(objectives.kl, mean.zero,
                                       kernel.squared_exponential, dataset2)
  assert len(prog2.launches) == 1
  monkeypatch.setattr(objectives, "KL_MULTI_RHS", True)
  prog2 = objectives.compile_objective(objectives.kl, mean.zero,
                                       kernel.squared_exponential, dataset2)
  assert len(prog2.launches) == 1


def _fd_grad(fn, model, d, h=1e-6):
  """central differences of fn(model) over the raw parameters (order of raw_vec)."""
  out = []
  for key, idx in ([("constant", None), ("signal_variance", None),
                    ("noise_variance", None)] + [("lengthscale", k) for k in range(d)]):
    vals = []
    if key not in model:
      out.append(0.0)
      continue
    for sgn in (1.0, -1.0):
      m = {k: np.array(v, dtype=np.float64, copy=True) for k, v in model.items()}
      if idx is None:
        m[key] = m[key] + sgn * h
      else:
        m[key][idx] += sgn * h
      vals.append(fn(m))
    out.append((vals[0] - vals[1]) / (2 * h))
  return np.array(out)


@pytest.mark.parametrize("name", H.golden_cases(kl=True))
def test_euclidean_regulariser_program(monkeypatch, name):
  """objectives.euc / nll_regeuc(c) as engine programs (hb_euclid_grad): the value
  is the committed fixture, the gradient matches central differences of the
  oracle's restatement of utils.euclidean_multivariate_normal."""
  fake_engine.install(monkeypatch)
  g, model, params, dataset = _case(name)
  mf, cf = MEANS[g["mean"]], COVS[g["cov"]]
  d = g["d"]
  val, grads = objectives.value_and_grad(objectives.euc, mf, cf, params, dataset, WF)
  assert abs(float(val) - g["euc"]) < 1e-11 * abs(g["euc"])
  model_full = dict(model)
  model_full["lengthscale"] = np.broadcast_to(
      np.asarray(model["lengthscale"], dtype=np.float64), (d,)).copy()
  fd = _fd_grad(lambda m: O.multivariate_normal_divergence(
      g["mean"], g["cov"], m, g["dataset"], WFO,
      distance=O.euclidean_multivariate_normal), model_full, d)
  got = H.grad_vec(grads, d)
  if g["mean"] == "zero":
    fd[0] = 0.0
  assert H.rel(got, fd) < 1e-6
  # weights of the two norms + the nll_regeuc(c) combination
  import functools
  dist = functools.partial(utils.euclidean_multivariate_normal, mean_weight=0.5,
                           cov_weight=2.0)
  objective = objectives.add(objectives.nll, objectives.mul(
      0.3, functools.partial(objectives.multivariate_normal_divergence,
                             distance=dist)))
  v2, _ = objectives.value_and_grad(objective, mf, cf, params, dataset, WF)
  want = O.neg_log_marginal_likelihood(g["mean"], g["cov"], model, g["dataset"], WFO) \
      + 0.3 * O.multivariate_normal_divergence(
          g["mean"], g["cov"], model, g["dataset"], WFO,
          distance=functools.partial(O.euclidean_multivariate_normal,
                                     mean_weight=0.5, cov_weight=2.0))
  assert abs(float(v2) - want) < 1e-10 * abs(want)


def test_multi_column_nll_value(monkeypatch):
  """exclude_aligned=False on y with m > 1 columns (objectives.py:153-155)."""
  eng = fake_engine.install(monkeypatch)

  def factorize(kid, mid, ds, raw, mask, want_chol=True, want_alpha=True):
    _, per_task = eng.nll_grad(kid, mid, ds, raw, mask, want_task_nll=True)
    return None, None, per_task, None

  monkeypatch.setattr(eng, "factorize", factorize, raising=False)
  g, model, params, dataset = _case("kl_m52_const_d3")
  total, key2nll = objectives.neg_log_marginal_likelihood(
      mean.constant, kernel.matern52, params, dataset, WF, exclude_aligned=False,
      return_key2nll=True)
  want, want_k2n = O.neg_log_marginal_likelihood(
      "constant", "matern52", model, g["dataset"], WFO, exclude_aligned=False,
      return_key2nll=True)
  assert abs(float(total) - want) < 1e-11 * abs(want)
  assert set(key2nll) == set(want_k2n)
  for k in want_k2n:
    assert abs(float(key2nll[k]) - want_k2n[k]) < 1e-10 * abs(want_k2n[k])


# ---- two ranks under gloo: sharded programs + ONE all-reduce -----------------
def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _worker(rank, world, port, out):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  from hyperbo_b200 import engine as _engine
  eng = fake_engine.FakeEngine()
  _engine.Engine.get = staticmethod(lambda *a, **k: eng)
  g, model, params, dataset = _case("kl_m52_const_d3")
  objective = objectives.nll_regkl(0.7)
  prog = objectives.compile_objective(objective, mean.constant, kernel.matern52,
                                      dataset, rank, world)
  raw, mask, _ = params_utils.pack_raw(params.model, g["d"], True, WF)
  sums = prog.sums(raw, mask)
  out[rank] = (sums.numpy().copy(),
               sum(l.ds.num_tasks for l in prog.launches))
  dist.destroy_process_group()


def test_two_rank_objective_program_matches_single_process(monkeypatch):
  world = 2
  mgr = mp.Manager()
  out = mgr.dict()
  mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
  (s0, n0), (s1, n1) = out[0], out[1]
  assert np.array_equal(s0, s1)          # every rank holds the reduced vector
  fake_engine.install(monkeypatch)
  g, model, params, dataset = _case("kl_m52_const_d3")
  prog = objectives.compile_objective(objectives.nll_regkl(0.7), mean.constant,
                                      kernel.matern52, dataset)
  raw, mask, _ = params_utils.pack_raw(params.model, g["d"], True, WF)
  ref = prog.sums(raw, mask).numpy()
  assert n0 + n1 == sum(l.ds.num_tasks for l in prog.launches)  # a partition
  assert abs(n0 - n1) <= 2                                      # balanced
  assert H.rel(s0[:-1], ref[:-1]) < 1e-12 and s0[-1] == 1.0


# ---- gp.infer_parameters on objective programs (host loop + Adam contract) ---
def _oracle_adam(model, ds_np, cov, c, lr, steps):
  opt, losses, m = O.Adam(lr), [], dict(model)
  for _ in range(steps):
    v1, g1 = O.nll_value_and_grad("constant", cov, m, ds_np, WFO)
    v2, g2 = O.kl_value_and_grad("constant", cov, m, ds_np, WFO)
    losses.append(v1 + c * v2)
    m = opt.update(m, {k: np.asarray(g1[k]) + c * np.asarray(g2[k]) for k in g1})
  return m, losses


def _train(rank=0, world=1):
  from hyperbo_b200.gp_utils import gp
  g, model, params, dataset = _case("kl_m52_const_d3")
  params.config = {"method": "adam", "learning_rate": 1e-2,
                   "max_training_step": 4, "batch_size": 10**6}
  losses = []
  out = gp.infer_parameters(mean.constant, kernel.matern52, params, dataset, WF,
                            objective=objectives.nll_regkl(0.5),
                            callback=lambda i, p, l: losses.append(l))
  return g, model, out, losses


def test_infer_parameters_on_a_program_matches_oracle_adam(monkeypatch):
  fake_engine.install(monkeypatch)
  g, model, out, losses = _train()
  ref_model, ref_losses = _oracle_adam(model, g["dataset"], "matern52", 0.5,
                                       1e-2, 4)
  assert H.rel(losses, ref_losses) < 1e-11
  for k in ref_model:
    assert H.rel(out.model[k], ref_model[k]) < 1e-10, k
  assert out.cache == {}


def _train_worker(rank, world, port, out):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  from hyperbo_b200 import engine as _engine
  eng = fake_engine.FakeEngine()
  _engine.Engine.get = staticmethod(lambda *a, **k: eng)
  g, model, res, losses = _train(rank, world)
  out[rank] = (losses, H.raw_vec({k: np.asarray(v) for k, v in res.model.items()},
                                 g["d"]))
  dist.destroy_process_group()


def test_two_rank_infer_parameters_on_a_program(monkeypatch):
  """Every rank passes the SAME dataset (gp.infer_parameters' contract); tasks
  of every launch are sharded, one all-reduce per step, identical replicas."""
  world = 2
  mgr = mp.Manager()
  out = mgr.dict()
  mp.spawn(_train_worker, args=(world, _free_port(), out), nprocs=world,
           join=True)
  (l0, p0), (l1, p1) = out[0], out[1]
  assert l0 == l1 and np.array_equal(p0, p1)
  g, model, params, dataset = _case("kl_m52_const_d3")
  ref_model, ref_losses = _oracle_adam(model, g["dataset"], "matern52", 0.5,
                                       1e-2, 4)
  assert H.rel(l0, ref_losses) < 1e-11
  assert H.rel(p0, H.raw_vec(ref_model, g["d"])) < 1e-10


def test_lbfgs_on_ekl_matches_the_driver_on_the_oracle(monkeypatch):
  from hyperbo_b200.basics import lbfgs as _lbfgs
  from hyperbo_b200.gp_utils import gp
  fake_engine.install(monkeypatch)
  g, model, params, dataset = _case("kl_m52_const_d3")
  params.config = {"method": "lbfgs", "max_training_step": 3, "batch_size": 10**6,
                   "alpha": 1.0, "objective": objectives.ekl}
  out = gp.infer_parameters(mean.constant, kernel.matern52, params, dataset, WF,
                            objective=objectives.ekl)
  d = g["d"]

  def val_and_grad(v):
    m = H.model_from_raw(v, d, "constant")
    val, gr = O.kl_value_and_grad("constant", "matern52", m, g["dataset"], WFO)
    return val, H.grad_vec(gr, d)

  _, v, _ = _lbfgs.lbfgs(val_and_grad, H.raw_vec(model, d), steps=3, alpha=1.0)
  assert H.rel(H.raw_vec({k: np.asarray(x) for k, x in out.model.items()}, d),
               v) < 1e-8


def test_subsampled_training_rebuilds_the_program_every_step(monkeypatch):
  """batch_size <= n: data_utils.sub_sample_dataset_iterator draws new rows of
  every sub-dataset per step (data_utils.py:72-100), so the objective program
  is recompiled per step; the losses equal the oracle's on the same batches.
  Also: the objective given by NAME (GP.initialize_params resolves strings)."""
  from hyperbo_b200.basics import data_utils
  from hyperbo_b200.gp_utils import gp
  fake_engine.install(monkeypatch)
  g, model, params, dataset = _case("kl_m52_const_d3")
  params.config = {"method": "adam", "learning_rate": 1e-2,
                   "max_training_step": 3, "batch_size": 25}
  losses = []
  gp.infer_parameters(mean.constant, kernel.matern52, params, dataset, WF,
                      objective="ekl", key=5,
                      callback=lambda i, p, l: losses.append(l))
  it = data_utils.sub_sample_dataset_iterator(
      5, {k: defs.SubDataset(*v) for k, v in dataset.items()}, 25)
  opt, m, want = O.Adam(1e-2), dict(model), []
  for _ in range(3):
    batch = next(it)
    b_np = {k: (np.asarray(s.x), np.asarray(s.y), s.aligned)
            for k, s in batch.items()}
    assert all(s[0].shape[0] <= 25 for s in b_np.values())
    v, gr = O.kl_value_and_grad("constant", "matern52", m, b_np, WFO)
    want.append(v)
    m = opt.update(m, gr)
  assert H.rel(losses, want) < 1e-10


# ---- candidate-axis sharding of an acquisition sweep ----------
def _acq_case():
  from hyperbo_b200.gp_utils import gp
  d = 3
  ds_np = O.make_dataset(3, 40, d, "matern32")
  model = gp.GP({k: defs.SubDataset(*v) for k, v in ds_np.items()},
                mean.constant, kernel.matern32,
                defs.GPParams(model=dict(O.init_raw_params(d))), WF)
  xq, _ = O.make_task(77, 101, d, "matern32")  # 101: not a multiple of 2
  return ds_np, model, xq


def _acq_worker(rank, world, port, out):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  from hyperbo_b200 import engine as _engine
  from hyperbo_b200.bo_utils import acfun
  eng = fake_engine.FakeEngine()
  _engine.Engine.get = staticmethod(lambda *a, **k: eng)
  _, model, xq = _acq_case()
  ei = acfun.shard_candidates(acfun.ei)(model=model, sub_dataset_key=1,
                                        x_queries=xq)
  out[rank] = ei.numpy().copy()
  dist.destroy_process_group()


def test_two_rank_candidate_sharding(monkeypatch):
  from hyperbo_b200.bo_utils import acfun
  world = 2
  mgr = mp.Manager()
  out = mgr.dict()
  mp.spawn(_acq_worker, args=(world, _free_port(), out), nprocs=world, join=True)
  assert np.array_equal(out[0], out[1]) and out[0].shape == (101, 1)
  fake_engine.install(monkeypatch)
  ds_np, model, xq = _acq_case()
  full = acfun.ei(model=model, sub_dataset_key=1, x_queries=xq).numpy()
  assert H.rel(out[0], full) < 1e-12  # (BLAS blocking differs with the slice)
  want = O.acquisition("ei", "constant", "matern32", O.init_raw_params(3), ds_np,
                       1, xq, WFO)
  assert H.rel(full, want) < 1e-9
  # no process group: the wrapper is the acquisition itself
  same = acfun.shard_candidates(acfun.ei)(model=model, sub_dataset_key=1,
                                          x_queries=xq).numpy()
  assert np.array_equal(same, full)


@pytest.mark.parametrize("case", H.load_kat_div(), ids=lambda c: "katdiv%d" % c["id"])
def test_divergence_programs_match_mpmath_known_answers(monkeypatch, case):
  """kl (hb_nll_grad_mrhs contract) and euc (hb_euclid_grad contract) programs, value
  and gradient, against 60-digit answers that were computed without the oracle."""
  fake_engine.install(monkeypatch)
  c = case
  wf = WF if c["warped"] else None
  model = H.model_from_raw(c["raw"], c["d"], c["mean"])
  params = defs.GPParams(model=dict(model))
  dataset = {k: defs.SubDataset(*v) for k, v in c["dataset"].items()}
  mf, cf = MEANS[c["mean"]], COVS[c["cov"]]
  for name, objective, tol in (("kl", objectives.kl, 1e-9), ("euc", objectives.euc, 1e-10)):
    val, grads = objectives.value_and_grad(objective, mf, cf, params, dataset, wf)
    assert abs(float(val) - c[name]) < 1e-10 * abs(c[name]), name
    assert H.rel(H.grad_vec(grads, c["d"]), c[name + "_grad"]) < tol, name
