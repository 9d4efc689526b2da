"""hb_nll_grad_mrhs: R right-hand-side columns on ONE factorisation per task (the
building block of the empirical-KL objective, gp_utils/objectives.py:29-101 of the
reference).  Checked against the dense closed form of tests/fake_engine.py, against
the hb_nll_grad_weighted decomposition (m + 2 tasks sharing x) and on both
factorisation paths (persistent kernel / launch per column)."""
import numpy as np
import pytest
import torch

from hyperbo_b200.basics import definitions as defs
from hyperbo_b200.engine import Engine
from hyperbo_b200.gp_utils import kernel, mean, objectives, utils
from oracle import hyperbo_oracle as O
from tests import fake_engine
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _problem(ns, d, R, seed, cov):
  rng = np.random.default_rng(seed)
  tasks = []
  for t, n in enumerate(ns):
    x, y = O.make_task(100 * seed + t, n, d, cov) if n else (np.zeros((0, d)), np.zeros(0))
    tasks.append((t, x, np.zeros(n)))
  B = np.concatenate([rng.standard_normal((R, n)).reshape(-1) for n in ns])
  # (Engine.pack drops empty tasks: weights exist for the non-empty ones only)
  nt = sum(1 for n in ns if n)
  cw = rng.uniform(-1.0, 2.0, size=(nt, R))
  w = rng.uniform(0.5, 2.0, size=nt)
  return tasks, B, cw, w


@pytest.mark.parametrize("kid,cov", [(0, "squared_exponential"), (1, "matern32"),
                                     (2, "matern52")])
@pytest.mark.parametrize("d,R,ns", [(1, 1, [1, 37]), (3, 5, [70, 20, 129, 64]),
                                    (9, 11, [130, 300]), (4, 9, [200, 0, 65])])
def test_mrhs_matches_dense_closed_form(kid, cov, d, R, ns):
  eng = Engine.get()
  fake = fake_engine.FakeEngine()
  tasks, B, cw, w = _problem(ns, d, R, 3 + d, cov)
  model = {"constant": 0.3, "signal_variance": 0.2, "noise_variance": -3.0,
           "lengthscale": np.linspace(-0.3, 0.4, d)}
  raw, mask = H.raw_vec(model, d), H.default_mask(d)
  col_mean = np.zeros(R, dtype=np.int32)
  col_mean[R - 1] = 1
  ds = eng.pack(tasks)
  ds_cpu = fake.pack(tasks)
  for mean_id, cm, jit in ((1, col_mean, 1e-6), (0, col_mean, 1e-4), (1, None, 1e-6)):
    got = eng.nll_grad_mrhs(kid, mean_id, ds, R, B, cw, cm, raw, mask, weights=w,
                            jitter=jit).cpu().numpy()
    want = fake.nll_grad_mrhs(kid, mean_id, ds_cpu, R, B, cw, cm, raw, mask,
                              weights=w, jitter=jit).numpy()
    assert abs(got[0] - want[0]) < 1e-10 * abs(want[0]), (mean_id, jit)
    assert H.rel(got[1:-1], want[1:-1]) < 1e-8, (mean_id, jit)
    assert got[-1] == want[-1] == sum(1 for n in ns if n)


def test_mrhs_same_on_both_factorisation_paths_and_repeatable():
  eng = Engine.get()
  d, R, ns = 4, 6, [256, 256, 200, 64]
  tasks, B, cw, w = _problem(ns, d, R, 11, "matern52")
  raw, mask = H.raw_vec(O.init_raw_params(d), d), H.default_mask(d)
  ds = eng.pack(tasks)
  res = {}
  try:
    for path in (0, 2):
      eng.h.set_option("fused", path)
      a = eng.nll_grad_mrhs(2, 1, ds, R, B, cw, None, raw, mask, weights=w).cpu().numpy()
      b = eng.nll_grad_mrhs(2, 1, ds, R, B, cw, None, raw, mask, weights=w).cpu().numpy()
      assert np.array_equal(a, b)          # fixed-order reductions
      res[path] = a
  finally:
    eng.h.set_option("fused", 1)
  assert H.rel(res[0], res[2]) < 1e-11


def test_mrhs_fp32_engine():
  eng = Engine.get(dtype=torch.float32)
  fake = fake_engine.FakeEngine()
  d, R, ns = 3, 5, [100, 150]
  tasks, B, cw, w = _problem(ns, d, R, 5, "matern52")
  cw = np.abs(cw) + 0.1     # (no cancellation between the columns: fp32 tolerances)
  raw, mask = H.raw_vec(O.init_raw_params(d), d), H.default_mask(d)
  col_mean = np.array([0, 0, 0, 0, 1], dtype=np.int32)
  got = eng.nll_grad_mrhs(2, 1, eng.pack(tasks), R, B, cw, col_mean, raw, mask,
                          weights=w).cpu().numpy().astype(np.float64)
  want = fake.nll_grad_mrhs(2, 1, fake.pack(tasks), R, B, cw, col_mean, raw, mask,
                            weights=w).numpy()
  assert abs(got[0] - want[0]) < 1e-4 * abs(want[0])
  assert H.rel(got[1:-1], want[1:-1]) < 5e-3


@pytest.mark.parametrize("name", H.golden_cases(kl=True))
def test_kl_program_multi_rhs_equals_weighted_decomposition(monkeypatch, name):
  """The KL objective through hb_nll_grad_mrhs (one factorisation per aligned
  sub-dataset) and through m + 2 weighted tasks: same value and gradient, and both
  equal the committed fixture."""
  g = H.load_golden_kl(name)
  model = H.model_from_raw(g["raw"], g["d"], g["mean"])
  params = defs.GPParams(model=dict(model))
  dataset = {k: defs.SubDataset(*v) for k, v in g["dataset"].items()}
  mf = {"constant": mean.constant, "zero": mean.zero}[g["mean"]]
  cf = {"squared_exponential": kernel.squared_exponential, "matern32": kernel.matern32,
        "matern52": kernel.matern52}[g["cov"]]
  wf = utils.DEFAULT_WARP_FUNC
  out = {}
  for flag in (True, False):
    monkeypatch.setattr(objectives, "KL_MULTI_RHS", flag)
    prog = objectives.compile_objective(objectives.kl, mf, cf, dataset)
    kinds = {type(l).__name__ for l in prog.launches}
    assert kinds == ({"_LaunchMRHS"} if flag else {"_Launch"})
    eng = Engine.get()
    l0 = eng.launch_count()
    val, grads = objectives.value_and_grad(objectives.kl, mf, cf, params, dataset, wf)
    out[flag] = (float(val), H.grad_vec(grads, g["d"]), eng.launch_count() - l0)
    assert abs(out[flag][0] - g["kl"]) < 1e-10 * abs(g["kl"])
    assert H.rel(out[flag][1], g["kl_grad"]) < 1e-8
  assert abs(out[True][0] - out[False][0]) < 1e-11 * abs(out[False][0])
  assert H.rel(out[True][1], out[False][1]) < 1e-9


# ---- hb_euclid_grad: the Euclidean regulariser (utils.py:151-173), value + grad --
@pytest.mark.parametrize("kid,cov", [(0, "squared_exponential"), (1, "matern32"),
                                     (2, "matern52")])
@pytest.mark.parametrize("d,R,ns", [(1, 1, [1, 37]), (3, 5, [70, 20, 129, 64]),
                                    (9, 11, [130, 300])])
def test_euclid_grad_matches_dense_closed_form(kid, cov, d, R, ns):
  eng = Engine.get()
  fake = fake_engine.FakeEngine()
  rng = np.random.default_rng(17 + d)
  tasks, mu0s = [], []
  for t, n in enumerate(ns):
    x, _ = O.make_task(300 + t, n, d, cov)
    mu0 = rng.standard_normal(n)
    tasks.append((t, x, mu0))
  Yc = np.concatenate([0.4 * rng.standard_normal((R, n)).reshape(-1) for n in ns])
  w = rng.uniform(0.5, 2.0, size=len(ns))
  model = {"constant": 0.3, "signal_variance": 0.2, "noise_variance": -3.0,
           "lengthscale": np.linspace(-0.3, 0.4, d)}
  raw, mask = H.raw_vec(model, d), H.default_mask(d)
  ds, ds_cpu = eng.pack(tasks), fake.pack(tasks)
  for mean_id, mw, cw in ((1, 1.0, 1.0), (0, 0.5, 2.0)):
    got = eng.euclid_grad(kid, mean_id, ds, R, Yc, ds.y, raw, mask, mw, cw,
                          weights=w).cpu().numpy()
    want = fake.euclid_grad(kid, mean_id, ds_cpu, R, Yc, ds_cpu.y, raw, mask, mw, cw,
                            weights=w).numpy()
    assert abs(got[0] - want[0]) < 1e-11 * abs(want[0]), mean_id
    assert H.rel(got[1:-1], want[1:-1]) < 1e-9, mean_id
    assert got[-1] == want[-1] == len(ns)
  eng32 = Engine.get(dtype=torch.float32)
  got = eng32.euclid_grad(kid, 1, eng32.pack(tasks), R, Yc, ds.y, raw, mask,
                          weights=w).cpu().numpy().astype(np.float64)
  want = fake.euclid_grad(kid, 1, ds_cpu, R, Yc, ds_cpu.y, raw, mask, weights=w).numpy()
  assert abs(got[0] - want[0]) < 1e-5 * abs(want[0])
  assert H.rel(got[1:-1], want[1:-1]) < 1e-3


@pytest.mark.parametrize("name", H.golden_cases(kl=True))
def test_euclidean_regulariser_value_and_grad(name):
  """objectives.euc through its engine program: value = the committed fixture (and
  the value-only path of multivariate_normal_divergence), gradient = the dense
  closed form; nll_regeuc(c) = nll + c * regeuc (objectives.py:238)."""
  g = H.load_golden_kl(name)
  model = H.model_from_raw(g["raw"], g["d"], g["mean"])
  params = defs.GPParams(model=dict(model))
  dataset = {k: defs.SubDataset(*v) for k, v in g["dataset"].items()}
  mf = {"constant": mean.constant, "zero": mean.zero}[g["mean"]]
  cf = {"squared_exponential": kernel.squared_exponential, "matern32": kernel.matern32,
        "matern52": kernel.matern52}[g["cov"]]
  wf = utils.DEFAULT_WARP_FUNC
  val, grads = objectives.value_and_grad(objectives.euc, mf, cf, params, dataset, wf)
  assert abs(float(val) - g["euc"]) < 1e-10 * abs(g["euc"])
  assert abs(float(val) - float(objectives.euc(mf, cf, params, dataset, wf))) \
      < 1e-11 * abs(g["euc"])
  # the same program on the CPU stand-in (dense closed form)
  import pytest as _pt
  mp = _pt.MonkeyPatch()
  try:
    fake_engine.install(mp)
    _, grads_cpu = objectives.value_and_grad(objectives.euc, mf, cf, params, dataset, wf)
  finally:
    mp.undo()
  assert H.rel(H.grad_vec(grads, g["d"]), H.grad_vec(grads_cpu, g["d"])) < 1e-9
  v_reg, _ = objectives.value_and_grad(objectives.nll_regeuc(0.7), mf, cf, params,
                                       dataset, wf)
  v_nll, _ = objectives.value_and_grad(objectives.nll, mf, cf, params, dataset, wf)
  assert abs(float(v_reg) - (float(v_nll) + 0.7 * g["euc"])) < 1e-10 * abs(float(v_reg))


def test_training_on_nll_regeuc_decreases_it():
  """nll_regeuc(c) could not be trained in round 1 (no gradient of the Euclidean
  term); Adam on it now lowers the objective."""
  from hyperbo_b200.gp_utils import gp
  g = H.load_golden_kl("kl_m52_const_d3")
  model = H.model_from_raw(g["raw"], g["d"], g["mean"])
  params = defs.GPParams(model=dict(model))
  dataset = {k: defs.SubDataset(*v) for k, v in g["dataset"].items()}
  objective = objectives.nll_regeuc(1.0)
  params.config = {"method": "adam", "learning_rate": 1e-2, "beta": 0.9,
                   "max_training_step": 40, "batch_size": 1000,
                   "objective": objective, "alpha": 1.0}
  wf = utils.DEFAULT_WARP_FUNC
  before = float(objective(mean.constant, kernel.matern52, params, dataset, wf))
  model_ = gp.GP(dataset=dataset, mean_func=mean.constant, cov_func=kernel.matern52,
                 params=params, warp_func=wf)
  seen = []
  out = model_.train(callback=lambda i, p, l: seen.append(l))
  after = float(objective(mean.constant, kernel.matern52, out, dataset, wf))
  assert len(seen) == 40 and abs(seen[0] - before) < 1e-9 * abs(before)
  assert after < before


@pytest.mark.parametrize("case", H.load_kat_div(), ids=lambda c: "katdiv%d_%s_%s" % (
    c["id"], c["cov"], c["mean"]))
def test_divergences_match_mpmath_known_answers(case):
  """hb_nll_grad_mrhs (kl) and hb_euclid_grad (euc), value and gradient, against the
  60-digit known answers of tests/golden/make_mpmath_kat_div.py (no oracle involved)."""
  c = case
  wf = utils.DEFAULT_WARP_FUNC if c["warped"] else None
  model = H.model_from_raw(c["raw"], c["d"], c["mean"])
  params = defs.GPParams(model=dict(model))
  dataset = {k: defs.SubDataset(*v) for k, v in c["dataset"].items()}
  mf = {"constant": mean.constant, "zero": mean.zero}[c["mean"]]
  cf = {"squared_exponential": kernel.squared_exponential, "matern32": kernel.matern32,
        "matern52": kernel.matern52}[c["cov"]]
  for name, objective in (("kl", objectives.kl), ("euc", objectives.euc)):
    val, grads = objectives.value_and_grad(objective, mf, cf, params, dataset, wf)
    assert abs(float(val) - c[name]) < 1e-9 * abs(c[name]), name
    assert H.rel(H.grad_vec(grads, c["d"]), c[name + "_grad"]) < 1e-7, name
