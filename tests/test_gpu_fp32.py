"""fp32 engine (HB_F32: the reference's JAX default dtype; tile products on the
tensor pipe as 3xTF32) against the fp64 oracle.

Tolerances (SURVEY.md 8d, fp32 engine vs fp64 oracle): nll 1e-5 relative (2e-5
for the tiny n = 40 case), gradients 1e-3, chol 1e-4, alpha 1e-3, predictions
and acquisition 1e-3 relative; measured: nll ~2e-6, grad ~3e-5, mu ~1e-5,
var ~5e-5, EI ~1e-4.  The fp64 engine keeps the tight tolerances of
test_gpu_parity.py."""
import numpy as np
import pytest
import torch

from oracle import hyperbo_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu
WF = O.DEFAULT_WARP_FUNC


@pytest.fixture(scope="module")
def eng():
  from hyperbo_b200.engine import Engine
  return Engine.get(dtype=torch.float32)


def _ids(cov, mean):
  from hyperbo_b200.engine import KERNEL_IDS, MEAN_IDS
  return KERNEL_IDS[cov], MEAN_IDS[mean]


def _pack(eng, ds):
  return eng.pack([(k, v[0], v[1]) for k, v in ds.items()])


@pytest.mark.parametrize("name", H.golden_cases())
def test_golden_fixtures_fp32(eng, name):
  g = H.load_golden(name)
  d, ns = g["d"], g["ns"]
  kid, mid = _ids(g["cov"], g["mean"])
  mask = H.default_mask(d)
  ds = _pack(eng, g["dataset"])
  assert ds.x.dtype == torch.float32
  chols, alpha, nll, info = eng.factorize(kid, mid, ds, g["raw"], mask)
  assert chols[0].dtype == torch.float32 and info.tolist() == [0] * len(ns)
  assert H.rel(nll.double().cpu().numpy(), g["nll_task"]) < 2e-5
  assert H.rel(chols[0].double().cpu().numpy(), g["chol0"]) < 1e-4
  assert H.rel(alpha[:ns[0]].double().cpu().numpy(), g["alpha0"]) < 1e-3
  sums = eng.nll_grad(kid, mid, ds, g["raw"], mask).double().cpu().numpy()
  T = len(ns)
  assert sums[-1] == T
  # (2e-5 of the SIZE OF THE NLL'S TERMS: with n ~ 100 the value itself nearly
  # cancels against .5 n log(2 pi), so it is no measure of the arithmetic's scale)
  scale_nll = abs(g["mean_nll"]) + 0.5 * np.mean(ns) * np.log(2 * np.pi)
  assert abs(sums[0] / T - g["mean_nll"]) < 2e-5 * scale_nll
  assert H.rel(sums[1:-1] / T, g["grad"]) < 1e-3
  cache, chol, kinvy, _, _ = eng.build_predictor(kid, mid, g["x0"], g["y0"],
                                                 g["raw"], mask)
  scale = T / (T - 1.0) if T > 1 else 1.0
  mu, var, acq = eng.predict(kid, mid, eng.tensor(g["x0"]), cache, g["raw"],
                             mask, g["xq"], noise_flag=1.0, var_scale=scale,
                             acq_id=1, acq_param=float(np.max(g["y0"])))
  assert H.rel(mu.double().cpu().numpy().ravel(), g["mu"]) < 1e-3
  assert H.rel(var.double().cpu().numpy().ravel(), g["var"]) < 1e-3
  assert H.rel(acq.double().cpu().numpy().ravel(), g["ei"]) < 2e-3


@pytest.mark.parametrize("cov", O.KERNELS)
def test_nll_grad_fp32_c2_shape_tasks(eng, cov):
  ns, d = [512, 300, 65], 8
  ds_np = {t: O.make_task(t, n, d, cov) for t, n in enumerate(ns)}
  model = O.init_raw_params(d)
  kid, mid = _ids(cov, "constant")
  sums = eng.nll_grad(kid, mid, _pack(eng, ds_np), H.raw_vec(model, d),
                      H.default_mask(d)).double().cpu().numpy()
  v_ref, g_ref = O.nll_value_and_grad("constant", cov, model, ds_np, WF)
  # the NLL is a sum of O(n) terms that partly cancel: 1e-5 of that magnitude
  assert abs(sums[0] / 3 - v_ref) < 1e-5 * max(abs(v_ref), float(np.mean(ns)))
  assert H.rel(sums[1:-1] / 3, H.grad_vec(g_ref, d)) < 1e-3


def test_fp32_no_worse_than_numpy_fp32_restatement(eng):
  """The engine's fp32 NLL must be at least as close to the fp64 truth as a
  plain NumPy float32 restatement of the reference (what JAX-default computes)."""
  x, y = O.make_task(4, 300, 4)
  model = O.init_raw_params(4)
  ref64 = O.nll_sub_dataset("constant", "squared_exponential", model, x, y, WF)
  # numpy float32 restatement
  ls = np.float32(O.default_softplus(0.0))
  sv, nv = np.float32(O.default_softplus(0.0)), np.float32(O.default_softplus(-4.0))
  xs = (x.astype(np.float32) / ls)
  diff = xs[:, None, :] - xs[None, :, :]
  k = sv * np.exp(-np.sum(diff * diff, axis=-1, dtype=np.float32) / np.float32(2))
  k = (k + np.eye(300, dtype=np.float32) * (nv + np.float32(1e-6))).astype(np.float32)
  chol = np.linalg.cholesky(k)
  r = (y.astype(np.float32) - np.float32(5.1))
  import scipy.linalg as spla
  a = spla.cho_solve((chol, True), r).astype(np.float32)
  nll32 = float(np.float32(0.5) * (r.T @ a).item() + np.sum(np.log(np.diag(chol)))
                + np.float32(0.5 * 300 * np.log(2 * np.pi)))
  kid, mid = _ids("squared_exponential", "constant")
  _, _, nll, _ = eng.factorize(kid, mid, _pack(eng, {0: (x, y)}),
                               H.raw_vec(model, 4), H.default_mask(4),
                               want_chol=False, want_alpha=False)
  err_engine = abs(nll[0].item() - ref64)
  err_numpy32 = abs(nll32 - ref64)
  assert err_engine <= max(2.0 * err_numpy32, 1e-5 * abs(ref64))


def test_adam_loop_fp32_tracks_oracle(eng):
  from hyperbo_b200.gp_utils.gp import AdamTrainer
  d = 3
  ds_np = {t: O.make_task(t, 60 + 40 * t, d) for t in range(4)}
  model = O.init_raw_params(d)
  kid, mid = _ids("squared_exponential", "constant")
  tr = AdamTrainer(eng, kid, mid, H.raw_vec(model, d), H.default_mask(d), d, 1e-2)
  assert tr.raw.dtype == torch.float32
  ds = _pack(eng, ds_np)
  losses = []
  for i in range(6):
    tr.step(ds, use_graph=(i >= 2))
    losses.append(tr.loss())
  _, ref = O.infer_parameters_adam("constant", "squared_exponential", model,
                                   ds_np, WF, 1e-2, 6, 10**6)
  assert H.rel(losses, ref) < 1e-4 and losses[-1] < losses[0]


def test_api_default_dtype_switch():
  from hyperbo_b200 import engine
  from hyperbo_b200.basics import definitions as defs
  from hyperbo_b200.gp_utils import kernel, mean, objectives, utils
  ds_np = O.make_dataset(3, 50, 2)
  dataset = {k: defs.SubDataset(*v) for k, v in ds_np.items()}
  params = defs.GPParams(model=dict(O.init_raw_params(2)))
  ref = O.neg_log_marginal_likelihood("constant", "squared_exponential",
                                      O.init_raw_params(2), ds_np, WF)
  try:
    engine.set_default_dtype(torch.float32)
    v32 = objectives.nll(mean.constant, kernel.squared_exponential, params,
                         dataset, utils.DEFAULT_WARP_FUNC)
    assert v32.dtype == torch.float32 and abs(float(v32) - ref) < 2e-5 * abs(ref)
  finally:
    engine.set_default_dtype(torch.float64)
  v64 = objectives.nll(mean.constant, kernel.squared_exponential, params,
                       dataset, utils.DEFAULT_WARP_FUNC)
  assert v64.dtype == torch.float64 and abs(float(v64) - ref) < 1e-10 * abs(ref)
