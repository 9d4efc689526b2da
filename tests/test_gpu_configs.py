"""BASELINE.json configs at (or near) full size against the oracle -- the cases
round 1 only spot-checked from scripts/:

  configs[1]  256 x 512 x 8, fp64: the GRADIENT (not only values) vs the oracle
              on a 32-task shard + shard linearity up to all 256 tasks
  configs[2]  fp32 engine at n = 512
  configs[3]  24 ragged tasks (n ~ U{450..550}, d = 4, Matern-5/2): training
              steps, then EI over 10 000 candidates
  configs[4]  one n = 4096, d = 16 Matern-5/2 task: value + gradient

Tolerances as in tests/test_gpu_parity.py (fp64) / test_gpu_fp32.py (fp32).
The oracle needs ~1 minute of CPU time in total here."""
import numpy as np
import pytest
import torch

from oracle import hyperbo_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu
WF = O.DEFAULT_WARP_FUNC


def _eng(dtype=torch.float64):
  from hyperbo_b200.engine import Engine
  return Engine.get(dtype=dtype)


def _ids(cov, mean):
  from hyperbo_b200.engine import KERNEL_IDS, MEAN_IDS
  return KERNEL_IDS[cov], MEAN_IDS[mean]


def _pack(eng, ds):
  return eng.pack([(k, v[0], v[1]) for k, v in ds.items()])


def test_c2_full_gradient_256x512x8():
  eng = _eng()
  T, n, d = 256, 512, 8
  ds = {t: O.make_task(t, n, d) for t in range(32)}          # GP draws (8d)
  rng = np.random.default_rng(1)
  for t in range(32, T):                                      # cheap filler tasks
    x = rng.random((n, d))
    ds[t] = (x, 5.0 + np.sin(3 * x.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((n, 1)))
  model = O.init_raw_params(d)
  model["lengthscale"] = np.linspace(-0.3, 0.4, d)
  raw, mask = H.raw_vec(model, d), H.default_mask(d)
  kid, mid = _ids("squared_exponential", "constant")
  shard = {t: ds[t] for t in range(32)}
  s_shard = eng.nll_grad(kid, mid, _pack(eng, shard), raw, mask).cpu().numpy()
  v_ref, g_ref = O.nll_value_and_grad("constant", "squared_exponential", model, shard, WF)
  assert abs(s_shard[0] / 32 - v_ref) < 1e-10 * abs(v_ref)
  assert H.rel(s_shard[1:-1] / 32, H.grad_vec(g_ref, d)) < 1e-8
  # all 256 tasks in one call == sum of eight 32-task calls (incl. the gradient)
  s_full = eng.nll_grad(kid, mid, _pack(eng, ds), raw, mask).cpu().numpy()
  parts = sum(eng.nll_grad(kid, mid, _pack(eng, {t: ds[t] for t in range(32 * r, 32 * r + 32)}),
                           raw, mask).cpu().numpy() for r in range(8))
  assert s_full[-1] == T
  assert H.rel(s_full, parts) < 1e-12
  assert np.array_equal(parts[:0], s_full[:0]) and H.rel(parts[:12], s_full[:12]) < 1e-12


def test_c3_fp32_engine_n512():
  eng = _eng(torch.float32)
  n, d = 512, 8
  ds = {t: O.make_task(t, n, d) for t in range(4)}
  model = O.init_raw_params(d)
  kid, mid = _ids("squared_exponential", "constant")
  sums = eng.nll_grad(kid, mid, _pack(eng, ds), H.raw_vec(model, d),
                      H.default_mask(d)).double().cpu().numpy()
  v_ref, g_ref = O.nll_value_and_grad("constant", "squared_exponential", model, ds, WF)
  # fp32 tolerance 1e-5 relative to the SIZE OF THE NLL'S TERMS: on GP-draw data
  # the quadratic term, the log-determinant and .5 n log(2 pi) = 470 nearly cancel
  # (nll = -66.6), so the value itself is no measure of the arithmetic's scale
  scale = abs(v_ref) + 0.5 * n * np.log(2 * np.pi)
  assert abs(sums[0] / 4 - v_ref) < 1e-5 * scale
  assert H.rel(sums[1:-1] / 4, H.grad_vec(g_ref, d)) < 1e-3


def test_c4_ragged_pd1_shape_train_then_ei_10k():
  from hyperbo_b200.gp_utils.gp import AdamTrainer
  eng = _eng()
  d, T = 4, 24
  ns = np.random.Generator(np.random.PCG64(7)).integers(450, 551, T)
  ds = {t: O.make_task(t, int(ns[t]), d, "matern52") for t in range(T)}
  model = O.init_raw_params(d)
  kid, mid = _ids("matern52", "constant")
  raw0, mask = H.raw_vec(model, d), H.default_mask(d)
  tr = AdamTrainer(eng, kid, mid, raw0, mask, d, 1e-3)
  packed = _pack(eng, ds)
  losses = []
  for i in range(5):
    tr.step(packed, use_graph=i >= 2)
    losses.append(tr.loss())
  ref_model, ref_losses = O.infer_parameters_adam("constant", "matern52", model, ds, WF,
                                                  1e-3, 5, 10**6)
  assert H.rel(losses, ref_losses) < 1e-9
  raw = tr.raw.cpu().numpy()
  assert H.rel(raw, H.raw_vec(ref_model, d)) < 1e-8
  # EI over 10 000 candidates on the query task t = 0 (GP.predict conventions)
  xq = np.random.Generator(np.random.PCG64(9)).random((10000, d))
  cache, _, _, _, info = eng.build_predictor(kid, mid, ds[0][0], ds[0][1], raw, mask)
  assert int(info[0]) == 0
  target = float(np.max(ds[0][1]))
  mu, var, ei = eng.predict(kid, mid, eng.tensor(ds[0][0]), cache, raw, mask, xq,
                            noise_flag=1.0, var_scale=T / (T - 1.0), acq_id=1,
                            acq_param=target)
  m = H.model_from_raw(raw, d, "constant")
  mu_ref, var_ref = O.gp_predict("constant", "matern52", m, ds, xq, 0, WF)
  ei_ref = O.acquisition("ei", "constant", "matern52", m, ds, 0, xq, WF)
  assert H.rel(mu.cpu().numpy().ravel(), np.ravel(mu_ref)) < 1e-6
  assert H.rel(var.cpu().numpy().ravel(), np.ravel(var_ref)) < 1e-6
  assert H.rel(ei.cpu().numpy().ravel(), np.ravel(ei_ref)) < 1e-6
  assert int(np.argmax(ei.cpu().numpy())) == int(np.argmax(ei_ref))


def test_c5_one_task_n4096_d16_matern52():
  eng = _eng()
  n, d = 4096, 16
  x, y = O.make_task(0, n, d, "matern52", surrogate=True)  # C5 surrogate
  model = O.init_raw_params(d)
  kid, mid = _ids("matern52", "constant")
  sums = eng.nll_grad(kid, mid, _pack(eng, {0: (x, y)}), H.raw_vec(model, d),
                      H.default_mask(d)).cpu().numpy()
  v_ref, g_ref = O.nll_value_and_grad("constant", "matern52", model, {0: (x, y)}, WF)
  assert abs(sums[0] - v_ref) < 1e-10 * abs(v_ref)
  assert H.rel(sums[1:-1], H.grad_vec(g_ref, d)) < 1e-8


def test_graph_survives_engine_use_between_steps():
  """a callback that uses the engine between replayed steps (predict
  on a LARGER task: workspace re-allocation + plan eviction) must not leave the
  trainer replaying a graph with stale workspace pointers."""
  from hyperbo_b200.gp_utils.gp import AdamTrainer
  eng = _eng()
  d = 3
  ds_np = {t: O.make_task(t, 60 + 10 * t, d) for t in range(3)}
  model = O.init_raw_params(d)
  kid, mid = _ids("squared_exponential", "constant")
  raw0, mask = H.raw_vec(model, d), H.default_mask(d)
  tr = AdamTrainer(eng, kid, mid, raw0, mask, d, 1e-2)
  ds = _pack(eng, ds_np)
  losses = []
  for i in range(8):
    tr.step(ds, use_graph=True)
    losses.append(tr.loss())
    if i in (2, 5):  # what a callback may do: factorise much larger tasks
      for k in range(5):  # > the 4 cached plans: evicts the trainer's plan
        big = O.make_task(90 + k, 700 + 64 * k + 200 * i, d)
        eng.build_predictor(kid, mid, big[0], big[1], raw0, mask)
  _, ref_losses = O.infer_parameters_adam("constant", "squared_exponential", model,
                                          ds_np, WF, 1e-2, 8, 10**6)
  assert H.rel(losses, ref_losses) < 1e-9
