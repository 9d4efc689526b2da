"""Device-resident BO loop (hb_bo_init / hb_bo_step):
rank-1 append of the packed inverse factor == full refactorisation, fused
arg-max == host arg-max, and bayesopt.simulated_bayesopt's fast path picks the
same candidates as the reference-shaped host loop (bayesopt.py:169-193)."""
import numpy as np
import pytest
import torch

from oracle import hyperbo_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu
WF = O.DEFAULT_WARP_FUNC


@pytest.mark.parametrize("cov,n0,d", [("matern52", 50, 4), ("squared_exponential", 0, 2),
                                      ("matern32", 63, 3), ("squared_exponential", 120, 8)])
def test_30_appends_equal_full_refactorisation(cov, n0, d):
  from hyperbo_b200.engine import BoSession, Engine, KERNEL_IDS
  eng = Engine.get()
  rng = np.random.default_rng(3)
  iters, nq = 30, 700
  x0 = rng.random((n0, d))
  y0 = 5 + np.sin(3 * x0.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((n0, 1))
  xq = rng.random((nq, d))
  yq = 5 + np.sin(3 * xq.sum(1)) + 0.1 * rng.standard_normal(nq)
  model = O.init_raw_params(d)
  raw, mask = H.raw_vec(model, d), H.default_mask(d)
  kid = KERNEL_IDS[cov]
  sess = BoSession(eng, kid, 1, x0 if n0 else None, y0 if n0 else None, raw, mask,
                   n0 + iters, d=d)
  xq_d, yq_d = eng.tensor(xq), eng.tensor(yq)
  # host-side replay with the ORACLE: EI over all candidates, arg-max, append
  xs, ys = x0.copy(), y0.copy()
  for it in range(iters):
    sess.step(xq_d, yq_d, 1, 0.0, True, noise_flag=1.0, var_scale=1.0)
    ds = {0: (xs, ys)} if len(xs) else {}
    ei = O.acquisition("ei", "constant", cov, model, ds, 0, xq, WF)
    pick = int(np.argmax(ei))
    got = int(sess.selected()[it])
    if got != pick:  # only acceptable for a numerical tie
      top = np.sort(np.ravel(ei))[-2:]
      assert abs(ei.ravel()[got] - ei.ravel()[pick]) <= 1e-9 * abs(top[-1]), (it, got, pick)
      pick = got
    xs = np.vstack([xs, xq[pick:pick + 1]])
    ys = np.vstack([ys, [[yq[pick]]]])
  # the appended factor predicts like a factorisation from scratch
  xo, yo = sess.observations()
  assert np.array_equal(xo.cpu().numpy(), xs) and np.array_equal(yo.cpu().numpy(), ys)
  xt = rng.random((64, d))
  mu, var, _ = sess.predict(xt, noise_flag=1.0)
  mu_ref, var_ref = O.gp_predict("constant", cov, model, {0: (xs, ys)}, xt, 0, WF,
                                 unbiased=False)
  assert H.rel(mu.cpu().numpy().ravel(), np.ravel(mu_ref)) < 1e-9
  assert H.rel(var.cpu().numpy().ravel(), np.ravel(var_ref)) < 1e-9


@pytest.mark.parametrize("acname", ["expected_improvement", "ucb", "probability_of_improvement"])
def test_simulated_bayesopt_device_loop_matches_host_loop(acname, monkeypatch):
  from hyperbo_b200.basics import definitions as defs
  from hyperbo_b200.bo_utils import acfun, bayesopt
  from hyperbo_b200.gp_utils import gp, kernel, mean, utils
  d = 3
  rng = np.random.default_rng(11)
  train = {t: O.make_task(t, 40, d, "matern52") for t in range(3)}
  xq = rng.random((500, d))
  yq = 5 + np.sin(3 * xq.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((500, 1))
  ac = getattr(acfun, acname)

  def run(device):
    monkeypatch.setenv("HB_BO_DEVICE", "1" if device else "0")
    dataset = {k: defs.SubDataset(*v) for k, v in train.items()}
    dataset["query"] = defs.SubDataset(xq[:5], yq[:5])
    params = defs.GPParams(model=dict(O.init_raw_params(d)), config={})
    model = gp.GP(dataset, mean.constant, kernel.matern52, params, utils.DEFAULT_WARP_FUNC)
    sub = bayesopt.simulated_bayesopt(model, "query", defs.SubDataset(xq, yq), ac, 12)
    return np.asarray(torch.as_tensor(sub.x).cpu()), np.asarray(torch.as_tensor(sub.y).cpu())

  xh, yh = run(False)
  xd, yd = run(True)
  assert xh.shape == (17, d) and xd.shape == (17, d)
  assert np.array_equal(xh, xd) and np.array_equal(yh, yd)
