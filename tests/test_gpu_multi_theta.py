"""Second batch axis: S hyper-parameter sets x the same data
in ONE launch sequence == S sequential calls; the acfun_test.py:74-118 scenario
(the reference vmaps the GP -> acquisition pipeline over 100 parameter vectors)
through gp.HGP."""
import numpy as np
import pytest
import torch

from oracle import hyperbo_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _raws(S, d, rng):
  base = H.raw_vec(O.init_raw_params(d), d)
  return base[None, :] + 0.3 * rng.standard_normal((S, 3 + d))


@pytest.mark.parametrize("cov", ["squared_exponential", "matern52"])
def test_nll_grad_multi_equals_sequential_calls(cov):
  from hyperbo_b200.engine import Engine, KERNEL_IDS
  eng = Engine.get()
  d, S = 4, 7
  rng = np.random.default_rng(0)
  ds_np = {t: O.make_task(t, n, d, cov) for t, n in enumerate([150, 64, 33, 200])}
  ds = eng.pack([(k, v[0], v[1]) for k, v in ds_np.items()])
  raws, mask, kid = _raws(S, d, rng), H.default_mask(d), KERNEL_IDS[cov]
  sums, nll_task = eng.nll_grad_multi(kid, 1, ds, raws, mask, want_task_nll=True)
  sums, nll_task = sums.cpu().numpy(), nll_task.cpu().numpy()
  for s in range(S):
    one, nt = eng.nll_grad(kid, 1, ds, raws[s], mask, want_task_nll=True)
    # (sequential small calls may take the launch-per-column path, whose sums are
    # ordered differently: agreement to rounding, not bitwise)
    assert H.rel(sums[s], one.cpu().numpy()) < 1e-12
    assert H.rel(nll_task[s], nt.cpu().numpy()) < 1e-12
  # and against the oracle for one of the sets
  m = H.model_from_raw(raws[3], d, "constant")
  v_ref, g_ref = O.nll_value_and_grad("constant", cov, m, ds_np, O.DEFAULT_WARP_FUNC)
  assert abs(sums[3, 0] / 4 - v_ref) < 1e-10 * abs(v_ref)
  assert H.rel(sums[3, 1:-1] / 4, H.grad_vec(g_ref, d)) < 1e-8


def test_acquisition_over_100_parameter_sets_like_acfun_test():
  from hyperbo_b200.basics import definitions as defs
  from hyperbo_b200.bo_utils import acfun
  from hyperbo_b200.gp_utils import gp, kernel, mean, utils
  d, S, nq = 2, 100, 50
  rng = np.random.default_rng(1)
  x, y = O.make_task(0, 40, d)
  xq = rng.random((nq, d))
  raws = _raws(S, d, rng)
  samples = [H.model_from_raw(r, d, "constant") for r in raws]
  dataset = {0: defs.SubDataset(x, y)}
  hgp = gp.HGP(dataset, mean.constant, kernel.squared_exponential,
               defs.GPParams(model=dict(samples[0]), samples=samples, config={}),
               utils.DEFAULT_WARP_FUNC)
  got = acfun.expected_improvement(model=hgp, sub_dataset_key=0, x_queries=xq)
  got = got.cpu().numpy().ravel()
  # S sequential single-parameter GPs (what the reference's vmap computes)
  seq = []
  for m in samples:
    g = gp.GP(dataset, mean.constant, kernel.squared_exponential,
              defs.GPParams(model=dict(m), config={}), utils.DEFAULT_WARP_FUNC)
    seq.append(acfun.expected_improvement(model=g, sub_dataset_key=0,
                                          x_queries=xq).cpu().numpy().ravel())
  want = np.mean(seq, axis=0)
  assert got.shape == (nq,)
  assert H.rel(got, want) < 1e-10
  ref = np.mean([np.ravel(O.acquisition("ei", "constant", "squared_exponential", m, {0: (x, y)},
                                        0, xq, O.DEFAULT_WARP_FUNC)) for m in samples[:10]], axis=0)
  assert H.rel(np.mean(seq[:10], axis=0), ref) < 1e-6
