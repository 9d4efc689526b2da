"""gp.predict(..., full_cov=True) / GP.predict(full_cov=True) on hb_predict_cov
(gp_utils/gp.py:295-300, :607-619 of the reference): V = L^-1 K* and K** - V'V on
the tensor pipe, against the CPU oracle."""
import numpy as np
import pytest
import torch

from hyperbo_b200 import engine as _engine
from hyperbo_b200.basics import definitions as defs
from hyperbo_b200.gp_utils import gp, kernel, mean, utils
from oracle import hyperbo_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu
COVS = {"squared_exponential": kernel.squared_exponential,
        "matern32": kernel.matern32, "matern52": kernel.matern52}


def _np(t):
  return t.detach().cpu().numpy().astype(np.float64)


@pytest.mark.parametrize("cov", sorted(COVS))
@pytest.mark.parametrize("n,nq,d", [(20, 10, 1), (130, 70, 4), (500, 300, 3),
                                    (64, 64, 2), (257, 129, 9)])
def test_full_cov_matches_oracle(cov, n, nq, d):
  x, y = O.make_task(5, n, d, cov)
  xq = np.random.default_rng(1).uniform(size=(nq, d))
  model = {"constant": 0.7, "lengthscale": np.linspace(0.6, 1.3, d),
           "signal_variance": 1.1, "noise_variance": 0.02}
  params = defs.GPParams(model=dict(model))
  mu, c = gp.predict(mean.constant, COVS[cov], params, x, y, xq,
                     full_cov=True)
  mu_o, c_o = O.predict("constant", cov, model, x, y, xq, full_cov=True)
  assert mu.shape == (nq, 1) and c.shape == (nq, nq)
  assert H.rel(_np(mu), mu_o) < 1e-6            # north_star: 1e-6 relative
  scale = np.abs(c_o).max()
  assert np.abs(_np(c) - c_o).max() < 1e-8 * scale
  assert np.array_equal(_np(c), _np(c).T)       # both triangles written
  # diagonal == the variance path
  _, var = gp.predict(mean.constant, COVS[cov], params, x, y, xq)
  assert np.abs(np.diag(_np(c)) - _np(var).ravel()).max() < 1e-9 * scale


def test_gp_predict_full_cov_noise_unbiased_and_prior():
  d = 3
  ds_np = {t: O.make_task(t, n, d, "matern52") for t, n in enumerate([90, 40, 70])}
  dataset = {t: defs.SubDataset(x, y) for t, (x, y) in ds_np.items()}
  model = {"constant": 0.2, "lengthscale": np.array([0.8, 1.0, 1.2]),
           "signal_variance": 0.9, "noise_variance": 0.05}
  m = gp.GP(dataset=dataset, mean_func=mean.constant, cov_func=kernel.matern52,
            params=defs.GPParams(model=dict(model)))
  xq = np.random.default_rng(3).uniform(size=(75, d))
  ds_o = {t: (x, y) for t, (x, y) in ds_np.items()}
  for key in (1, "unknown"):
    for with_noise in (True, False):
      for unbiased in (True, False):
        mu, c = m.predict(xq, key, full_cov=True, with_noise=with_noise,
                          unbiased=unbiased)
        mu_o, c_o = O.gp_predict("constant", "matern52", model, ds_o, xq, key,
                                 full_cov=True, with_noise=with_noise,
                                 unbiased=unbiased)
        assert H.rel(_np(mu), mu_o) < 1e-6
        assert np.abs(_np(c) - c_o).max() < 1e-8 * np.abs(c_o).max()


def test_full_cov_fp32_engine():
  prev = _engine.get_default_dtype()
  _engine.set_default_dtype(torch.float32)
  try:
    n, nq, d = 200, 100, 4
    x, y = O.make_task(2, n, d, "matern52")
    xq = np.random.default_rng(4).uniform(size=(nq, d))
    model = {"constant": 0.7, "lengthscale": np.ones(d), "signal_variance": 1.0,
             "noise_variance": 0.05}
    mu, c = gp.predict(mean.constant, kernel.matern52, defs.GPParams(model=model),
                       x, y, xq, full_cov=True)
    mu_o, c_o = O.predict("constant", "matern52", model, x, y, xq,
                          full_cov=True)
    assert c.dtype == torch.float32
    assert H.rel(_np(mu), mu_o) < 1e-4
    assert np.abs(_np(c) - c_o).max() < 2e-4 * np.abs(c_o).max()
  finally:
    _engine.set_default_dtype(prev)
