"""GP linear algebra -- mirrors hyperbo/basics/linalg.py:29-110 on the engine."""
from __future__ import annotations

import torch

from hyperbo_b200 import engine as _engine
from hyperbo_b200.basics import params_utils
from hyperbo_b200.gp_utils import kernel as _kernel
from hyperbo_b200.gp_utils import mean as _mean

EPS = 1e-10


def _col(y):
  y = torch.as_tensor(y)
  return y.reshape(y.shape[0], -1)


def compute_delta_y_and_cov(mean_func, cov_func, params, x, y, warp_func=None,
                            eps=1e-6):
  """y - mu(x) and cov(x,x) + I*(sigma^2 + eps)  (linalg.py:36-69)."""
  kid = _kernel.kernel_id_of(cov_func)
  need_mean = _mean.mean_id_of(mean_func) == 1
  eng = _engine.Engine.get()
  x = eng.tensor(x)
  raw, mask, _ = params_utils.pack_raw(params.model, x.shape[1], need_mean,
                                       warp_func)
  cov = eng.kernel_matrix(kid, x, None, raw, mask, add_noise=True, jitter=eps)
  dy = eng.tensor(_col(y)) - mean_func(params, x, warp_func=warp_func).to(
      eng.device)
  return dy, cov


def solve_gp_linear_system(mean_func, cov_func, params, x, y, warp_func=None,
                           eps=1e-6, return_cache=False):
  """Solve m + K v = y with the Cholesky factor of K = cov(x,x) + I*(noise+eps)
  (linalg.py:72-110).  Returns (chol, kinvy, y - mean); with return_cache also
  the engine's packed predictor cache."""
  if eps != 1e-6:
    raise NotImplementedError("the engine uses the reference's eps=1e-6 jitter")
  kid = _kernel.kernel_id_of(cov_func)
  mid = _mean.mean_id_of(mean_func)
  eng = _engine.Engine.get()
  x = eng.tensor(x)
  y = eng.tensor(_col(y))
  if y.shape[1] != 1:
    raise NotImplementedError("the hot path handles y with one column (m=1)")
  raw, mask, _ = params_utils.pack_raw(params.model, x.shape[1], mid == 1,
                                       warp_func)
  cache, chol, kinvy, _, _ = eng.build_predictor(kid, mid, x, y, raw, mask)
  dy = y - mean_func(params, x, warp_func=warp_func).to(eng.device)
  if return_cache:
    return chol, kinvy, dy, cache
  return chol, kinvy, dy


def solve_linear_system(coeff, b):
  raise NotImplementedError(
      "solve_linear_system on an explicit matrix is not part of the engine's "
      "hot path; use solve_gp_linear_system (the kernel matrix is built and "
      "factorised on the GPU without being materialised)")
