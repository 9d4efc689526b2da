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


# ---- explicit-matrix helpers (linalg.py:29-33,112-197) ----------------------
# Off the hot path: the engine never materialises K (solve_gp_linear_system
# builds and factorises it tile by tile).  These operate on matrices the caller
# already holds (e.g. the data-side moments of the KL objective) and are plain
# torch calls (cuSOLVER / cuBLAS on CUDA tensors).
def cholesky_cache(spd_matrix, cached_cholesky):
  """Cholesky factor of `spd_matrix` unless one is given (linalg.py:129-136)."""
  if cached_cholesky is not None:
    return cached_cholesky
  return torch.linalg.cholesky(spd_matrix)


def inverse_spdmatrix_vector_product(spd_matrix, x, cached_cholesky=None):
  """spd_matrix^-1 x through the Cholesky factor (linalg.py:139-145)."""
  chol = cholesky_cache(spd_matrix, cached_cholesky)
  x = torch.as_tensor(x, dtype=chol.dtype, device=chol.device)
  if x.dim() == 1:
    return torch.cholesky_solve(x[:, None], chol)[:, 0]
  return torch.cholesky_solve(x, chol)


def solve_linear_system(coeff, b):
  """Solve A x = b for SPD A = coeff -> (chol, x)  (linalg.py:29-33)."""
  chol = torch.linalg.cholesky(coeff)
  return chol, inverse_spdmatrix_vector_product(coeff, b, cached_cholesky=chol)


def svd_matrix_sqrt(cov):
  """A with A A' = cov and full column rank (linalg.py:112-126)."""
  u, s, _ = torch.linalg.svd(cov)
  factor = u * torch.sqrt(s)[None, :]
  tol = s.max() * torch.finfo(s.dtype).eps / 2.0 * (2 * cov.shape[0] + 1.0)**0.5
  rank = int((s > tol).sum())
  return factor[:, :rank]


def safe_l2norm(x):
  """l2 norm (linalg.py:194-197; the custom gradient at 0 only matters to
  autodiff, which the engine replaces by closed forms)."""
  return torch.sqrt(torch.sum(torch.as_tensor(x)**2))
