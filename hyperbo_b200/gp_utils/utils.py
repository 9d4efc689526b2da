"""Warp functions and MVN distances -- mirrors hyperbo/gp_utils/utils.py:27-173.

The engine differentiates the objective in closed form on the GPU, so it has to
*recognise* a warp rather than trace it: the callables below carry an `hb_warp`
tag ('identity' or 'softplus_eps').  They are also ordinary torch callables.
"""
from __future__ import annotations

import torch

from hyperbo_b200.basics import data_utils as _data_utils
from hyperbo_b200.basics import definitions as _defs

EPS = 1e-10  # utils.py:28
SubDataset = _defs.SubDataset  # utils.py:23
sub_sample_dataset_iterator = _data_utils.sub_sample_dataset_iterator  # utils.py:32


def identity_warp(x):
  return x


identity_warp.hb_warp = "identity"


def softplus_warp(x):  # utils.py:29 (jax.nn.softplus); not used by the default
  return torch.nn.functional.softplus(torch.as_tensor(x, dtype=torch.float64))


softplus_warp.hb_warp = "softplus"


def DEFAULT_SOFTPLUS(x):  # utils.py:73
  return torch.nn.functional.softplus(
      torch.as_tensor(x, dtype=torch.float64)) + EPS


DEFAULT_SOFTPLUS.hb_warp = "softplus_eps"


def squareplus_warp(x):  # utils.py:30 -- a plain callable: the engine does not
  x = torch.as_tensor(x, dtype=torch.float64)  # differentiate it (warp_kind raises)
  return 0.5 * (x + torch.sqrt(x**2 + 4))

# utils.py:75-81
DEFAULT_WARP_FUNC = {
    "constant": identity_warp,
    "lengthscale": DEFAULT_SOFTPLUS,
    "signal_variance": DEFAULT_SOFTPLUS,
    "noise_variance": DEFAULT_SOFTPLUS,
    "dot_prod_sigma": DEFAULT_SOFTPLUS,
}


def warp_kind(warp_func, key: str) -> str:
  """'identity' | 'softplus_eps' for `key`; raises for warps the engine cannot
  differentiate (anything that is not one of the tagged callables)."""
  if not warp_func or key not in warp_func:
    return "identity"
  kind = getattr(warp_func[key], "hb_warp", None)
  if kind in ("identity", "softplus_eps"):
    return kind
  raise NotImplementedError(
      f"warp function for '{key}' is not one of hyperbo_b200.gp_utils.utils."
      "{identity_warp, DEFAULT_SOFTPLUS}; the engine differentiates in closed "
      "form and cannot trace arbitrary Python warps")


# ------------------------------------------------- MVN distances (:84-173) --
# objectives.multivariate_normal_divergence selects its arithmetic by the
# `distance` callable.  The engine evaluates the partial KL through its batched
# factorisation (see objectives.py), so -- like the warps above -- the distance
# has to be RECOGNISED: the callables carry an `hb_distance` tag, and
# functools.partial(...) of them with the reference's keyword arguments
# (weight / eps / partial, mean_weight / cov_weight) is understood.  Called
# directly with explicit moments they are ordinary torch (device) functions.
def partial_kl_mvn(mu0, cov0, mu1, cov1):
  """utils.py:84-106 on explicit moments (torch; cuSOLVER Cholesky)."""
  mu_diff = mu1 - mu0
  chol1 = torch.linalg.cholesky(cov1)
  tr = torch.trace(torch.cholesky_solve(cov0, chol1))
  mah = mu_diff @ torch.cholesky_solve(mu_diff[:, None], chol1)[:, 0]
  return tr + mah + torch.sum(2 * torch.log(torch.diagonal(chol1)))


def svd_matrix_sqrt(cov):
  """basics/linalg.py:112-126."""
  from hyperbo_b200.basics import linalg as _linalg
  return _linalg.svd_matrix_sqrt(cov)


def kl_multivariate_normal(mu0, cov0, mu1, cov1, weight=1.0, eps=0.0,
                           partial=True):
  """utils.py:109-148."""
  cov0 = torch.atleast_2d(cov0)
  cov1 = torch.atleast_2d(cov1)
  if eps > 0.0:
    eye = torch.eye(cov0.shape[0], device=cov0.device, dtype=cov0.dtype)
    cov0 = cov0 + eye * eps
    cov1 = cov1 + eye * eps
  if partial:
    return weight * partial_kl_mvn(mu0, cov0, mu1, cov1)
  chol0 = svd_matrix_sqrt(cov0)
  chol0inv = torch.linalg.pinv(chol0)
  mu1 = chol0inv @ (mu1 - mu0)
  cov1 = chol0inv @ cov1 @ chol0inv.T
  mu0 = torch.zeros_like(mu1)
  cov0 = torch.eye(cov1.shape[0], device=cov1.device, dtype=cov1.dtype)
  return weight * 0.5 * (partial_kl_mvn(mu0, cov0, mu1, cov1) - chol0.shape[1])


kl_multivariate_normal.hb_distance = "kl"


def euclidean_multivariate_normal(mu0, cov0, mu1, cov1, mean_weight=1.0,
                                  cov_weight=1.0, **unused_kwargs):
  """utils.py:151-173."""
  mean_diff = torch.sqrt(torch.sum((mu0 - mu1)**2))
  cov_diff = torch.sqrt(torch.sum((torch.atleast_2d(cov0) - cov1)**2))
  return mean_weight * mean_diff + cov_weight * cov_diff


euclidean_multivariate_normal.hb_distance = "euc"


def distance_spec(distance):
  """(kind, kwargs) of a distance callable: 'kl' | 'euc' plus the keyword
  arguments bound by functools.partial; raises for anything else."""
  import functools
  kw = {}
  while isinstance(distance, functools.partial):
    if distance.args:
      raise NotImplementedError("positional functools.partial of a distance")
    kw = {**distance.keywords, **kw}
    distance = distance.func
  kind = getattr(distance, "hb_distance", None)
  if kind not in ("kl", "euc"):
    raise NotImplementedError(
        "distance must be hyperbo_b200.gp_utils.utils.kl_multivariate_normal or "
        "euclidean_multivariate_normal (optionally through functools.partial)")
  return kind, kw
