"""Objective functions -- mirrors hyperbo/gp_utils/objectives.py:109-210
(neg_log_marginal_likelihood, Cholesky branch) on the engine, plus the
value-and-gradient entry that replaces jax.value_and_grad (gp.py:134)."""
from __future__ import annotations

from typing import Dict, Tuple

import torch

from hyperbo_b200 import engine as _engine
from hyperbo_b200.basics import params_utils
from hyperbo_b200.gp_utils import kernel as _kernel
from hyperbo_b200.gp_utils import mean as _mean

retrieve_params = params_utils.retrieve_params


def _select(dataset, exclude_aligned=True):
  """Task filter of objectives.py:181-185 (skip aligned and empty)."""
  out = []
  for k, s in dataset.items():
    aligned = s[2] if len(s) > 2 else None
    if exclude_aligned and aligned is not None:
      continue
    if torch.as_tensor(s[0]).shape[0] == 0:
      continue
    y = torch.as_tensor(s[1])
    if y.dim() == 2 and y.shape[1] != 1:
      raise NotImplementedError(
          "the engine's NLL handles y with one column (m=1); "
          f"dataset[{k}].y has shape {tuple(y.shape)}")
    out.append((k, s[0], s[1]))
  return out


def _prepare(mean_func, cov_func, params, dataset, warp_func, exclude_aligned):
  kid = _kernel.kernel_id_of(cov_func)
  mid = _mean.mean_id_of(mean_func)
  eng = _engine.Engine.get()
  ds = eng.pack(_select(dataset, exclude_aligned))
  d = ds.d if ds.num_tasks else 1
  raw, mask, _ = params_utils.pack_raw(params.model, d, mid == 1, warp_func)
  return eng, kid, mid, ds, raw, mask


def neg_log_marginal_likelihood(mean_func, cov_func, params, dataset,
                                warp_func=None, exclude_aligned=True,
                                return_key2nll=False, use_cholesky=True):
  """Negative log marginal likelihood of a (multi-task) GP, averaged over the
  non-empty, non-aligned sub-datasets (objectives.py:109-210)."""
  if not use_cholesky:
    raise NotImplementedError(
        "the SVD branch (objectives.py:157-176) is not on the engine's hot "
        "path; use use_cholesky=True")
  if "priors" in params.config:
    raise NotImplementedError("log-prior terms (objectives.py:197-207)")
  eng, kid, mid, ds, raw, mask = _prepare(mean_func, cov_func, params, dataset,
                                          warp_func, exclude_aligned)
  if ds.num_tasks == 0:
    total = torch.zeros((), device=eng.device, dtype=eng.dtype)
    return (total, {}) if return_key2nll else total
  _, _, nll, _ = eng.factorize(kid, mid, ds, raw, mask, want_chol=False,
                               want_alpha=False)
  total = nll.sum() / ds.num_tasks
  if return_key2nll:
    return total, {k: nll[i] for i, k in enumerate(ds.keys)}
  return total


def nll_value_and_grad(mean_func, cov_func, params, dataset, warp_func=None,
                       exclude_aligned=True) -> Tuple[torch.Tensor, Dict]:
  """(mean NLL, d mean NLL / d params.model) -- what
  jax.value_and_grad(loss_func)(model_param, batch) returns at gp.py:134.  The
  gradient dict has the keys and shapes of params.model."""
  eng, kid, mid, ds, raw, mask = _prepare(mean_func, cov_func, params, dataset,
                                          warp_func, exclude_aligned)
  d = ds.d if ds.num_tasks else 1
  if ds.num_tasks == 0:
    zero = torch.zeros(3 + d, dtype=torch.float64)
    return torch.zeros((), device=eng.device, dtype=eng.dtype), \
        params_utils.unpack_like(params.model, zero, d, mid == 1, is_grad=True)
  sums = eng.nll_grad(kid, mid, ds, raw, mask)
  cnt = sums[-1]
  grads = params_utils.unpack_like(params.model, (sums[1:-1] / cnt), d,
                                   mid == 1, is_grad=True)
  return sums[0] / cnt, grads


nll = neg_log_marginal_likelihood


def _unsupported(name):

  def f(*args, **kwargs):
    raise NotImplementedError(
        f"objective '{name}' (EKL / Euclidean regulariser on aligned data, "
        "objectives.py:29-101) is outside the B200 hot path")

  f.__name__ = name
  return f


multivariate_normal_divergence = _unsupported("multivariate_normal_divergence")
multivariate_normal_euc_distance = _unsupported(
    "multivariate_normal_euc_distance")
kl = ekl = regkl = multivariate_normal_divergence
euc = regeuc = multivariate_normal_euc_distance
