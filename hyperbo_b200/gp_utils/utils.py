"""Warp functions -- mirrors hyperbo/gp_utils/utils.py:27-81.

The engine differentiates the objective in closed form on the GPU, so it has to
*recognise* a warp rather than trace it: the callables below carry an `hb_warp`
tag ('identity' or 'softplus_eps').  They are also ordinary torch callables.
"""
from __future__ import annotations

import torch

EPS = 1e-10  # utils.py:28


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
