"""Mean function library -- mirrors hyperbo/gp_utils/mean.py:30-64.

mean_func(params, vx, warp_func=None) -> (n, 1).  `hb_mean_id` tags the
callables the batched engine understands (zero, constant)."""
from __future__ import annotations

import torch

from hyperbo_b200.basics import params_utils

retrieve_params = params_utils.retrieve_params


def _device_of(vx):
  if isinstance(vx, torch.Tensor):
    return vx.device
  return torch.device("cuda") if torch.cuda.is_available() else torch.device(
      "cpu")


def zero(params, vx, warp_func=None):
  """Zero mean function (mean.py:54-57)."""
  del params, warp_func
  n = torch.as_tensor(vx).shape[0]
  return torch.zeros((n, 1), dtype=torch.float64, device=_device_of(vx))


zero.hb_mean_id = 0


def constant(params, vx, warp_func=None):
  """Constant mean function (mean.py:60-64)."""
  val, = retrieve_params(params, ["constant"], warp_func)
  n = torch.as_tensor(vx).shape[0]
  val = float(torch.as_tensor(val, dtype=torch.float64).reshape(-1)[0])
  return torch.full((n, 1), val, dtype=torch.float64, device=_device_of(vx))


constant.hb_mean_id = 1


def _unsupported(name):

  def f(params, x, warp_func=None):
    raise NotImplementedError(
        f"mean function '{name}' (Flax MLP / dense mean, mean.py:67-79) is "
        "outside the B200 hot path; supported: zero, constant")

  f.__name__ = name
  return f


linear = _unsupported("linear")
linear_mlp = _unsupported("linear_mlp")


def mean_vector(mean_func):
  """Decorator of the reference that maps a per-point mean over the rows of x
  (mean.py:30-51).  The engine's means are built in (zero, constant)."""
  raise NotImplementedError(
      f"custom mean function {getattr(mean_func, '__name__', mean_func)!r}: "
      "the engine's means are built in (zero, constant)")


def mean_id_of(mean_func) -> int:
  mid = getattr(mean_func, "hb_mean_id", None)
  if mid is None:
    raise NotImplementedError(
        f"mean_func {getattr(mean_func, '__name__', mean_func)!r} is not an "
        "engine mean; supported: zero, constant")
  return mid
