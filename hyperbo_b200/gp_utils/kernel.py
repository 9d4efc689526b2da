"""Kernel library -- mirrors hyperbo/gp_utils/kernel.py.

Same call signature as the reference's `covariance_matrix.matrix_map`
(kernel.py:33-58):  cov_func(params, vx1, vx2=None, warp_func=None, diag=False)
-> (n1, n2) torch CUDA tensor (or (n1,) for diag=True with vx2=None).
Each callable carries `hb_kernel_id`, which is how the batched engine entry
points (objectives, gp, acfun) recognise it -- they never call it per pair.
"""
from __future__ import annotations

import torch

from hyperbo_b200 import engine as _engine
from hyperbo_b200.basics import params_utils

retrieve_params = params_utils.retrieve_params


def _as_2d(v):
  if not isinstance(v, torch.Tensor):
    import numpy as np
    v = np.asarray(v, dtype=np.float64)
  t = torch.as_tensor(v)
  if t.dim() == 1:
    t = t[None, :]
  return t


def _make(name: str, kernel_id: int, doc: str):

  def matrix_map(params, vx1, vx2=None, warp_func=None, diag=False):
    vx1 = _as_2d(vx1)
    d = vx1.shape[1]
    raw, mask, _ = params_utils.pack_raw(params.model, d, need_mean=False,
                                         warp_func=warp_func, need_noise=False)
    eng = _engine.Engine.get()
    if vx2 is not None:
      vx2 = _as_2d(vx2)
    # kernel.py:54-58: diag is honoured only when vx2 is None
    return eng.kernel_matrix(kernel_id, vx1, vx2, raw, mask,
                             diag=bool(diag and vx2 is None))

  matrix_map.__name__ = name
  matrix_map.__qualname__ = name
  matrix_map.__doc__ = doc
  matrix_map.hb_kernel_id = kernel_id
  matrix_map.hb_kernel_name = name
  return matrix_map


squared_exponential = _make(
    "squared_exponential", 0,
    "Squared exponential kernel, Eq.(4.9/13) of GPML (kernel.py:63-81).")
matern32 = _make("matern32", 1,
                 "Matern 3/2 kernel, Eq.(4.17) of GPML (kernel.py:84-102).")
matern52 = _make("matern52", 2,
                 "Matern 5/2 kernel, Eq.(4.17) of GPML (kernel.py:105-123).")


def _unsupported(name: str, why: str):

  def matrix_map(params, vx1, vx2=None, warp_func=None, diag=False):
    raise NotImplementedError(
        f"kernel '{name}' is outside the B200 hot path ({why}); "
        "supported: squared_exponential, matern32, matern52")

  matrix_map.__name__ = name
  matrix_map.__qualname__ = name
  return matrix_map


# kernel.py:126-222 -- registered so that const.KERNEL keeps the reference's
# names, but they raise: no CPU fallback by design.
dot_product = _unsupported("dot_product", "not a stationary ARD kernel")
dot_product_mlp = _unsupported("dot_product_mlp", "learned MLP input warp")
squared_exponential_mlp = _unsupported("squared_exponential_mlp",
                                       "learned MLP input warp")
matern32_mlp = _unsupported("matern32_mlp", "learned MLP input warp")
matern52_mlp = _unsupported("matern52_mlp", "learned MLP input warp")
dot_product_kumar = _unsupported("dot_product_kumar", "Kumaraswamy warp")
squared_exponential_kumar = _unsupported("squared_exponential_kumar",
                                         "Kumaraswamy warp")
matern32_kumar = _unsupported("matern32_kumar", "Kumaraswamy warp")
matern52_kumar = _unsupported("matern52_kumar", "Kumaraswamy warp")


def covariance_matrix(kernel):
  """Decorator of the reference that turns a pairwise kernel into a matrix map
  (kernel.py:29-60).  The engine evaluates its kernels tile by tile on the GPU
  and differentiates them in closed form, so it cannot adopt an arbitrary Python
  pairwise function: use squared_exponential / matern32 / matern52."""
  raise NotImplementedError(
      f"custom pairwise kernel {getattr(kernel, '__name__', kernel)!r}: the "
      "engine's kernels are built in (squared_exponential, matern32, matern52)")


def with_mlp_bases(kernel):  # kernel.py:148-176
  raise NotImplementedError("learned MLP input warps are outside the hot path")


def with_kumar_bases(kernel):  # kernel.py:179-209
  raise NotImplementedError("Kumaraswamy input warps are outside the hot path")


def kernel_id_of(cov_func) -> int:
  kid = getattr(cov_func, "hb_kernel_id", None)
  if kid is None:
    raise NotImplementedError(
        f"cov_func {getattr(cov_func, '__name__', cov_func)!r} is not an engine "
        "kernel; supported: squared_exponential, matern32, matern52")
  return kid
