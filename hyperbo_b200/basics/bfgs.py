"""BFGS driver -- role of hyperbo/basics/bfgs.py:24-53 (which wraps
jax.scipy.optimize.minimize(method='BFGS')): here scipy's BFGS on the host with
the engine's value-and-gradient as the objective."""
from __future__ import annotations

import numpy as np
import scipy.optimize


def bfgs(val_and_grad_fn, x0, tol=1e-8, max_training_step=100):
  """Returns (x, final value)."""
  res = scipy.optimize.minimize(
      lambda v: val_and_grad_fn(np.asarray(v, dtype=np.float64)), np.asarray(
          x0, dtype=np.float64), jac=True, method="BFGS", tol=tol,
      options={"maxiter": int(max_training_step)})
  return res.x, float(res.fun)
