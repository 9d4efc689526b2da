"""L-BFGS driver -- same algorithm as hyperbo/basics/lbfgs.py (two-directional
backtracking line search with the Armijo and Wolfe-curvature tests :51-139,
Nocedal two-loop direction :141-183, driver :186-349), restated on FLAT numpy
vectors: the objective is a black box `val_and_grad_fn(x) -> (float, ndarray)`,
which here is one batched engine call (hb_nll_grad_batched) plus a read-back of
P+2 scalars.  All of this is host logic; nothing runs on the device but the
objective.
"""
from __future__ import annotations

import logging
import math
from typing import Callable, List, Optional, Tuple

import numpy as np

ValGrad = Callable[[np.ndarray], Tuple[float, np.ndarray]]
MultiValGrad = Callable[[List[np.ndarray]], List[Tuple[float, np.ndarray]]]


class _Evaluator:
  """The objective behind a small memo of recent points.

  * The driver re-evaluates the point its line search just accepted
    (lbfgs.py:301 after :136): same bytes, same deterministic engine call, so the
    memo returns the stored pair instead of a second factorisation pass.
  * With `multi_fn` (S points in ONE engine call, hb_nll_grad_multi) the line
    search evaluates a trial step together with its two possible successors
    (alpha * tau, alpha * 2.1) from the second trial on: the next trial is then
    already known and a long search needs half the device round trips.  The sequence of trial
    steps and every decision are those of the plain search.
  """

  def __init__(self, fn: ValGrad, multi_fn: Optional[MultiValGrad] = None,
               keep: int = 16):
    self.fn, self.multi_fn, self.keep = fn, multi_fn, keep
    self.memo = {}
    self.calls = 0        # engine round trips
    self.points = 0       # points evaluated (speculative ones included)

  def _store(self, x, res):
    self.memo[x.tobytes()] = res
    while len(self.memo) > self.keep:
      self.memo.pop(next(iter(self.memo)))

  def __call__(self, x: np.ndarray):
    hit = self.memo.get(x.tobytes())
    if hit is not None:
      return hit
    res = self.fn(x)
    self.calls += 1
    self.points += 1
    self._store(x, res)
    return res

  def prefetch(self, xs: List[np.ndarray]):
    """xs[0] is the next trial, the rest its possible successors: if the trial is
    not known yet, evaluate all unknown points in one call."""
    if self.multi_fn is None or xs[0].tobytes() in self.memo:
      return
    todo = [x for x in xs if x.tobytes() not in self.memo]
    if len(todo) < 2:
      return
    out = self.multi_fn(todo)
    self.calls += 1
    self.points += len(todo)
    for x, res in zip(todo, out):
      self._store(x, res)


def backtracking_linesearch(val_and_grad_fn: ValGrad, cur_val: float,
                            x: np.ndarray, grads: np.ndarray,
                            direction: np.ndarray, alpha: float = 1.0,
                            c1: float = 1e-4, c2: float = 0.9, tau: float = 0.5,
                            max_steps: int = 50) -> Tuple[float, float]:
  """Returns (new value, step size); (cur_val, 0.) when every trial was
  non-finite (lbfgs.py:136-139)."""
  slope = float(np.dot(grads, direction))
  if slope > 0.0:
    # not a descent direction: the caller sees "no progress" (lbfgs.py:92-95)
    logging.info("Incorrect descent direction %f. Exiting linesearch", slope)
    return cur_val, 0.0
  new_val = cur_val
  prefetch = getattr(val_and_grad_fn, "prefetch", None)
  for trial in range(max_steps):
    # (the first trial is accepted most of the time: speculate only once it was
    # not) this trial + both possible next trials in one call
    if prefetch is not None and trial > 0:
      prefetch([x + a * direction for a in (alpha, alpha * tau, alpha * 2.1)])
    new_val, new_grads = val_and_grad_fn(x + alpha * direction)
    armijo = math.isfinite(new_val) and cur_val + alpha * c1 * slope >= new_val
    if armijo:
      if float(np.dot(new_grads, direction)) >= c2 * slope:
        return new_val, alpha
      alpha *= 2.1  # sufficient decrease but still steep: lengthen
    else:
      alpha *= tau
  if math.isfinite(new_val):
    return new_val, alpha
  return cur_val, 0.0


def descent_direction(grads: np.ndarray, s: List[np.ndarray],
                      y: List[np.ndarray]) -> np.ndarray:
  """Nocedal's two-loop recursion (lbfgs.py:141-183)."""
  q = -grads
  rho = [1.0 / float(np.dot(yi, si)) for si, yi in zip(s, y)]
  a = [0.0] * len(s)
  for i in range(len(s) - 1, -1, -1):
    a[i] = rho[i] * float(np.dot(s[i], q))
    q = q - a[i] * y[i]
  gamma = float(np.dot(s[-1], y[-1])) / float(np.dot(y[-1], y[-1]))
  r = gamma * q
  for i in range(len(s)):
    b = rho[i] * float(np.dot(y[i], r))
    r = r + (a[i] - b) * s[i]
  return r


def lbfgs(val_and_grad_fn: ValGrad, x0: np.ndarray, memory: int = 10,
          ls_steps: int = 50, steps: int = 100, alpha: float = 1.0,
          tol: float = 1e-6, ls_tau: float = 0.5, state=None,
          callback: Optional[Callable] = None,
          multi_fn: Optional[MultiValGrad] = None, stats: Optional[dict] = None):
  """Minimise with L-BFGS.  Returns (value, x, state) like lbfgs.py:186-349;
  `state = (s, y, old_grads, old_x)` resumes the Hessian estimate.  `multi_fn`
  (several points per engine call) enables the speculative line search of
  _Evaluator; `stats` receives the number of engine calls / points."""
  val_and_grad_fn = _Evaluator(val_and_grad_fn, multi_fn)
  try:
    return _lbfgs(val_and_grad_fn, x0, memory, ls_steps, steps, alpha, tol, ls_tau,
                  state, callback)
  finally:
    if stats is not None:
      stats["calls"] = val_and_grad_fn.calls
      stats["points"] = val_and_grad_fn.points


def _lbfgs(val_and_grad_fn, x0, memory, ls_steps, steps, alpha, tol, ls_tau, state,
           callback):
  x = np.array(x0, dtype=np.float64)
  if state is None:
    s_k: List[np.ndarray] = []
    y_k: List[np.ndarray] = []
    val, grads = val_and_grad_fn(x)
    if callback is not None:
      callback(step=0, model_params=x, loss=val)
    gnorm = float(np.dot(grads, grads))
    if gnorm <= tol:
      return val, x, None
    old_x, old_g = x.copy(), grads.copy()
    new_val, step = backtracking_linesearch(
        val_and_grad_fn, val, x, grads, -grads, 1.0 / math.sqrt(gnorm),
        tau=ls_tau, max_steps=ls_steps)
    if new_val < val:
      x = x - step * grads
    else:
      return new_val, x, (s_k, y_k, old_g, old_x)
  else:
    s_k, y_k, old_g, old_x = state
    s_k, y_k = list(s_k), list(y_k)
  new_val = None
  for i in range(1, steps + 1):
    val, grads = val_and_grad_fn(x)
    if float(np.dot(grads, grads)) <= tol:
      new_val = val
      break
    if old_g is not None:
      y_k.append(grads - old_g)
      s_k.append(x - old_x)
    if len(s_k) > memory:
      s_k, y_k = s_k[-memory:], y_k[-memory:]
    old_x, old_g = x.copy(), grads.copy()
    curv = float(np.dot(y_k[-1], s_k[-1]))
    if callback is not None:
      callback(step=i, model_params=x, loss=val)
    if math.isfinite(curv) and curv >= tol:
      d = descent_direction(grads, s_k, y_k)
      new_val, step = backtracking_linesearch(
          val_and_grad_fn, val, x, grads, d, alpha, tau=ls_tau,
          max_steps=ls_steps)
      if new_val >= val:
        break  # the line search made no progress
      x = x + step * d
    else:
      new_val = val  # unstable curvature estimate: stop (lbfgs.py:339-342)
      break
  return new_val, x, (s_k, y_k, old_g, old_x)
