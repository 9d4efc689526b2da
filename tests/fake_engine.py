"""A CPU stand-in for hyperbo_b200.engine.Engine -- TEST INFRASTRUCTURE ONLY.

It answers `nll_grad` (the hb_nll_grad_batched / hb_nll_grad_weighted contract:
weighted sums of per-task NLL values and raw-parameter gradients, runtime
jitter) with the oracle's arithmetic on CPU tensors, so that the HOST logic
built on top of the engine -- objective programs, task sharding, the
all-reduce contract -- runs under `-m "not gpu"` and under gloo.  The product
never imports this module; on a GPU box the same host logic drives the CUDA
kernels (tests/test_gpu_objectives.py)."""
import math

import numpy as np
import scipy.linalg as spla
import torch

from hyperbo_b200 import engine as _engine
from oracle import hyperbo_oracle as O

_KERNELS = {v: k for k, v in _engine.KERNEL_IDS.items()}


def task_nll_and_grad(kid, mid, x, y, raw, mask, jitter):
  """nll and d nll / d raw (length 3 + d) of one task, closed form."""
  n, d = x.shape
  raw = np.asarray(raw, dtype=np.float64)
  warped = np.array([(mask >> p) & 1 for p in range(3 + d)], dtype=bool)
  theta = np.where(warped, O.softplus(raw) + O.EPS_WARP, raw)
  chain = np.where(warped, O.sigmoid(raw), 1.0)
  c = theta[0] if mid == 1 else 0.0
  sv, nv, ls = theta[1], theta[2], theta[3:]
  name = _KERNELS[kid]
  r2, diff = O._scaled_sqdist(x, x, ls)  # pylint: disable=protected-access
  k = O._kernel_from_r2(name, r2, sv)  # pylint: disable=protected-access
  try:
    chol = np.linalg.cholesky(k + np.eye(n) * (nv + jitter))
  except np.linalg.LinAlgError:  # engine contract: NaN propagates, no raise
    return float("nan"), np.full(3 + d, np.nan)
  r = (y - c)[:, None]
  alpha = spla.cho_solve((chol, True), r)
  nll = float(0.5 * (r.T @ alpha).item() + np.sum(np.log(np.diag(chol))) +
              0.5 * n * math.log(2 * math.pi))
  g = 0.5 * (spla.cho_solve((chol, True), np.eye(n)) - alpha @ alpha.T)
  w = O._pair_weight(name, r2, k, sv)  # pylint: disable=protected-access
  grad = np.zeros(3 + d)
  grad[0] = -float(np.sum(alpha)) if mid == 1 else 0.0
  grad[1] = float(np.sum(g * k) / sv)
  grad[2] = float(np.trace(g))
  grad[3:] = np.einsum("ij,ijk->k", g * w, diff * diff) / ls
  return nll, grad * chain


class FakeEngine(_engine.Engine):
  """Engine whose arithmetic is the oracle's (CPU, fp64)."""

  def __init__(self):  # pylint: disable=super-init-not-called
    self.device = torch.device("cpu")
    self.dtype = torch.float64
    self.max_dim = 32
    self.calls = 0

  def nll_grad(self, kernel_id, mean_id, ds, raw, mask, sums_out=None,
               want_task_nll=False, weights=None, jitter=None):
    raw = np.asarray(torch.as_tensor(raw).detach().cpu(), dtype=np.float64)
    jitter = O.JITTER if jitter is None else float(jitter)
    T, P = ds.num_tasks, 3 + ds.d
    w = np.ones(T) if weights is None else np.asarray(
        torch.as_tensor(weights).cpu(), dtype=np.float64)
    out = np.zeros(P + 2)
    per_task = np.zeros(T)
    x, y = ds.x.numpy(), ds.y.numpy()
    for t in range(T):
      lo, hi = ds.offs[t], ds.offs[t + 1]
      v, g = task_nll_and_grad(kernel_id, mean_id, x[lo:hi], y[lo:hi], raw, mask,
                               jitter)
      out[0] += w[t] * v
      out[1:-1] += w[t] * g
      out[-1] += 1.0
      per_task[t] = v
    self.calls += 1
    res = torch.from_numpy(out)
    if sums_out is not None:
      sums_out.copy_(res)
      res = sums_out
    if want_task_nll:
      return res, torch.from_numpy(per_task)
    return res


  def nll_grad_mrhs(self, kernel_id, mean_id, ds, R, B, col_weight, col_mean, raw,
                    mask, weights=None, jitter=None, sums_out=None):
    """hb_nll_grad_mrhs contract, written out directly (dense inverse): one K~ per
    task, R residual columns, value and raw-parameter gradient."""
    raw = self._np(raw)
    jitter = O.JITTER if jitter is None else float(jitter)
    T, d = ds.num_tasks, ds.d
    P = 3 + d
    warped = np.array([(mask >> p) & 1 for p in range(P)], dtype=bool)
    theta = np.where(warped, O.softplus(raw) + O.EPS_WARP, raw)
    chain = np.where(warped, O.sigmoid(raw), 1.0)
    c = theta[0] if mean_id == 1 else 0.0
    sv, nv, ls = theta[1], theta[2], theta[3:]
    name = _KERNELS[kernel_id]
    B = self._np(B).reshape(-1)
    cw = self._np(col_weight).reshape(T, R)
    cm = np.zeros(R) if col_mean is None else self._np(col_mean).reshape(R)
    w = np.ones(T) if weights is None else self._np(weights).reshape(T)
    out = np.zeros(P + 2)
    x = ds.x.numpy()
    for t in range(T):
      lo, hi = ds.offs[t], ds.offs[t + 1]
      n = hi - lo
      if n == 0:
        continue
      xt = x[lo:hi]
      res = B[lo * R:hi * R].reshape(R, n).T - c * cm[None, :]   # (n, R)
      r2, diff = O._scaled_sqdist(xt, xt, ls)  # pylint: disable=protected-access
      k = O._kernel_from_r2(name, r2, sv)  # pylint: disable=protected-access
      chol = np.linalg.cholesky(k + np.eye(n) * (nv + jitter))
      kinv = spla.cho_solve((chol, True), np.eye(n))
      a = kinv @ res
      out[0] += w[t] * (np.sum(np.log(np.diag(chol))) + 0.5 * n * math.log(2 * math.pi))
      out[0] += 0.5 * np.sum(cw[t] * np.sum(res * a, axis=0))
      g = 0.5 * (w[t] * kinv - (a * cw[t][None, :]) @ a.T)
      pw = O._pair_weight(name, r2, k, sv)  # pylint: disable=protected-access
      grad = np.zeros(P)
      grad[0] = -float(np.sum(cw[t] * cm * np.sum(a, axis=0))) if mean_id == 1 else 0.0
      grad[1] = float(np.sum(g * k) / sv)
      grad[2] = float(np.trace(g))
      grad[3:] = np.einsum("ij,ijk->k", g * pw, diff * diff) / ls
      out[1:-1] += grad * chain
      out[-1] += 1.0
    self.calls += 1
    res_t = torch.from_numpy(out)
    if sums_out is not None:
      sums_out.copy_(res_t)
      res_t = sums_out
    return res_t

  def euclid_grad(self, kernel_id, mean_id, ds, R, Yc, mu0, raw, mask,
                  mean_weight=1.0, cov_weight=1.0, weights=None, sums_out=None):
    """hb_euclid_grad contract, dense closed form (d ||E||_F = <E, dE> / ||E||_F)."""
    raw = self._np(raw)
    T, d = ds.num_tasks, ds.d
    P = 3 + d
    warped = np.array([(mask >> p) & 1 for p in range(P)], dtype=bool)
    theta = np.where(warped, O.softplus(raw) + O.EPS_WARP, raw)
    chain = np.where(warped, O.sigmoid(raw), 1.0)
    c = theta[0] if mean_id == 1 else 0.0
    sv, nv, ls = theta[1], theta[2], theta[3:]
    name = _KERNELS[kernel_id]
    Yc, mu0 = self._np(Yc).reshape(-1), self._np(mu0).reshape(-1)
    w = np.ones(T) if weights is None else self._np(weights).reshape(T)
    out = np.zeros(P + 2)
    x = ds.x.numpy()
    for t in range(T):
      lo, hi = ds.offs[t], ds.offs[t + 1]
      n = hi - lo
      if n == 0:
        continue
      xt = x[lo:hi]
      yc = Yc[lo * R:hi * R].reshape(R, n).T
      r2, diff = O._scaled_sqdist(xt, xt, ls)  # pylint: disable=protected-access
      k = O._kernel_from_r2(name, r2, sv)  # pylint: disable=protected-access
      e = yc @ yc.T - k - nv * np.eye(n)
      f = math.sqrt(float(np.sum(e * e)))
      dm = mu0[lo:hi] - c
      nm = math.sqrt(float(np.sum(dm * dm)))
      out[0] += w[t] * (mean_weight * nm + cov_weight * f)
      pw = O._pair_weight(name, r2, k, sv)  # pylint: disable=protected-access
      grad = np.zeros(P)
      if mean_id == 1 and nm > 0:
        grad[0] = -mean_weight * float(np.sum(dm)) / nm
      if f > 0:
        grad[1] = -cov_weight * float(np.sum(e * k)) / sv / f
        grad[2] = -cov_weight * float(np.trace(e)) / f
        grad[3:] = -cov_weight * np.einsum("ij,ijk->k", e * pw, diff * diff) / ls / f
      out[1:-1] += w[t] * grad * chain
      out[-1] += 1.0
    self.calls += 1
    res_t = torch.from_numpy(out)
    if sums_out is not None:
      sums_out.copy_(res_t)
      res_t = sums_out
    return res_t

  # ---- the rest of the Engine surface (same contracts as engine.py) ---------
  @staticmethod
  def _np(a):
    return np.asarray(torch.as_tensor(a).detach().cpu(), dtype=np.float64)

  @staticmethod
  def _theta(raw, mask, d):
    raw = FakeEngine._np(raw)
    warped = np.array([(mask >> p) & 1 for p in range(3 + d)], dtype=bool)
    return np.where(warped, O.softplus(raw) + O.EPS_WARP, raw)

  def _gram(self, kid, x1, x2, theta):
    r2, _ = O._scaled_sqdist(x1, x2, theta[3:])  # pylint: disable=protected-access
    return O._kernel_from_r2(_KERNELS[kid], r2, theta[1])  # pylint: disable=protected-access

  def kernel_matrix(self, kernel_id, x1, x2, raw, mask, diag=False,
                    add_noise=False, jitter=O.JITTER):
    x1 = self._np(x1)
    theta = self._theta(raw, mask, x1.shape[1])
    if x2 is None:
      if diag:
        return torch.full((x1.shape[0],), float(theta[1]), dtype=torch.float64)
      k = self._gram(kernel_id, x1, x1, theta)
      if add_noise:
        k = k + np.eye(x1.shape[0]) * (theta[2] + jitter)
      return torch.from_numpy(k)
    return torch.from_numpy(self._gram(kernel_id, x1, self._np(x2), theta))

  def _solve(self, kid, mid, x, y, raw, mask):
    theta = self._theta(raw, mask, x.shape[1])
    n = x.shape[0]
    cov = self._gram(kid, x, x, theta) + np.eye(n) * (theta[2] + O.JITTER)
    r = (y.reshape(-1) - (theta[0] if mid == 1 else 0.0))[:, None]
    try:
      chol = np.linalg.cholesky(cov)
    except np.linalg.LinAlgError:
      return np.full((n, n), np.nan), np.full((n, 1), np.nan), np.nan, 1
    alpha = spla.cho_solve((chol, True), r)
    nll = float(0.5 * (r.T @ alpha).item() + np.sum(np.log(np.diag(chol))) +
                0.5 * n * math.log(2 * math.pi))
    return chol, alpha, nll, 0

  def factorize(self, kernel_id, mean_id, ds, raw, mask, want_chol=True,
                want_alpha=True):
    x, y = ds.x.numpy(), ds.y.numpy()
    chols, alphas, nlls, infos = [], [], [], []
    for t in range(ds.num_tasks):
      lo, hi = ds.offs[t], ds.offs[t + 1]
      c, a, v, info = self._solve(kernel_id, mean_id, x[lo:hi], y[lo:hi], raw, mask)
      chols.append(torch.from_numpy(c))
      alphas.append(a.reshape(-1))
      nlls.append(v)
      infos.append(info)
    alpha = torch.from_numpy(np.concatenate(alphas)) if alphas else torch.zeros(0)
    return (chols if want_chol else None, alpha if want_alpha else None,
            torch.tensor(nlls, dtype=torch.float64),
            torch.tensor(infos, dtype=torch.int32))

  def build_predictor(self, kernel_id, mean_id, x, y, raw, mask):
    x, y = self._np(x), self._np(y)
    chol, alpha, nll, info = self._solve(kernel_id, mean_id, x, y, raw, mask)
    cache = (chol, alpha)  # opaque to the callers, like the packed device cache
    return (cache, torch.from_numpy(chol), torch.from_numpy(alpha),
            torch.tensor([nll]), torch.tensor([info], dtype=torch.int32))

  def _acq(self, acq_id, param, mu, var):
    std = np.sqrt(var)
    if acq_id == 1:
      return O.expected_improvement_sub(mu, std, param)
    if acq_id == 2:
      return O.probability_of_improvement_sub(mu, std, param)
    return O.ucb_sub(mu, std, param)

  def predict(self, kernel_id, mean_id, x, cache, raw, mask, xq, noise_flag=0.0,
              var_scale=1.0, acq_id=0, acq_param=0.0, want_mu=True,
              want_var=True):
    xq = self._np(xq)
    theta = self._theta(raw, mask, xq.shape[1])
    mu = np.full((xq.shape[0], 1), theta[0] if mean_id == 1 else 0.0)
    v2 = np.zeros((xq.shape[0], 1))
    if x is not None and torch.as_tensor(x).shape[0] > 0:
      chol, alpha = cache
      ks = self._gram(kernel_id, self._np(x), xq, theta)
      mu = mu + ks.T @ alpha
      v = spla.solve_triangular(chol, ks, lower=True)
      v2 = np.sum(v * v, axis=0)[:, None]
    var = (theta[1] - v2 + noise_flag * theta[2]) * var_scale
    acq = torch.from_numpy(self._acq(acq_id, acq_param, mu, var)) if acq_id \
        else None
    return (torch.from_numpy(mu) if want_mu else None,
            torch.from_numpy(var) if want_var else None, acq)

  def predict_cov(self, kernel_id, mean_id, x, cache, raw, mask, xq, noise_flag=0.0,
                  var_scale=1.0):
    xq = self._np(xq)
    theta = self._theta(raw, mask, xq.shape[1])
    mu = np.full((xq.shape[0], 1), theta[0] if mean_id == 1 else 0.0)
    cov = self._gram(kernel_id, xq, xq, theta)
    if x is not None and torch.as_tensor(x).shape[0] > 0:
      chol, alpha = cache
      ks = self._gram(kernel_id, self._np(x), xq, theta)
      mu = mu + ks.T @ alpha
      v = spla.solve_triangular(chol, ks, lower=True)
      cov = cov - v.T @ v
    cov = (cov + noise_flag * theta[2] * np.eye(xq.shape[0])) * var_scale
    return torch.from_numpy(mu), torch.from_numpy(cov)

  def acquisition(self, acq_id, param, mu, var):
    mu, var = self._np(mu).reshape(-1, 1), self._np(var).reshape(-1, 1)
    return torch.from_numpy(self._acq(acq_id, float(param), mu, var))

  def adam_step(self, P, raw, m, v, accepted, sums, scal, lr, b1=0.9, b2=0.999,
                eps=1e-8, tie_lengthscale=False):
    """hb_adam_step (k_adam): optax.adam + the accept / stop rule of
    gp.py:135-146, in place on CPU tensors."""
    cnt = float(sums[1 + P])
    loss = float(sums[0]) / cnt if cnt > 0 else 0.0
    stopped = bool(scal[2] != 0)
    scal[0] = loss
    if stopped:
      return
    if not math.isfinite(loss):
      scal[2] = 1.0
      return
    t = float(scal[1]) + 1.0
    scal[1] = t
    scal[3] += 1.0
    g = sums[1:1 + P].clone() / cnt if cnt > 0 else torch.zeros(P, dtype=raw.dtype)
    if tie_lengthscale:
      g[3:] = g[3:].sum()
    accepted.copy_(raw)
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    mhat = m / (1 - b1**t)
    vhat = v / (1 - b2**t)
    raw.sub_(lr * mhat / (torch.sqrt(vhat) + eps))


def install(monkeypatch):
  """Route Engine.get() to one FakeEngine for the duration of a test."""
  eng = FakeEngine()
  monkeypatch.setattr(_engine.Engine, "get", staticmethod(lambda *a, **k: eng))
  return eng
