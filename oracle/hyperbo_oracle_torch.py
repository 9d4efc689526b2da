"""Second, independent CPU oracle: op-by-op torch restatement with AUTOGRAD.

TEST INFRASTRUCTURE ONLY (see oracle/hyperbo_oracle.py header; PARITY UNPINNED
for the same reason).  Where hyperbo_oracle.py uses the closed-form gradient,
this file follows the reference's *operations* -- including its two custom
VJPs -- and lets reverse-mode autodiff produce the gradient, exactly as
jax.value_and_grad does at gp_utils/gp.py:134.  The two oracles are checked
against each other in tests/test_oracle.py.

It is also the `cpu_baseline` / `--impl reference` arm of bench.py: the
reference (JAX) cannot be installed in this image, so the baseline is this
port run on the host cores with all threads ("kind": "port").
"""
from __future__ import annotations

import math
from typing import Dict

import torch

EPS_WARP = 1e-10
JITTER = 1e-6


def default_softplus(x):  # gp_utils/utils.py:73
  return torch.nn.functional.softplus(x) + EPS_WARP


DEFAULT_WARP_FUNC = {
    "constant": lambda x: x,
    "lengthscale": default_softplus,
    "signal_variance": default_softplus,
    "noise_variance": default_softplus,
}


def retrieve(model, keys, warp_func):  # basics/params_utils.py:97-111
  if warp_func:
    return [warp_func[k](model[k]) if k in warp_func else model[k]
            for k in keys]
  return [model[k] for k in keys]


class _SafeSqrt(torch.autograd.Function):
  """linalg._safe_sqrt, basics/linalg.py:175-190: the cotangent at x == 0 is
  REPLACED by the constant 1e6 (not scaled by the incoming tangent)."""

  @staticmethod
  def forward(ctx, x):
    r = torch.sqrt(x)
    ctx.save_for_backward(x, r)
    return r

  @staticmethod
  def backward(ctx, g):
    x, r = ctx.saved_tensors
    safe = torch.where(x != 0, r, torch.ones_like(r))
    return torch.where(x != 0, g / (2 * safe), torch.full_like(x, 1e6))


class _InvSpdMatVec(torch.autograd.Function):
  """linalg.inverse_spdmatrix_vector_product with its custom VJP,
  basics/linalg.py:139-171 (no gradient into the cached Cholesky factor)."""

  @staticmethod
  def forward(ctx, spd, x, chol):
    out = torch.cholesky_solve(x, chol, upper=False)
    ctx.save_for_backward(chol, x)
    return out

  @staticmethod
  def backward(ctx, g):
    chol, x = ctx.saved_tensors
    inv_x = torch.cholesky_solve(x, chol, upper=False)
    inv_g = torch.cholesky_solve(g, chol, upper=False)
    return -(inv_x @ inv_g.T), inv_g, None


def cov_matrix(name, model, vx1, vx2=None, warp_func=None):
  """kernel.py:33-123 -- vmap(vmap(scalar kernel)) == broadcast (n1,n2,d)."""
  ls, sv = retrieve(model, ["lengthscale", "signal_variance"], warp_func)
  if vx2 is None:
    vx2 = vx1
  diff = (vx1[:, None, :] - vx2[None, :, :]) / ls
  r2 = torch.sum(diff**2, dim=-1)
  if name == "squared_exponential":
    return torch.squeeze(sv) * torch.exp(-r2 / 2)
  if name == "matern32":
    r = math.sqrt(3.0) * _SafeSqrt.apply(r2)
    return torch.squeeze(sv) * (1 + r) * torch.exp(-r)
  if name == "matern52":
    r = math.sqrt(5.0) * _SafeSqrt.apply(r2)
    return sv * (1 + r + r**2 / 3) * torch.exp(-r)
  raise NotImplementedError(name)


def nll_sub_dataset(mean_name, cov_name, model, vx, vy, warp_func):
  """objectives.py:144-156 through linalg.py:36-110."""
  n = vx.shape[0]
  if mean_name == "constant":
    (c,) = retrieve(model, ["constant"], warp_func)
    vy = vy - c
  (nv,) = retrieve(model, ["noise_variance"], warp_func)
  cov = cov_matrix(cov_name, model, vx, None, warp_func) + torch.eye(
      n, dtype=vx.dtype) * (nv + JITTER)
  chol = torch.linalg.cholesky(cov)
  kinvy = _InvSpdMatVec.apply(cov, vy, chol)
  return torch.sum(0.5 * (vy.T @ kinvy) + torch.sum(
      torch.log(torch.diagonal(chol))) + 0.5 * n * math.log(2 * math.pi))


def neg_log_marginal_likelihood(mean_name, cov_name, model, dataset,
                                warp_func):
  """objectives.py:178-195 -- Python task loop, mean over non-empty tasks."""
  total, num = 0.0, 0
  for _, s in dataset.items():
    if len(s) > 2 and s[2] is not None:
      continue
    if s[0].shape[0] == 0:
      continue
    total = total + nll_sub_dataset(mean_name, cov_name, model, s[0], s[1],
                                    warp_func)
    num += 1
  return total / num if num else torch.zeros(())


def neg_log_marginal_likelihood_batched(mean_name, cov_name, model, x, y,
                                        warp_func):
  """Same maths for T equal-size tasks as ONE batched program (x:(T,n,d),
  y:(T,n,1)) -- generous to the CPU baseline (the reference loops tasks)."""
  t, n, _ = x.shape
  ls, sv = retrieve(model, ["lengthscale", "signal_variance"], warp_func)
  (nv,) = retrieve(model, ["noise_variance"], warp_func)
  if mean_name == "constant":
    (c,) = retrieve(model, ["constant"], warp_func)
    y = y - c
  diff = (x[:, :, None, :] - x[:, None, :, :]) / ls
  r2 = torch.sum(diff**2, dim=-1)
  if cov_name == "squared_exponential":
    k = torch.squeeze(sv) * torch.exp(-r2 / 2)
  elif cov_name == "matern32":
    r = math.sqrt(3.0) * _SafeSqrt.apply(r2)
    k = torch.squeeze(sv) * (1 + r) * torch.exp(-r)
  else:
    r = math.sqrt(5.0) * _SafeSqrt.apply(r2)
    k = sv * (1 + r + r**2 / 3) * torch.exp(-r)
  cov = k + torch.eye(n, dtype=x.dtype) * (nv + JITTER)
  chol = torch.linalg.cholesky(cov)
  kinvy = torch.cholesky_solve(y, chol)
  nll = 0.5 * torch.sum(y * kinvy, dim=(1, 2)) + torch.sum(
      torch.log(torch.diagonal(chol, dim1=1, dim2=2)), dim=1) + \
      0.5 * n * math.log(2 * math.pi)
  return torch.mean(nll)


def multivariate_normal_divergence(mean_name, cov_name, model, dataset,
                                   warp_func, eps=0.0, partial=True, euc=False):
  """objectives.py:29-101 with utils.kl_multivariate_normal (utils.py:84-148) or
  euclidean_multivariate_normal (utils.py:151-173), op by op (autograd)."""
  total, num = 0.0, 0
  for _, s in dataset.items():
    if len(s) < 3 or s[2] is None or s[0].shape[0] == 0:
      continue
    x, y = s[0], s[1]
    n = x.shape[0]
    mu0 = torch.mean(y, dim=1)
    yc = y - mu0[:, None]
    cov0 = yc @ yc.T / y.shape[1]  # jnp.cov(y, bias=True)
    if mean_name == "constant":
      (c,) = retrieve(model, ["constant"], warp_func)
      mu1 = c * torch.ones(n, dtype=x.dtype)
    else:
      mu1 = torch.zeros(n, dtype=x.dtype)
    (nv,) = retrieve(model, ["noise_variance"], warp_func)
    cov1 = cov_matrix(cov_name, model, x, None, warp_func) + torch.eye(
        n, dtype=x.dtype) * nv
    if euc:
      val = _SafeSqrt.apply(torch.sum((mu0 - mu1)**2)) + _SafeSqrt.apply(
          torch.sum((cov0 - cov1)**2))
    else:
      if eps > 0:
        cov0 = cov0 + torch.eye(n, dtype=x.dtype) * eps
        cov1 = cov1 + torch.eye(n, dtype=x.dtype) * eps
      if not partial:
        u, sg, _ = torch.linalg.svd(cov0.detach())
        tol = sg.max() * torch.finfo(sg.dtype).eps / 2.0 * math.sqrt(2 * n + 1.0)
        rank = int((sg > tol).sum())
        chol0 = (u * torch.sqrt(sg)[None, :])[:, :rank]
        chol0inv = torch.linalg.pinv(chol0)
        mu1 = chol0inv @ (mu1 - mu0)
        cov1 = chol0inv @ cov1 @ chol0inv.T
        mu0 = torch.zeros_like(mu1)
        cov0 = torch.eye(rank, dtype=x.dtype)
      mu_diff = mu1 - mu0
      chol1 = torch.linalg.cholesky(cov1)
      val = torch.trace(torch.cholesky_solve(cov0, chol1)) + mu_diff @ \
          torch.cholesky_solve(mu_diff[:, None], chol1)[:, 0] + \
          torch.sum(2 * torch.log(torch.diagonal(chol1)))
      if not partial:
        val = 0.5 * (val - rank)
    total = total + val
    num += 1
  return total / num if num else torch.zeros(())


def divergence_value_and_grad(mean_name, cov_name, model_np: Dict, dataset_np,
                              warp_func=DEFAULT_WARP_FUNC, **kw):
  model = to_torch_model(model_np, torch.float64)
  ds = {k: (torch.as_tensor(s[0], dtype=torch.float64), torch.as_tensor(
      s[1], dtype=torch.float64)) + tuple(s[2:]) for k, s in dataset_np.items()}
  loss = multivariate_normal_divergence(mean_name, cov_name, model, ds,
                                        warp_func, **kw)
  loss.backward()
  grads = {k: (v.grad.numpy().copy() if v.grad is not None else
               torch.zeros_like(v).numpy()) for k, v in model.items()}
  return float(loss.detach()), grads


def to_torch_model(model: Dict, dtype=torch.float64, requires_grad=True):
  return {k: torch.tensor(v, dtype=dtype, requires_grad=requires_grad)
          for k, v in model.items()}


def value_and_grad(mean_name, cov_name, model_np: Dict, dataset_np,
                   warp_func=DEFAULT_WARP_FUNC, dtype=torch.float64):
  """jax.value_and_grad(loss_func)(model_param, batch), gp.py:134."""
  model = to_torch_model(model_np, dtype)
  ds = {k: (torch.as_tensor(s[0], dtype=dtype), torch.as_tensor(
      s[1], dtype=dtype)) + tuple(s[2:]) for k, s in dataset_np.items()}
  loss = neg_log_marginal_likelihood(mean_name, cov_name, model, ds, warp_func)
  loss.backward()
  grads = {k: (v.grad.numpy().copy() if v.grad is not None else
               torch.zeros_like(v).numpy()) for k, v in model.items()}
  return float(loss.detach()), grads


class AdamTrainer:
  """gp.infer_parameters Adam branch (gp.py:114-157) as a timed CPU loop.
  torch.optim.Adam with eps=1e-8, betas=(.9,.999) == optax.adam defaults."""

  def __init__(self, mean_name, cov_name, model_np, x, y, lr=1e-3,
               dtype=torch.float64, batched=True,
               warp_func=DEFAULT_WARP_FUNC):
    self.mean_name, self.cov_name = mean_name, cov_name
    self.model = to_torch_model(model_np, dtype)
    self.x = torch.as_tensor(x, dtype=dtype)
    self.y = torch.as_tensor(y, dtype=dtype)
    self.batched, self.warp_func = batched, warp_func
    self.opt = torch.optim.Adam(list(self.model.values()), lr=lr, eps=1e-8)

  def step(self) -> float:
    self.opt.zero_grad(set_to_none=True)
    if self.batched:
      loss = neg_log_marginal_likelihood_batched(
          self.mean_name, self.cov_name, self.model, self.x, self.y,
          self.warp_func)
    else:
      ds = {t: (self.x[t], self.y[t]) for t in range(self.x.shape[0])}
      loss = neg_log_marginal_likelihood(self.mean_name, self.cov_name,
                                         self.model, ds, self.warp_func)
    loss.backward()
    val = float(loss.detach())  # the isfinite host read of gp.py:135-142
    if not math.isfinite(val):
      raise FloatingPointError("non-finite loss")
    self.opt.step()
    return val
