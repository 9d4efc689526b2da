"""CPU fp64 oracle for the HyperBO GP hot path -- TEST INFRASTRUCTURE ONLY.

This file restates, in plain NumPy/SciPy float64, the arithmetic of the
reference path  kernel matrix -> Cholesky(K + s^2 I) -> solves -> NLL (+ its
hyper-parameter gradient) -> predict -> acquisition  of google-research/hyperbo
(reference @ e720fc1).  Each function cites the reference file:line it follows
(paths relative to /root/reference/hyperbo/).

PARITY UNPINNED: the reference ships no golden vectors / known-answer tests for
this path (its tests assert shapes, inequalities and identities only) and JAX
is not installable in this image, so the reference itself cannot be executed
here.  The oracle is therefore pinned against (a) the reference's own test
identities, (b) closed-form known answers, (c) central finite differences and
(d) an independent op-by-op torch-autograd restatement
(oracle/hyperbo_oracle_torch.py).  See tests/test_oracle.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
leg may import this module.  The product (hyperbo_b200) never does.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import scipy.linalg as spla
from scipy.special import erf

EPS_WARP = 1e-10  # gp_utils/utils.py:28   (EPS)
JITTER = 1e-6  # basics/linalg.py:42     (eps default)

KERNELS = ("squared_exponential", "matern32", "matern52")
MEANS = ("constant", "zero")


# ----------------------------------------------------------------------------
# warps / parameter retrieval
# ----------------------------------------------------------------------------
def softplus(x):
  """jax.nn.softplus == logaddexp(x, 0)  (gp_utils/utils.py:29)."""
  return np.logaddexp(np.asarray(x, dtype=np.float64), 0.0)


def sigmoid(x):
  x = np.asarray(x, dtype=np.float64)
  return 0.5 * (1.0 + np.tanh(0.5 * x))


def default_softplus(x):
  """DEFAULT_SOFTPLUS = softplus(x) + EPS  (gp_utils/utils.py:73)."""
  return softplus(x) + EPS_WARP


identity_warp = lambda x: x  # gp_utils/utils.py:27

# gp_utils/utils.py:75-81
DEFAULT_WARP_FUNC = {
    "constant": identity_warp,
    "lengthscale": default_softplus,
    "signal_variance": default_softplus,
    "noise_variance": default_softplus,
    "dot_prod_sigma": default_softplus,
}


def retrieve_params(model: Dict, keys: Sequence[str], warp_func=None) -> List:
  """basics/params_utils.py:90-111 (ValueError on missing key; per-key warp)."""
  if not set(keys).issubset(set(model.keys())):
    raise ValueError(f"Expected parameters are {sorted(keys)}, "
                     f"but received {sorted(model.keys())}.")
  if warp_func:
    return [
        warp_func[k](model[k]) if k in warp_func else model[k] for k in keys
    ]
  return [model[k] for k in keys]


# ----------------------------------------------------------------------------
# kernels  (gp_utils/kernel.py:29-123)
# ----------------------------------------------------------------------------
def _scaled_sqdist(vx1, vx2, lengthscale):
  """sum(((x1-x2)/l)**2) for every pair -- DIRECT differences, kernel.py:80."""
  vx1 = np.asarray(vx1, dtype=np.float64)
  vx2 = np.asarray(vx2, dtype=np.float64)
  ls = np.asarray(lengthscale, dtype=np.float64)
  diff = (vx1[:, None, :] - vx2[None, :, :]) / ls
  return np.sum(diff * diff, axis=-1), diff


def _kernel_from_r2(name: str, r2, signal_variance):
  sv = float(np.squeeze(signal_variance))
  if name == "squared_exponential":  # kernel.py:63-81
    return sv * np.exp(-r2 / 2.0)
  if name == "matern32":  # kernel.py:84-102, linalg.safe_l2norm
    r = math.sqrt(3.0) * np.sqrt(r2)
    return sv * (1.0 + r) * np.exp(-r)
  if name == "matern52":  # kernel.py:105-123
    r = math.sqrt(5.0) * np.sqrt(r2)
    return sv * (1.0 + r + r * r / 3.0) * np.exp(-r)
  raise NotImplementedError(name)


def cov_matrix(name: str, model: Dict, vx1, vx2=None, warp_func=None,
               diag: bool = False):
  """covariance_matrix.matrix_map, kernel.py:33-58.

  diag=True is honoured only when vx2 is None (kernel.py:54-56) and then
  returns an (n1,) vector.
  """
  lengthscale, signal_variance = retrieve_params(
      model, ["lengthscale", "signal_variance"], warp_func)
  vx1 = np.asarray(vx1, dtype=np.float64)
  if vx2 is None:
    if diag:
      return np.full((vx1.shape[0],), float(np.squeeze(signal_variance)))
    vx2 = vx1
  r2, _ = _scaled_sqdist(vx1, vx2, lengthscale)
  return _kernel_from_r2(name, r2, signal_variance)


def mean_vector(name: str, model: Dict, vx, warp_func=None):
  """mean.constant / mean.zero through mean_vector, mean.py:30-64 -> (n,1)."""
  n = np.asarray(vx).shape[0]
  if name == "zero":
    return np.zeros((n, 1))
  if name == "constant":
    (val,) = retrieve_params(model, ["constant"], warp_func)
    return np.full((n, 1), float(np.squeeze(val)))
  raise NotImplementedError(name)


# ----------------------------------------------------------------------------
# linear algebra  (basics/linalg.py:29-110)
# ----------------------------------------------------------------------------
def compute_delta_y_and_cov(mean_name, cov_name, model, x, y, warp_func=None,
                            eps=JITTER):
  """linalg.py:36-69: y - m(x),  K(x,x) + I*(noise_variance + eps)."""
  y = np.asarray(y, dtype=np.float64) - mean_vector(mean_name, model, x,
                                                    warp_func)
  (noise_variance,) = retrieve_params(model, ["noise_variance"], warp_func)
  cov = cov_matrix(cov_name, model, x, warp_func=warp_func)
  cov = cov + np.eye(len(x)) * (float(np.squeeze(noise_variance)) + eps)
  return y, cov


def solve_gp_linear_system(mean_name, cov_name, model, x, y, warp_func=None,
                           eps=JITTER):
  """linalg.py:72-110 -> (chol lower, kinvy, y - mean)."""
  y, cov = compute_delta_y_and_cov(mean_name, cov_name, model, x, y,
                                   warp_func, eps)
  chol = np.linalg.cholesky(cov)  # linalg.py:31 (lower=True)
  kinvy = spla.cho_solve((chol, True), y)  # linalg.py:144
  return chol, kinvy, y


# ----------------------------------------------------------------------------
# objective  (gp_utils/objectives.py:109-210, Cholesky branch)
# ----------------------------------------------------------------------------
def nll_sub_dataset(mean_name, cov_name, model, vx, vy, warp_func=None):
  """objectives.py:144-156."""
  chol, kinvy, vy = solve_gp_linear_system(mean_name, cov_name, model, vx, vy,
                                           warp_func)
  return float(
      np.sum(0.5 * np.dot(vy.T, kinvy) + np.sum(np.log(np.diag(chol))) +
             0.5 * len(vx) * np.log(2 * np.pi)))


def nll_sub_dataset_svd(mean_name, cov_name, model, vx, vy, warp_func=None):
  """objectives.py:157-176 (the branch GP.stats() uses)."""
  vy, cov = compute_delta_y_and_cov(mean_name, cov_name, model, vx, vy,
                                    warp_func)
  u, s, vt = np.linalg.svd(cov)
  kinv = vt.T @ (np.diag(1.0 / s) @ u.T)
  kinvy = kinv @ vy
  return float(0.5 * np.sum(
      np.dot(vy.T, kinvy) + np.sum(np.log(s)) + len(vx) * np.log(2 * np.pi)))


def neg_log_marginal_likelihood(mean_name, cov_name, model, dataset,
                                warp_func=None, exclude_aligned=True,
                                return_key2nll=False, use_cholesky=True):
  """objectives.py:178-195.  dataset: dict key -> (x, y[, aligned])."""
  total, key2nll, num = 0.0, {}, 0
  for k, s in dataset.items():
    x, y = s[0], s[1]
    aligned = s[2] if len(s) > 2 else None
    if exclude_aligned and aligned is not None:
      continue
    if np.asarray(x).shape[0] == 0:
      continue
    f = nll_sub_dataset if use_cholesky else nll_sub_dataset_svd
    key2nll[k] = f(mean_name, cov_name, model, x, y, warp_func)
    total += key2nll[k]
    num += 1
  total = 0.0 if num == 0 else total / num
  if return_key2nll:
    return total, key2nll
  return total


# ----------------------------------------------------------------------------
# closed-form gradient of the mean NLL w.r.t. the RAW (un-warped) parameters
# (what jax.value_and_grad(loss_func) yields at gp.py:134)
# ----------------------------------------------------------------------------
def _pair_weight(cov_name: str, r2, k_nodiag, signal_variance):
  """W such that dK/dl_k = W * Delta_k^2 / l_k^3  (Delta = x1-x2, unscaled).

  SE:  W = K.  M32: W = 3 sv e^{-r}.  M52: W = (5/3) sv (1+r) e^{-r}.
  The reference's _safe_sqrt (linalg.py:175-197) makes the r=0 entries
  contribute exactly 0, which these closed forms reproduce because
  Delta_k^2 = 0 there.
  """
  sv = float(np.squeeze(signal_variance))
  if cov_name == "squared_exponential":
    return k_nodiag
  if cov_name == "matern32":
    r = math.sqrt(3.0) * np.sqrt(r2)
    return 3.0 * sv * np.exp(-r)
  if cov_name == "matern52":
    r = math.sqrt(5.0) * np.sqrt(r2)
    return (5.0 / 3.0) * sv * (1.0 + r) * np.exp(-r)
  raise NotImplementedError(cov_name)


def nll_and_grad_sub_dataset(mean_name, cov_name, model, vx, vy,
                             warp_func=None):
  """Per-task nll and d nll / d raw for keys constant, lengthscale (d,),
  signal_variance, noise_variance.  Lengthscale may be scalar or (d,)."""
  vx = np.asarray(vx, dtype=np.float64)
  n, d = vx.shape
  ls_raw = np.asarray(model["lengthscale"], dtype=np.float64)
  ls, sv, nv = retrieve_params(
      model, ["lengthscale", "signal_variance", "noise_variance"], warp_func)
  ls_full = np.broadcast_to(np.asarray(ls, dtype=np.float64), (d,))
  sv = float(np.squeeze(sv))
  nv = float(np.squeeze(nv))
  r = np.asarray(vy, dtype=np.float64) - mean_vector(mean_name, model, vx,
                                                     warp_func)
  r2, diff = _scaled_sqdist(vx, vx, ls_full)
  k = _kernel_from_r2(cov_name, r2, sv)
  cov = k + np.eye(n) * (nv + JITTER)
  chol = np.linalg.cholesky(cov)
  alpha = spla.cho_solve((chol, True), r)
  nll = float(0.5 * (r.T @ alpha).item() + np.sum(np.log(np.diag(chol))) +
              0.5 * n * np.log(2 * np.pi))
  kinv = spla.cho_solve((chol, True), np.eye(n))
  g = 0.5 * (kinv - alpha @ alpha.T)  # d nll / d cov
  w = _pair_weight(cov_name, r2, k, sv)
  gw = g * w
  # d/d lengthscale_k = sum_ij G_ij W_ij Delta_k^2 / l_k^3 ; diff = Delta / l
  d_ls = np.einsum("ij,ijk->k", gw, diff * diff) / ls_full
  d_sv = float(np.sum(g * k) / sv)
  d_nv = float(np.trace(g))
  d_c = -float(np.sum(alpha)) if mean_name == "constant" else 0.0

  # chain rule through the warp: softplus'(raw) = sigmoid(raw)
  def warped(key):
    return bool(warp_func) and key in warp_func and \
        warp_func[key] is not identity_warp

  if warped("lengthscale"):
    d_ls = d_ls * sigmoid(np.broadcast_to(ls_raw, (d,)))
  if warped("signal_variance"):
    d_sv *= float(sigmoid(model["signal_variance"]))
  if warped("noise_variance"):
    d_nv *= float(sigmoid(model["noise_variance"]))
  if ls_raw.size == 1:  # scalar lengthscale broadcast over d (kernel.py:80)
    d_ls_out = np.sum(d_ls).reshape(ls_raw.shape)
  else:
    d_ls_out = d_ls.reshape(ls_raw.shape)
  grad = {
      "lengthscale": d_ls_out,
      "signal_variance": d_sv,
      "noise_variance": d_nv,
  }
  if "constant" in model:
    grad["constant"] = d_c
  return nll, grad


def nll_value_and_grad(mean_name, cov_name, model, dataset, warp_func=None):
  """Mean NLL over non-empty, non-aligned tasks and its raw-param gradient."""
  total, num = 0.0, 0
  gsum: Dict[str, np.ndarray] = {}
  for _, s in dataset.items():
    x, y = s[0], s[1]
    aligned = s[2] if len(s) > 2 else None
    if aligned is not None or np.asarray(x).shape[0] == 0:
      continue
    v, g = nll_and_grad_sub_dataset(mean_name, cov_name, model, x, y,
                                    warp_func)
    total += v
    num += 1
    for k2, gv in g.items():
      gsum[k2] = gsum.get(k2, 0.0) + np.asarray(gv, dtype=np.float64)
  if num == 0:
    return 0.0, {k2: np.zeros_like(np.asarray(v, dtype=np.float64))
                 for k2, v in model.items()}
  return total / num, {k2: gv / num for k2, gv in gsum.items()}


# ----------------------------------------------------------------------------
# divergence objectives on ALIGNED sub-datasets
# gp_utils/objectives.py:29-101, gp_utils/utils.py:84-173
# ----------------------------------------------------------------------------
def svd_matrix_sqrt(cov):
  """basics/linalg.py:112-126: A with A A' = cov, full column rank."""
  u, s, _ = np.linalg.svd(cov)
  factor = u * np.sqrt(s[..., None, :])
  tol = s.max() * np.finfo(s.dtype).eps / 2.0 * np.sqrt(2 * cov.shape[0] + 1.0)
  rank = int(np.count_nonzero(s > tol))
  return factor[:, :rank]


def partial_kl_mvn(mu0, cov0, mu1, cov1):
  """utils.py:84-106: tr(cov1^-1 cov0) + (mu1-mu0)' cov1^-1 (mu1-mu0) + logdet cov1."""
  mu_diff = mu1 - mu0
  chol1 = np.linalg.cholesky(cov1)
  cov1invmudiff = spla.cho_solve((chol1, True), mu_diff)
  trcov1invcov0 = float(np.trace(spla.cho_solve((chol1, True), cov0)))
  mahalanobis = float(np.dot(mu_diff, cov1invmudiff))
  logdetcov1 = float(np.sum(2 * np.log(np.diag(chol1))))
  return trcov1invcov0 + mahalanobis + logdetcov1


def kl_multivariate_normal(mu0, cov0, mu1, cov1, weight=1.0, eps=0.0,
                           partial=True):
  """utils.py:109-148."""
  cov0 = np.atleast_2d(np.asarray(cov0, dtype=np.float64))
  cov1 = np.atleast_2d(np.asarray(cov1, dtype=np.float64))
  if eps > 0.0:
    cov0 = cov0 + np.eye(cov0.shape[0]) * eps
    cov1 = cov1 + np.eye(cov1.shape[0]) * eps
  if partial:
    return weight * partial_kl_mvn(mu0, cov0, mu1, cov1)
  chol0 = svd_matrix_sqrt(cov0)
  chol0inv = np.linalg.pinv(chol0)
  mu1 = chol0inv @ (mu1 - mu0)
  cov1 = chol0inv @ cov1 @ chol0inv.T
  mu0 = np.zeros_like(mu1)
  cov0 = np.eye(cov1.shape[0])
  return weight * 0.5 * (partial_kl_mvn(mu0, cov0, mu1, cov1) - chol0.shape[1])


def euclidean_multivariate_normal(mu0, cov0, mu1, cov1, mean_weight=1.0,
                                  cov_weight=1.0, **unused):
  """utils.py:151-173 (safe_l2norm = plain l2 norm in the forward pass)."""
  mean_diff = math.sqrt(float(np.sum((mu0 - mu1)**2)))
  cov_diff = math.sqrt(float(np.sum((np.atleast_2d(cov0) - cov1)**2)))
  return mean_weight * mean_diff + cov_weight * cov_diff


def _data_moments(y):
  """objectives.py:69-70: sample mean / biased sample covariance over the m
  columns of an aligned y (n, m)."""
  y = np.asarray(y, dtype=np.float64)
  mu = np.mean(y, axis=1)
  yc = y - mu[:, None]
  return mu, (yc @ yc.T) / y.shape[1]


def multivariate_normal_divergence(mean_name, cov_name, model, dataset,
                                   warp_func=None,
                                   distance=kl_multivariate_normal):
  """objectives.py:29-101: mean over the non-empty ALIGNED sub-datasets of
  distance(N(data mean, data cov), N(m(x), K(x,x) + noise I))."""
  total, num = 0.0, 0
  for k, s in dataset.items():
    x, y = np.asarray(s[0], dtype=np.float64), np.asarray(s[1], dtype=np.float64)
    aligned = s[2] if len(s) > 2 else None
    if aligned is None or x.shape[0] == 0:
      continue
    if y.shape[1] == 0 or y.shape[0] != x.shape[0]:
      raise ValueError(f"dataset[{k}].x has shape {x.shape} but "
                       f"dataset[{k}].y has shape {y.shape}")
    mu_data, cov_data = _data_moments(y)
    mu_model = mean_vector(mean_name, model, x, warp_func).flatten()
    (nv,) = retrieve_params(model, ["noise_variance"], warp_func)
    cov_model = cov_matrix(cov_name, model, x, warp_func=warp_func) + \
        np.eye(x.shape[0]) * float(np.squeeze(nv))
    total += distance(mu0=mu_data, cov0=cov_data, mu1=mu_model, cov1=cov_model)
    num += 1
  return 0.0 if num == 0 else total / num


def kl_value_and_grad(mean_name, cov_name, model, dataset, warp_func=None,
                      eps=0.0, weight=1.0):
  """Partial-KL divergence (the `kl` / `ekl` / `regkl` objective) and its
  gradient w.r.t. the raw parameters in closed form (matrix calculus, not the
  engine's NLL decomposition): with K1 = K + (nv + eps) I, S = cov0 + eps I +
  d d', d = mu1 - mu0:   G = K1^-1 - K1^-1 S K1^-1,   d kl/d theta = <G, dK1>,
  d kl/d constant = 2 sum(K1^-1 d)."""
  total, num = 0.0, 0
  gsum: Dict[str, np.ndarray] = {}
  for _, s in dataset.items():
    x, y = np.asarray(s[0], dtype=np.float64), np.asarray(s[1], dtype=np.float64)
    aligned = s[2] if len(s) > 2 else None
    if aligned is None or x.shape[0] == 0:
      continue
    n, d = x.shape
    mu0, cov0 = _data_moments(y)
    ls_raw = np.asarray(model["lengthscale"], dtype=np.float64)
    ls, sv, nv = retrieve_params(
        model, ["lengthscale", "signal_variance", "noise_variance"], warp_func)
    ls_full = np.broadcast_to(np.asarray(ls, dtype=np.float64), (d,))
    sv, nv = float(np.squeeze(sv)), float(np.squeeze(nv))
    dvec = mean_vector(mean_name, model, x, warp_func).flatten() - mu0
    r2, diff = _scaled_sqdist(x, x, ls_full)
    k = _kernel_from_r2(cov_name, r2, sv)
    k1 = k + np.eye(n) * (nv + eps)
    chol = np.linalg.cholesky(k1)
    kinv = spla.cho_solve((chol, True), np.eye(n))
    smat = cov0 + np.eye(n) * eps + np.outer(dvec, dvec)
    val = float(np.trace(kinv @ (cov0 + np.eye(n) * eps)) + dvec @ kinv @ dvec +
                2 * np.sum(np.log(np.diag(chol))))
    g = kinv - kinv @ smat @ kinv
    w = _pair_weight(cov_name, r2, k, sv)
    d_ls = np.einsum("ij,ijk->k", g * w, diff * diff) / ls_full
    d_sv = float(np.sum(g * k) / sv)
    d_nv = float(np.trace(g))
    d_c = 2.0 * float(np.sum(kinv @ dvec)) if mean_name == "constant" else 0.0

    def warped(key):
      return bool(warp_func) and key in warp_func and \
          warp_func[key] is not identity_warp

    if warped("lengthscale"):
      d_ls = d_ls * sigmoid(np.broadcast_to(ls_raw, (d,)))
    if warped("signal_variance"):
      d_sv *= float(sigmoid(model["signal_variance"]))
    if warped("noise_variance"):
      d_nv *= float(sigmoid(model["noise_variance"]))
    grad = {
        "lengthscale": (np.sum(d_ls).reshape(ls_raw.shape) if ls_raw.size == 1
                        else d_ls.reshape(ls_raw.shape)),
        "signal_variance": d_sv,
        "noise_variance": d_nv,
    }
    if "constant" in model:
      grad["constant"] = d_c
    total += weight * val
    num += 1
    for k2, gv in grad.items():
      gsum[k2] = gsum.get(k2, 0.0) + weight * np.asarray(gv, dtype=np.float64)
  if num == 0:
    return 0.0, {k2: np.zeros_like(np.asarray(v, dtype=np.float64))
                 for k2, v in model.items()}
  return total / num, {k2: gv / num for k2, gv in gsum.items()}


# ----------------------------------------------------------------------------
# Adam loop  (gp_utils/gp.py:114-157; optax.adam defaults)
# ----------------------------------------------------------------------------
class Adam:
  """optax.adam(lr): b1=.9 b2=.999 eps=1e-8 eps_root=0, bias-corrected."""

  def __init__(self, lr, b1=0.9, b2=0.999, eps=1e-8):
    self.lr, self.b1, self.b2, self.eps = lr, b1, b2, eps
    self.t = 0
    self.m: Dict[str, np.ndarray] = {}
    self.v: Dict[str, np.ndarray] = {}

  def update(self, params: Dict, grads: Dict) -> Dict:
    self.t += 1
    out = {}
    for k, p in params.items():
      g = np.asarray(grads[k], dtype=np.float64)
      m = self.b1 * self.m.get(k, 0.0) + (1 - self.b1) * g
      v = self.b2 * self.v.get(k, 0.0) + (1 - self.b2) * g * g
      self.m[k], self.v[k] = m, v
      mhat = m / (1 - self.b1**self.t)
      vhat = v / (1 - self.b2**self.t)
      out[k] = np.asarray(p, dtype=np.float64) - self.lr * mhat / (
          np.sqrt(vhat) + self.eps)
    return out


def sub_sample_dataset_iterator(rng: np.random.Generator, dataset,
                                batch_size):
  """basics/data_utils.py:72-100 with a NumPy Generator instead of jax.random
  (threefry is unavailable here; only the *shape* of the draw is pinned)."""
  while True:
    out = {}
    for i, (k, s) in enumerate(dataset.items()):
      x, y = s[0], s[1]
      aligned = s[2] if len(s) > 2 else None
      if x.shape[0] >= batch_size:
        idx = rng.permutation(x.shape[0])[:batch_size]
        x, y = x[idx, :], y[idx, :]
      if isinstance(aligned, str):
        aligned = i
      out[k] = (x, y, aligned)
    yield out


def infer_parameters_adam(mean_name, cov_name, model, dataset, warp_func,
                          learning_rate, max_training_step, batch_size,
                          rng=None, callback=None):
  """gp.py:114-157.  Returns (final model dict, list of per-step losses)."""
  if not dataset or max_training_step <= 0:
    return dict(model), []
  rng = rng or np.random.default_rng(0)
  it = sub_sample_dataset_iterator(rng, dataset, batch_size)
  opt = Adam(learning_rate)
  accepted = dict(model)
  cur = dict(model)
  losses = []
  batch = None
  for i in range(max_training_step):
    batch = next(it)
    loss, grads = nll_value_and_grad(mean_name, cov_name, cur, batch,
                                     warp_func)
    if np.isnan(loss) and i == 0:
      raise ValueError("Encountered NaN in loss function.")
    if np.isfinite(loss):
      accepted = cur
    else:
      break
    losses.append(loss)
    cur = opt.update(cur, grads)
    if callback:
      callback(i, accepted, loss)
  if batch is not None:
    loss = neg_log_marginal_likelihood(mean_name, cov_name, cur, batch,
                                       warp_func)
    if np.isfinite(loss):
      accepted = cur
  return accepted, losses


# ----------------------------------------------------------------------------
# predict  (gp_utils/gp.py:242-305, 562-620)
# ----------------------------------------------------------------------------
def predict(mean_name, cov_name, model, x_observed, y_observed, x_query,
            warp_func=None, full_cov=False, cache=None):
  """gp.predict, gp.py:242-305.  cache = (chol, kinvy) or None."""
  x_query = np.asarray(x_query, dtype=np.float64)
  if x_observed is None or np.asarray(x_observed).shape[0] == 0:
    mu = mean_vector(mean_name, model, x_query, warp_func)
    cov = cov_matrix(cov_name, model, x_query, warp_func=warp_func,
                     diag=not full_cov)
    return (mu, cov) if full_cov else (mu, cov[:, None])
  if cache is None:
    chol, kinvy, _ = solve_gp_linear_system(mean_name, cov_name, model,
                                            x_observed, y_observed, warp_func)
  else:
    chol, kinvy = cache
  cov = cov_matrix(cov_name, model, x_observed, x_query, warp_func=warp_func)
  mu = cov.T @ kinvy + mean_vector(mean_name, model, x_query, warp_func)
  v = spla.solve_triangular(chol, cov, lower=True)
  if full_cov:
    return mu, cov_matrix(cov_name, model, x_query,
                          warp_func=warp_func) - v.T @ v
  var = cov_matrix(cov_name, model, x_query, warp_func=warp_func,
                   diag=True) - np.sum(v * v, axis=0)
  return mu, var[:, None]


def gp_predict(mean_name, cov_name, model, dataset, x_query, sub_dataset_key=0,
               warp_func=None, full_cov=False, with_noise=True, unbiased=True):
  """GP.predict, gp.py:562-620 (noise without jitter; N/(N-1) inflation)."""
  if sub_dataset_key not in dataset:
    mu, cov = predict(mean_name, cov_name, model, None, None, x_query,
                      warp_func, full_cov)
  else:
    s = dataset[sub_dataset_key]
    mu, cov = predict(mean_name, cov_name, model, s[0], s[1], x_query,
                      warp_func, full_cov)
  cov = np.array(cov, dtype=np.float64)
  if with_noise:
    (nv,) = retrieve_params(model, ["noise_variance"], warp_func)
    nv = float(np.squeeze(nv))
    if full_cov:
      cov = cov + np.eye(cov.shape[0]) * nv
    else:
      cov = cov + nv
  if unbiased:
    n_ds = len([k for k, v in dataset.items()
                if (v[2] if len(v) > 2 else None) is None])
    if n_ds > 1:
      cov = cov * (n_ds / (n_ds - 1.0))
  return mu, cov


# ----------------------------------------------------------------------------
# acquisition  (bo_utils/acfun.py:96-165)
# ----------------------------------------------------------------------------
def _norm_pdf(z):
  return np.exp(-0.5 * z * z) / math.sqrt(2 * math.pi)


def _norm_cdf(z):
  return 0.5 * (1.0 + erf(z / math.sqrt(2.0)))


def expected_improvement_sub(mu, std, target):  # acfun.py:96-110
  gamma = (target - mu) / std
  return (_norm_pdf(gamma) - gamma * (1 - _norm_cdf(gamma))) * std


def probability_of_improvement_sub(mu, std, target):  # acfun.py:113-126
  return -((target - mu) / std)


def ucb_sub(mu, std, beta=3.0):  # acfun.py:129-142
  return mu + beta * std


def ei_target(dataset, key):  # acfun.py:145-148
  if key not in dataset or np.asarray(dataset[key][1]).shape[0] == 0:
    return 0.0
  return float(np.max(dataset[key][1]))


def pi_target(dataset, key, zeta=0.1, use_std=False):  # acfun.py:159-165
  if key not in dataset or np.asarray(dataset[key][1]).shape[0] == 0:
    return 0.0
  y = np.asarray(dataset[key][1], dtype=np.float64)
  if use_std:
    return float(np.max(y) + zeta * np.std(y))
  return float(np.max(y) + zeta)


def acquisition(name, mean_name, cov_name, model, dataset, key, x_queries,
                warp_func=None):
  """acfun_wrapper.acquisition_function, acfun.py:51-91 (non-HGP branch)."""
  mu, var = gp_predict(mean_name, cov_name, model, dataset, x_queries, key,
                       warp_func, full_cov=False, with_noise=True)
  std = np.sqrt(var)
  if name in ("expected_improvement", "ei"):
    return expected_improvement_sub(mu, std, ei_target(dataset, key))
  if name in ("probability_of_improvement", "pi"):
    return probability_of_improvement_sub(mu, std, pi_target(dataset, key))
  if name == "pi2":
    return probability_of_improvement_sub(
        mu, std, pi_target(dataset, key, use_std=True))
  if name == "pi3":
    return probability_of_improvement_sub(
        mu, std, pi_target(dataset, key, zeta=0.05))
  if name in ("ucb", "ucb3"):
    return ucb_sub(mu, std, 3.0)
  if name == "ucb2":
    return ucb_sub(mu, std, 2.0)
  if name == "ucb4":
    return ucb_sub(mu, std, 4.0)
  raise NotImplementedError(name)


# ----------------------------------------------------------------------------
# synthetic workloads  (recipe of bo_utils/data.py:720-775 + gp.py:198-239,
# PCG64 streams instead of jax threefry -- SURVEY.md 8d)
# ----------------------------------------------------------------------------
GROUND_TRUTH = {  # gp_test.py:64-70, used un-warped
    "constant": 5.0,
    "lengthscale": 1.0,
    "signal_variance": 1.0,
    "noise_variance": 0.01,
}


def init_raw_params(d: int) -> Dict:
  """gp_test.py:102-108 initial raw parameters, ARD-broadcast (gp.py:395-400)."""
  return {
      "constant": 5.1,
      "lengthscale": np.zeros(d),
      "signal_variance": 0.0,
      "noise_variance": -4.0,
  }


def make_task(t: int, n: int, d: int, cov_name="squared_exponential",
              surrogate=False):
  """One synthetic task: X ~ U[0,1]^{n x d} (seed 1000+t); y = one draw from
  the ground-truth GP (seed 2000+t), or the cheap surrogate for large n."""
  x = np.random.Generator(np.random.PCG64(1000 + t)).random((n, d))
  z = np.random.Generator(np.random.PCG64(2000 + t)).standard_normal((n, 1))
  if surrogate:
    y = GROUND_TRUTH["constant"] + np.sum(np.sin(2 * np.pi * x), axis=1,
                                          keepdims=True) + 0.1 * z
    return x, y
  k = cov_matrix(cov_name, GROUND_TRUTH, x)
  chol = np.linalg.cholesky(
      k + np.eye(n) * (GROUND_TRUTH["noise_variance"] + JITTER))
  y = GROUND_TRUTH["constant"] + chol @ z
  return x, y


def make_dataset(num_tasks, n, d, cov_name="squared_exponential",
                 surrogate=False, ragged_seed=None, ragged_lo=None,
                 ragged_hi=None):
  ns = [n] * num_tasks
  if ragged_seed is not None:
    rr = np.random.Generator(np.random.PCG64(ragged_seed))
    ns = [int(v) for v in rr.integers(ragged_lo, ragged_hi + 1, num_tasks)]
  return {t: make_task(t, ns[t], d, cov_name, surrogate)
          for t in range(num_tasks)}
