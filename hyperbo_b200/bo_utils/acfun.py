"""Acquisition functions -- mirrors hyperbo/bo_utils/acfun.py:36-187.

acquisition(*, model, sub_dataset_key, x_queries, acfun_callback=...) -> (nq,1).
For a gp.GP model the predictive mean/variance and the acfun_sub epilogue run
as one fused engine call (hb_predict with acq_id)."""
from __future__ import annotations

import functools
from typing import Any, Callable, Union

import torch

from hyperbo_b200 import engine as _engine
from hyperbo_b200.gp_utils import gp

partial = functools.partial


def random_search(model, x_queries, **unused_kwargs):
  """Returns a uniformly sampled random array (acfun.py:27-33)."""
  assert model.rng is not None, "Random search requires random key."
  gen = model.rng if isinstance(model.rng, torch.Generator) else \
      torch.Generator().manual_seed(int(model.rng))
  if not isinstance(model.rng, torch.Generator):
    model.rng = int(model.rng) + 1
  n = torch.as_tensor(x_queries).shape[0]
  return torch.rand((n, 1), generator=gen, dtype=torch.float64)


def _sub(acq_id):

  def f(mu, std, param):
    eng = _engine.Engine.get()
    std = eng.tensor(std)
    return eng.acquisition(acq_id, float(param), mu, std * std)

  return f


_ei_sub, _pi_sub, _ucb_sub = _sub(1), _sub(2), _sub(3)


def expected_improvement_sub(mu, std, target):
  """(pdf(g) - g (1 - cdf(g))) std,  g = (target - mu)/std  (acfun.py:96-110)."""
  return _ei_sub(mu, std, target)


def probability_of_improvement_sub(mu, std, target):
  """-g  (acfun.py:113-126)."""
  return _pi_sub(mu, std, target)


def ucb_sub(mu, std, beta=3.):
  """mu + beta std  (acfun.py:129-142)."""
  return _ucb_sub(mu, std, beta)


expected_improvement_sub.hb_acq_id = 1
probability_of_improvement_sub.hb_acq_id = 2
ucb_sub.hb_acq_id = 3


def acfun_wrapper(acfun_sub: Callable[..., Any],
                  acfun_callback_default: Callable[..., Any]):
  """Wrapper for sub acquisition function (acfun.py:36-93)."""

  def acquisition_function(*, model, sub_dataset_key: Union[int, str],
                           x_queries, acfun_callback=acfun_callback_default):
    acq_id = getattr(acfun_sub, "hb_acq_id", None)
    if isinstance(model, gp.HGP):
      predicts = model.predict(x_queries, sub_dataset_key=sub_dataset_key,
                               full_cov=False, with_noise=True)
      acfun_param = acfun_callback(model, sub_dataset_key)
      ac_vals = [acfun_sub(mu, torch.sqrt(var), acfun_param)
                 for mu, var in predicts]
      return torch.mean(torch.stack(ac_vals), dim=0)
    acfun_param = float(acfun_callback(model, sub_dataset_key))
    if acq_id is not None and hasattr(model, "acquisition"):
      return model.acquisition(x_queries, sub_dataset_key, acq_id, acfun_param)
    mu, var = model.predict(x_queries, sub_dataset_key=sub_dataset_key,
                            full_cov=False, with_noise=True)
    return acfun_sub(mu, torch.sqrt(var), acfun_param)

  return acquisition_function


def ei_callback_default(model, key, **unused_kwargs):
  """acfun.py:145-148."""
  if key not in model.dataset or model.dataset[key].y.shape[0] == 0:
    return 0.0
  return float(torch.max(model.dataset[key].y))


expected_improvement = acfun_wrapper(
    acfun_sub=expected_improvement_sub,
    acfun_callback_default=ei_callback_default)
ei = expected_improvement


def pi_callback_default(model, key, zeta=0.1, use_std=False, **unused_kwargs):
  """acfun.py:159-165 (jnp.std is the population std)."""
  if key not in model.dataset or model.dataset[key].y.shape[0] == 0:
    return 0.0
  y = model.dataset[key].y
  if use_std:
    return float(torch.max(y) + zeta * torch.std(y, unbiased=False))
  return float(torch.max(y) + zeta)


probability_of_improvement = acfun_wrapper(
    acfun_sub=probability_of_improvement_sub,
    acfun_callback_default=pi_callback_default)
pi = probability_of_improvement
pi2 = acfun_wrapper(
    acfun_sub=probability_of_improvement_sub,
    acfun_callback_default=partial(pi_callback_default, use_std=True))
pi3 = acfun_wrapper(
    acfun_sub=probability_of_improvement_sub,
    acfun_callback_default=partial(pi_callback_default, zeta=0.05))

ucb4 = acfun_wrapper(acfun_sub=ucb_sub, acfun_callback_default=lambda a, b: 4.)
ucb3 = acfun_wrapper(acfun_sub=ucb_sub, acfun_callback_default=lambda a, b: 3.)
ucb2 = acfun_wrapper(acfun_sub=ucb_sub, acfun_callback_default=lambda a, b: 2.)
ucb = ucb3

# what the device-resident BO loop (bayesopt.simulated_bayesopt fast path) needs
# to know about an acquisition whose callback is the default one:
# (acq_id, parameter, parameter is an offset on max(y_observed))
expected_improvement.hb_device = (1, 0.0, True)
probability_of_improvement.hb_device = (2, 0.1, True)
pi3.hb_device = (2, 0.05, True)
ucb4.hb_device = (3, 4.0, False)
ucb3.hb_device = (3, 3.0, False)
ucb2.hb_device = (3, 2.0, False)

random_search.__name__ = "random_search"
rand = random_search


def shard_candidates(ac_func):
  """Candidate-axis sharding of an acquisition sweep over the ranks of
  torch.distributed: the factorisation (L, alpha) is replicated
  -- every rank holds the same model -- each rank evaluates a contiguous slice
  of `x_queries`, and one all-gather gives every rank the full (nq, 1) vector,
  so a following arg-max picks the same candidate everywhere.  Collective: all
  ranks must call it with the same arguments.  Without an initialised process
  group (or with one rank) it is `ac_func` itself."""

  def acquisition_function(*, model, sub_dataset_key, x_queries, **kwargs):
    import torch.distributed as dist
    rank, world = gp._dist_world()  # pylint: disable=protected-access
    nq = torch.as_tensor(x_queries).shape[0]
    if world == 1 or nq < world:
      return ac_func(model=model, sub_dataset_key=sub_dataset_key,
                     x_queries=x_queries, **kwargs)
    chunk = -(-nq // world)
    lo, hi = min(rank * chunk, nq), min((rank + 1) * chunk, nq)
    local = ac_func(model=model, sub_dataset_key=sub_dataset_key,
                    x_queries=x_queries[lo:hi], **kwargs).reshape(-1)
    padded = torch.zeros(chunk, dtype=local.dtype, device=local.device)
    padded[:hi - lo] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    return torch.cat(parts)[:nq].reshape(-1, 1)

  acquisition_function.__name__ = getattr(ac_func, "__name__",
                                          "acquisition_function")
  return acquisition_function
