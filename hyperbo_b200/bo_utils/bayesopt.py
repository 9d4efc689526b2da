"""Bayesian-optimisation drivers -- mirror hyperbo/bo_utils/bayesopt.py:34-302
(get_best_datapoint, retrain_model, simulated_bayesopt, bayesopt, run_bayesopt)
on top of the engine-backed gp.GP.  Host logic only: per iteration one append,
one (re)factorisation of the queried task and one fused predict+acquisition
sweep over all candidates on the GPU, then an arg-max read-back."""
from __future__ import annotations

import logging
import time

import numpy as np
import scipy.optimize
import torch

from hyperbo_b200.basics import definitions as defs
from hyperbo_b200.gp_utils import gp
from hyperbo_b200.gp_utils import objectives as obj

SubDataset = defs.SubDataset
INPUT_SAMPLERS = {}


def _gen(key):
  if isinstance(key, torch.Generator):
    return key
  g = torch.Generator()
  g.manual_seed(int(key) if key is not None else 0)
  return g


def get_best_datapoint(sub_dataset: SubDataset):
  """Best (x, y) in a sub-dataset (bayesopt.py:34-43)."""
  best_idx = int(torch.argmax(torch.as_tensor(sub_dataset.y)))
  return sub_dataset.x[best_idx], sub_dataset.y[best_idx]


def retrain_model(model: gp.GP, sub_dataset_key, random_key=None,
                  get_params_path=None, callback=None):
  """Retrain with more observations when config['retrain'] > 0
  (bayesopt.py:46-72)."""
  cfg = model.params.config
  if not ("retrain" in cfg and cfg["retrain"] > 0 and
          model.dataset[sub_dataset_key].x.shape[0] > 0):
    return
  if cfg["objective"] in (obj.regkl, obj.regeuc):
    raise ValueError("Objective must include NLL to retrain.")
  cfg["max_training_step"] = cfg["retrain"]
  model.train(random_key, get_params_path=get_params_path, callback=callback)


def simulated_bayesopt(model: gp.GP, sub_dataset_key,
                       queried_sub_dataset: SubDataset, ac_func, iters: int,
                       random_key=None, get_params_path=None, callback=None
                       ) -> SubDataset:
  """Simulated BO over a finite, pre-evaluated candidate set
  (bayesopt.py:137-193)."""
  xq = queried_sub_dataset.x
  yq = queried_sub_dataset.y
  gen = _gen(random_key) if random_key is not None else None
  if _device_loop_ok(model, ac_func, iters):
    return _simulated_bayesopt_device(model, sub_dataset_key, xq, yq, ac_func, iters)
  for _ in range(iters):
    retrain_model(model, sub_dataset_key=sub_dataset_key, random_key=gen,
                  get_params_path=get_params_path, callback=callback)
    if ac_func.__name__ in ("rand", "random_search"):
      if gen is None:
        raise ValueError("Must specify a random key for random search.")
      select_idx = int(torch.randint(0, xq.shape[0], (1,), generator=gen))
    else:
      evals = ac_func(model=model, sub_dataset_key=sub_dataset_key,
                      x_queries=xq)
      select_idx = int(evals.argmax())
    model.update_sub_dataset((xq[select_idx], yq[select_idx]),
                             sub_dataset_key=sub_dataset_key, is_append=True)
  return model.dataset.get(sub_dataset_key,
                           SubDataset(torch.empty(0), torch.empty(0)))


def _device_loop_ok(model, ac_func, iters) -> bool:
  """The device-resident loop (hb_bo_step) serves the plain-GP case with a
  default-callback EI / PI / UCB and no per-iteration retraining."""
  import os
  from hyperbo_b200 import engine as _engine
  cfg = model.params.config or {}
  if getattr(_engine.Engine.get(), "h", None) is None:  # (the CPU test double)
    return False
  return (iters > 0 and getattr(ac_func, "hb_device", None) is not None and
          type(model) is gp.GP and not cfg.get("retrain", 0) and
          hasattr(model, "engine_ids") and
          os.environ.get("HB_BO_DEVICE", "1") != "0")


def _simulated_bayesopt_device(model, sub_dataset_key, xq, yq, ac_func, iters):
  """bayesopt.py:169-193 with everything on the device: per iteration ONE
  engine call (acquisition over all candidates, arg-max, append, O(n^2) rank-1
  re-conditioning); the chosen indices are read back once at the end."""
  from hyperbo_b200 import engine as _engine
  eng = _engine.Engine.get()
  if sub_dataset_key in model.dataset:
    x0, y0 = model.dataset[sub_dataset_key].x, model.dataset[sub_dataset_key].y
  else:
    x0 = y0 = None
  xq_d = eng.tensor(xq)
  yq_d = eng.tensor(yq).reshape(-1)
  kid, mid, raw, mask = model.engine_ids(int(xq_d.shape[1]))
  n0 = 0 if x0 is None else int(torch.as_tensor(x0).shape[0])
  sess = _engine.BoSession(eng, kid, mid, x0 if n0 else None, y0 if n0 else None, raw,
                           mask, n0 + iters, d=int(xq_d.shape[1]))
  acq_id, param, on_ymax = ac_func.hb_device
  # GP.predict conventions (gp.py:607-619): noise without jitter, N/(N-1) with N
  # = number of non-aligned sub-datasets INCLUDING the queried one once it exists
  base = len([k for k, v in model.dataset.items() if v.aligned is None])
  key_is_new = sub_dataset_key not in model.dataset
  for it in range(iters):
    # (the first append creates the queried key, gp.py:441-445)
    nds = base + (1 if key_is_new and it > 0 else 0)
    scale = nds / (nds - 1.0) if nds > 1 else 1.0
    sess.step(xq_d, yq_d, acq_id, param, on_ymax, noise_flag=1.0, var_scale=scale)
  sel = sess.selected().long()
  xq_t, yq_t = torch.as_tensor(xq), torch.as_tensor(yq)
  for i in sel.tolist():  # replay the appends on the host-side dataset (bookkeeping)
    model.update_sub_dataset((xq_t[i], yq_t[i]), sub_dataset_key=sub_dataset_key,
                             is_append=True)
  return model.dataset.get(sub_dataset_key,
                           SubDataset(torch.empty(0), torch.empty(0)))


def bayesopt(key, model: gp.GP, sub_dataset_key, query_oracle, ac_func,
             iters: int, input_sampler) -> SubDataset:
  """BO over a continuous domain [0,1]^d (bayesopt.py:75-134): arg-max over
  sampled starting points, then a bounded L-BFGS-B refinement of the
  acquisition.  The reference differentiates the acquisition with autodiff
  (jaxopt); here scipy estimates the gradient by finite differences, each
  evaluation being one single-query engine call -- host-bound, off the hot
  path."""
  gen = _gen(key)
  input_dim = model.input_dim
  for i in range(iters):
    start = time.time()
    retrain_model(model, sub_dataset_key=sub_dataset_key)
    x_samples = torch.as_tensor(input_sampler(gen, input_dim),
                                dtype=torch.float64)
    if ac_func.__name__ in ("rand", "random_search"):
      select_idx = int(torch.randint(0, x_samples.shape[0], (1,),
                                     generator=gen))
    else:
      evals = ac_func(model=model, sub_dataset_key=sub_dataset_key,
                      x_queries=x_samples)
      select_idx = int(evals.argmax())
    x_init = x_samples[select_idx].cpu().numpy()

    def f(x):
      return -float(ac_func(model=model, sub_dataset_key=sub_dataset_key,
                            x_queries=np.asarray(x)[None, :]).reshape(-1)[0])

    res = scipy.optimize.minimize(f, x_init, method="L-BFGS-B",
                                  bounds=[(0.0, 1.0)] * input_dim)
    x_new = torch.as_tensor(res.x, dtype=torch.float64)
    eval_datapoint = x_new, query_oracle(x_new[None, :])
    logging.info(msg=f"{i}-th iter, x_init={x_init}, "
                 f"eval_datapoint={eval_datapoint}, "
                 f"elpased_time={time.time() - start}")
    model.update_sub_dataset(eval_datapoint, sub_dataset_key=sub_dataset_key,
                             is_append=True)
  return model.dataset.get(sub_dataset_key,
                           SubDataset(torch.empty(0), torch.empty(0)))


def run_bayesopt(dataset, sub_dataset_key, queried_sub_dataset, mean_func,
                 cov_func, init_params: defs.GPParams, ac_func, iters: int,
                 warp_func=None, init_random_key=None, method: str = "hyperbo",
                 init_model: bool = False, data_loader_name: str = "",
                 get_params_path=None, callback=None,
                 save_retrain_model: bool = False):
  """BO experiment (bayesopt.py:196-302).  Returns ((x, y) observations, best
  query, params).  The HGP / slice-sampling methods are not available (the
  sampler is absent from the reference snapshot itself, gp.py:192-193)."""
  if method in ("hyperbo_ss",):
    raise NotImplementedError("slice-sampling HGP (absent from the reference)")
  model = gp.GP(dataset=dataset, mean_func=mean_func, cov_func=cov_func,
                params=init_params, warp_func=warp_func)
  gen = _gen(init_random_key)
  if init_model:
    assert init_random_key is not None, ("Cannot initialize with "
                                         "init_random_key == None.")
    model.initialize_params(gen)
    model.train(gen, get_params_path, callback=callback)
  else:
    model.rng = gen
  if isinstance(queried_sub_dataset, SubDataset):
    best_query = get_best_datapoint(queried_sub_dataset)
    sub_dataset = simulated_bayesopt(
        model=model, sub_dataset_key=sub_dataset_key,
        queried_sub_dataset=queried_sub_dataset, ac_func=ac_func, iters=iters,
        random_key=gen,
        get_params_path=get_params_path if save_retrain_model else None,
        callback=callback if save_retrain_model else None)
    return (sub_dataset.x, sub_dataset.y), best_query, model.params
  if data_loader_name not in INPUT_SAMPLERS:
    raise NotImplementedError(
        f"Input sampler for {data_loader_name} not found.")
  sub_dataset = bayesopt(key=gen, model=model, sub_dataset_key=sub_dataset_key,
                         query_oracle=queried_sub_dataset, ac_func=ac_func,
                         iters=iters,
                         input_sampler=INPUT_SAMPLERS[data_loader_name])
  return (sub_dataset.x, sub_dataset.y), None, model.params
