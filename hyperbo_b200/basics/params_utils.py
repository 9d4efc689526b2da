"""Parameter retrieval, saving / loading and logging -- mirrors
hyperbo/basics/params_utils.py:33-111,193-207, plus the model-dict <-> raw-vector
packing of the C ABI (include/hyperbo_b200.h)."""
from __future__ import annotations

import logging
import os
import pickle
from typing import Any, Callable, Dict, List, Optional, Tuple, Union

import numpy as np
import torch

from hyperbo_b200.basics import definitions as defs
from hyperbo_b200.gp_utils import utils

GPParams = defs.GPParams


def _verify_params(model_params: Dict[str, Any], expected_keys: List[str]):
  """params_utils.py:90-94."""
  if not set(expected_keys).issubset(set(model_params.keys())):
    raise ValueError(f"Expected parameters are {sorted(expected_keys)}, "
                     f"but received {sorted(model_params.keys())}.")


def retrieve_params(params: GPParams, keys: List[str],
                    warp_func: Optional[Dict[str, Callable[[Any], Any]]] = None
                    ) -> List[Any]:
  """Returns a list of parameter values (warped if specified) by keys' order
  (params_utils.py:97-111)."""
  model_params = params.model
  _verify_params(model_params, keys)
  if warp_func:
    return [warp_func[k](model_params[k]) if k in warp_func else model_params[k]
            for k in keys]
  return [model_params[k] for k in keys]


# ------------------------------------------------ checkpoints (host only) ---
FINAL_PARAM_FILE_INFO = "FINAL"


def _portable(tree):
  """Callables -> their str (params_utils.py:78-80); device tensors -> numpy, so a
  file written on a GPU box loads anywhere."""
  if isinstance(tree, dict):
    return {k: _portable(v) for k, v in tree.items()}
  if isinstance(tree, (list, tuple)):
    return type(tree)(_portable(v) for v in tree)
  if isinstance(tree, torch.Tensor):
    return tree.detach().cpu().numpy()
  return str(tree) if callable(tree) else tree


def save_to_file(filenm: str, state: Any = None):
  """params_utils.py:45-53 (plain files instead of gfile)."""
  if not state:
    return
  dirnm = os.path.dirname(filenm)
  if dirnm and not os.path.exists(dirnm):
    os.makedirs(dirnm, exist_ok=True)
  with open(filenm, "wb") as f:
    pickle.dump(state, f)


def load_from_file(filenm: str):
  """params_utils.py:56-61."""
  if not os.path.exists(filenm):
    raise FileNotFoundError(f"{filenm} does not exist.")
  with open(filenm, "rb") as f:
    return pickle.load(f)


def save_params(filenm: str, params: Union[GPParams, Dict[str, Any]],
                state: Any = None):
  """params_utils.py:64-73: (params as a dict, state) pickled; the GPCache
  entries (device factors) are not part of a checkpoint."""
  if not isinstance(params, dict):
    params = dict(params.__dict__)
  params = dict(params)
  params["cache"] = {}
  save_to_file(filenm, (_portable(params), _portable(state) if state else state))


def load_params(filenm: str, use_gpparams: bool = True,
                include_state: bool = False):
  """params_utils.py:76-87."""
  params_dict, state = load_from_file(filenm)
  params = GPParams(**params_dict) if use_gpparams else params_dict
  if include_state:
    return params, state
  return params


def log_params_loss(step: int, params: GPParams, loss: float,
                    warp_func: Optional[Dict[str, Callable[[Any], Any]]] = None,
                    params_save_file: Optional[str] = None):
  """Log (and optionally checkpoint) the parameters and the loss
  (params_utils.py:193-207)."""
  keys = list(params.model.keys())
  retrieved = dict(zip(keys, retrieve_params(params, keys, warp_func=warp_func)))
  logging.info(msg=f"logging iter={step}, loss={loss}, "
               f"params.model after warping={retrieved}")
  if params_save_file is not None:
    logging.info(msg=f"Saving params to {params_save_file}.")
    save_params(params_save_file, params, state=(step, loss))


# ---------------------------------------------------------------- packing ---
def _to_np(v) -> np.ndarray:
  if isinstance(v, torch.Tensor):
    return v.detach().cpu().numpy().astype(np.float64)
  return np.asarray(v, dtype=np.float64)


def pack_raw(model: Dict[str, Any], d: int, need_mean: bool, warp_func,
             need_noise: bool = True) -> Tuple[np.ndarray, int, bool]:
  """model dict -> (raw[3+d], warp_mask, scalar_lengthscale)."""
  keys = ["lengthscale", "signal_variance"]
  if need_noise:
    keys = keys + ["noise_variance"]
  if need_mean:
    keys = ["constant"] + keys
  _verify_params(model, keys)
  ls = _to_np(model["lengthscale"]).reshape(-1)
  scalar_ls = ls.size == 1
  if not scalar_ls and ls.size != d:
    raise ValueError(f"lengthscale has {ls.size} entries but inputs have d={d}")
  raw = np.empty(3 + d, dtype=np.float64)
  raw[0] = float(_to_np(model["constant"]).reshape(-1)[0]) if need_mean else 0.0
  raw[1] = float(_to_np(model["signal_variance"]).reshape(-1)[0])
  raw[2] = float(_to_np(model["noise_variance"]).reshape(-1)[0]) \
      if "noise_variance" in model else 0.0
  raw[3:] = ls if not scalar_ls else ls[0]
  mask = 0
  if need_mean and utils.warp_kind(warp_func, "constant") == "softplus_eps":
    mask |= 1
  if utils.warp_kind(warp_func, "signal_variance") == "softplus_eps":
    mask |= 2
  if "noise_variance" in model and utils.warp_kind(
      warp_func, "noise_variance") == "softplus_eps":
    mask |= 4
  if utils.warp_kind(warp_func, "lengthscale") == "softplus_eps":
    mask |= ((1 << d) - 1) << 3
  return raw, mask, scalar_ls


def unpack_like(model: Dict[str, Any], vec, d: int, need_mean: bool,
                is_grad: bool = False) -> Dict[str, Any]:
  """raw / gradient vector [3+d] -> dict shaped like `model` (same keys; a
  scalar lengthscale receives the summed ARD gradient).  Keys the engine does
  not own are passed through (zeros for gradients)."""
  vec = _to_np(vec)
  out = {}
  for k, v in model.items():
    tmpl = _to_np(v)
    if k == "constant" and need_mean:
      val = np.asarray(vec[0]).reshape(tmpl.shape)
    elif k == "signal_variance":
      val = np.asarray(vec[1]).reshape(tmpl.shape)
    elif k == "noise_variance":
      val = np.asarray(vec[2]).reshape(tmpl.shape)
    elif k == "lengthscale":
      if tmpl.size == 1:
        val = np.asarray(vec[3:].sum() if is_grad else vec[3]).reshape(tmpl.shape)
      else:
        val = vec[3:].reshape(tmpl.shape)
    else:
      val = np.zeros_like(tmpl) if is_grad else tmpl
    out[k] = float(val) if val.ndim == 0 else val
  return out
