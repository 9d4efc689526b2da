"""Parameter retrieval -- mirrors hyperbo/basics/params_utils.py:90-111, plus the
model-dict <-> raw-vector packing of the C ABI (include/hyperbo_b200.h)."""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Tuple

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
