"""Data contract of the boundary -- mirrors hyperbo/basics/definitions.py:23-46
(GPCache, SubDataset, GPParams), with torch tensors as the array type."""
from __future__ import annotations

import dataclasses
from typing import Any, Callable, Dict, List, NamedTuple, Optional, Tuple, Union

import torch


@dataclasses.dataclass
class GPCache:
  """Caching intermediate results for GP (definitions.py:23-28).

  `chol` (n,n) lower factor and `kinvy` (n,1) are the reference-visible fields;
  `packed` is the engine's opaque predictor cache (packed L^{-1} tiles + alpha)
  that hb_predict consumes.
  """
  chol: torch.Tensor
  kinvy: torch.Tensor
  needs_update: bool
  packed: Optional[torch.Tensor] = None


class SubDataset(NamedTuple):
  """Sub dataset with x: n x d and y: n x m; d, m>=1 (definitions.py:31-35)."""
  x: Any
  y: Any
  aligned: Optional[Union[int, str, bool, Tuple[str, ...]]] = None


@dataclasses.dataclass
class GPParams:
  """Parameters in a GP (definitions.py:38-46)."""
  config: Dict[str, Any] = dataclasses.field(default_factory=lambda: {})
  model: Dict[str, Any] = dataclasses.field(default_factory=lambda: {})
  cache: Dict[Union[int, str], GPCache] = dataclasses.field(
      default_factory=lambda: {})
  samples: List[Dict[str, Any]] = dataclasses.field(default_factory=lambda: [])


AllowedDatasetTypes = Union[List[Union[Tuple[Any, ...], SubDataset]],
                            Dict[Union[str, int], Union[Tuple[Any, ...],
                                                        SubDataset]]]
WarpFuncType = Optional[Dict[str, Callable[[Any], Any]]]
