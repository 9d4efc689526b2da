"""Per-step sub-sampling and dataset logging -- mirrors
hyperbo/basics/data_utils.py:29-100."""
from __future__ import annotations

import logging

import numpy as np
import torch

from hyperbo_b200.basics import definitions as defs

SubDataset = defs.SubDataset


def log_dataset(dataset):
  """Log size, shapes and per-column mean / median / min / max of every
  sub-dataset (data_utils.py:29-69; called by GP.train and the data loaders).
  Empty arrays are reported as nan, non-array fields (aligned) as they are."""

  def stat(f, a):
    if not isinstance(a, (np.ndarray, torch.Tensor)):
      return a
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a
    return float("nan") if a.shape[0] == 0 else f(a)

  logging.info(msg=f"dataset len = {len(dataset)}.")
  for name, f in (("shape", np.shape),
                  ("mean", lambda a: np.mean(a, axis=0)),
                  ("median", lambda a: np.median(a, axis=0)),
                  ("min", lambda a: np.min(a, axis=0)),
                  ("max", lambda a: np.max(a, axis=0))):
    summary = {k: tuple(stat(f, field) for field in tuple(s))
               for k, s in dataset.items()}
    logging.info(msg=f"dataset {name}: {summary}")


def sub_sample_dataset_iterator(key, dataset, batch_size):
  """Iterator that sub-samples every sub-dataset with n >= batch_size to
  batch_size random points per step (data_utils.py:72-100).

  `key` is a torch.Generator or an int seed (jax.random keys do not exist
  here; the stream differs from threefry, the distribution does not).
  """
  if isinstance(key, torch.Generator):
    gen = key
  else:
    gen = torch.Generator(device="cpu")
    gen.manual_seed(int(key) if key is not None else 0)
  while True:
    sub_sampled_dataset = {}
    for i, (sub_dataset_key, sub_dataset) in enumerate(dataset.items()):
      if sub_dataset.x.shape[0] >= batch_size:
        indices = torch.randperm(sub_dataset.x.shape[0], generator=gen)
        idx = indices[:batch_size].to(sub_dataset.x.device) if isinstance(
            sub_dataset.x, torch.Tensor) else indices[:batch_size].numpy()
        new_sub_dataset = SubDataset(
            x=sub_dataset.x[idx, :], y=sub_dataset.y[idx, :],
            aligned=sub_dataset.aligned)
      else:
        new_sub_dataset = sub_dataset
      if isinstance(new_sub_dataset.aligned, str):
        new_sub_dataset = SubDataset(
            x=new_sub_dataset.x, y=new_sub_dataset.y, aligned=i)
      sub_sampled_dataset[sub_dataset_key] = new_sub_dataset
    yield sub_sampled_dataset
