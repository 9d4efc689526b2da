"""Per-step sub-sampling -- mirrors hyperbo/basics/data_utils.py:72-100."""
from __future__ import annotations

import torch

from hyperbo_b200.basics import definitions as defs

SubDataset = defs.SubDataset


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
