"""Synthetic data generator -- mirrors hyperbo/bo_utils/data.py:720-775
(`random`).  The PD1 / HPO-B dataframe wrangling of that file is out of scope
(the datasets are not part of the reference tree)."""
from __future__ import annotations

import torch

from hyperbo_b200.basics import definitions as defs
from hyperbo_b200.gp_utils import gp

SubDataset = defs.SubDataset


def _gen(key):
  if isinstance(key, torch.Generator):
    return key
  g = torch.Generator()
  g.manual_seed(int(key) if key is not None else 0)
  return g


def random(key, mean_func, cov_func, params, dim, n_observed, n_queries,
           n_func_historical=0, m_points_historical=0, warp_func=None):
  """Random historical data and observed data for the current function
  (data.py:720-775): X ~ U[0,1]^dim, y = one GP draw per function.
  Returns (dataset dict, key of the queried sub-dataset, queried SubDataset)."""
  gen = _gen(key)
  dataset = {}
  for i in range(n_func_historical):
    vx = torch.rand((m_points_historical, dim), generator=gen,
                    dtype=torch.float64)
    vy = gp.sample_from_gp(gen, mean_func, cov_func, params, vx,
                           warp_func=warp_func)
    dataset[i] = SubDataset(x=vx.to(vy.device), y=vy)
  vx = torch.rand((n_observed + n_queries, dim), generator=gen,
                  dtype=torch.float64)
  vy = gp.sample_from_gp(gen, mean_func, cov_func, params, vx,
                         warp_func=warp_func)
  vx = vx.to(vy.device)
  x_queries, x_observed = vx[:n_queries], vx[n_queries:]
  y_queries, y_observed = vy[:n_queries], vy[n_queries:]
  dataset[n_func_historical] = SubDataset(x=x_observed, y=y_observed)
  return dataset, n_func_historical, SubDataset(x=x_queries, y=y_queries)
