"""Device-side per-step sub-sampling (hb_subsample; basics/data_utils.py:72-100):
the gathered batch is exactly the keyed permutation's prefix, is a subset without
replacement, leaves small tasks untouched, does not depend on the task sharding,
and infer_parameters trains on it with the step's CUDA graph intact."""
import numpy as np
import pytest
import torch

from oracle import hyperbo_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _setup(ns, d=3):
  from hyperbo_b200.engine import Engine
  eng = Engine.get()
  ds = {t: O.make_task(t, n, d) for t, n in enumerate(ns)}
  return eng, ds, eng.pack([(k, v[0], v[1]) for k, v in ds.items()])


def test_gather_is_the_keyed_permutation_prefix():
  from hyperbo_b200 import _C
  from hyperbo_b200.engine import DeviceSampler
  ns, B, seed = [100, 30, 64, 257], 64, 12345
  eng, ds, src = _setup(ns)
  smp = DeviceSampler(eng, src, B, seed)
  assert smp.dst.offs == [0, 64, 94, 158, 222]
  for step in (0, 1, 7):
    out = smp.sample(step=step)
    x, y = out.x.cpu().numpy(), out.y.cpu().numpy()
    for t, n in enumerate(ns):
      lo, hi = out.offs[t], out.offs[t + 1]
      if n <= B:  # (n == B: a permutation is a no-op for the NLL; rows stay in place)
        idx = np.arange(n)
      else:
        idx = np.array([_C.subsample_perm(i, n, seed, step, t) for i in range(B)])
        assert len(set(idx.tolist())) == B          # without replacement
      assert np.array_equal(x[lo:hi], ds[t][0][idx])
      assert np.array_equal(y[lo:hi], ds[t][1][idx, 0])
  a = smp.sample(step=3).x.clone()
  b = smp.sample(step=4).x.clone()
  assert not torch.equal(a, b)                      # a new sample every step


def test_sample_does_not_depend_on_the_sharding():
  from hyperbo_b200.engine import DeviceSampler
  ns, B = [90, 120, 70, 200, 64, 130], 64
  eng, ds, src = _setup(ns)
  full = DeviceSampler(eng, src, B, 5).sample(step=9)
  fx = full.x.cpu().numpy()
  for rank in range(2):
    mine = [t for t in range(len(ns)) if t % 2 == rank]
    part = eng.pack([(t, ds[t][0], ds[t][1]) for t in mine])
    got = DeviceSampler(eng, part, B, 5, task_ids=mine).sample(step=9).x.cpu().numpy()
    for k, t in enumerate(mine):
      assert np.array_equal(got[64 * k:64 * k + 64], fx[full.offs[t]:full.offs[t + 1]])


def test_infer_parameters_with_batch_size_below_n():
  from hyperbo_b200.basics import definitions as defs
  from hyperbo_b200.engine import DeviceSampler, Engine
  from hyperbo_b200.gp_utils import gp, kernel, mean, objectives, utils
  ns, B, d, steps, lr, seed = [150, 40, 96], 48, 3, 6, 1e-2, 3
  eng, ds, src = _setup(ns, d)
  dataset = {k: defs.SubDataset(*v) for k, v in ds.items()}
  params = defs.GPParams(
      model=dict(O.init_raw_params(d)),
      config={"method": "adam", "learning_rate": lr, "max_training_step": steps,
              "batch_size": B, "objective": objectives.nll})
  res = gp.infer_parameters(mean.constant, kernel.squared_exponential, params, dataset,
                            warp_func=utils.DEFAULT_WARP_FUNC, objective=objectives.nll,
                            key=seed)
  got = H.raw_vec({k: np.asarray(v, dtype=np.float64) for k, v in res.model.items()}, d)
  # the same loop by hand: sample(step) with the step passed from the host
  smp = DeviceSampler(eng, src, B, seed)
  raw0, mask = H.raw_vec(O.init_raw_params(d), d), H.default_mask(d)
  tr = gp.AdamTrainer(eng, 0, 1, raw0, mask, d, lr)
  dst = smp.dst
  dst.sampler = None
  losses = []
  for i in range(steps):
    smp.sample(step=i)
    tr.step(dst)
    losses.append(tr.loss())
  assert all(np.isfinite(losses))
  # the oracle's Adam loop over the same six sampled batches
  model, opt, ref_losses = dict(O.init_raw_params(d)), O.Adam(lr), []
  for i in range(steps):
    smp.sample(step=i)
    x, y = dst.x.cpu().numpy(), dst.y.cpu().numpy()
    batch = {t: (x[dst.offs[t]:dst.offs[t + 1]], y[dst.offs[t]:dst.offs[t + 1], None])
             for t in range(len(ns))}
    v, g = O.nll_value_and_grad("constant", "squared_exponential", model, batch,
                                O.DEFAULT_WARP_FUNC)
    ref_losses.append(v)
    model = opt.update(model, g)
  ref = H.raw_vec(model, d)
  assert H.rel(losses, ref_losses) < 1e-9
  assert H.rel(tr.raw.cpu().numpy(), ref) < 1e-8, "hand-driven loop vs oracle"
  assert H.rel(got, ref) < 1e-8, "infer_parameters vs oracle"
  assert H.rel(got, tr.raw.cpu().numpy()) < 1e-12
