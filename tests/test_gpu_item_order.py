"""The persistent kernel's work queue must be a topological order of the
per-task tile DAG (an item popped before an item it waits for could leave every
resident CTA waiting): checked on the item lists the library builds, for the
many-task and the few-task ordering, ragged batches included."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
DIAG, PANEL, TRTRI, LAUUM, ALPHA = range(5)


@pytest.mark.parametrize("ns", [[512] * 256, [512] * 32, [200] * 300, [130, 64, 1, 700, 65] * 9,
                                [4096] * 3, [64]])
@pytest.mark.parametrize("variant", [0, 1, 2])
def test_item_order_is_topological(ns, variant):
  from hyperbo_b200.engine import Engine
  eng = Engine.get()
  offs = [0] + list(np.cumsum(ns))
  items = np.asarray(eng.h.debug_items([int(o) for o in offs], 4, variant)).reshape(-1, 4)
  pos = {tuple(it): k for k, it in enumerate(items.tolist())}
  assert len(pos) == len(items)                      # every item exactly once
  nb = {t: (n + 63) // 64 for t, n in enumerate(ns)}
  want = sum(b + b * (b - 1) // 2 * (2 if variant >= 1 else 1) +
             ((b * (b + 1) // 2 + 1) if variant == 2 else 0) for b in nb.values())
  assert len(items) == want
  for (t, kind, a, b), k in pos.items():
    deps = []
    if kind == PANEL:
      deps.append((t, DIAG, b, b))
      if b > 0:
        deps.append((t, PANEL, a, b - 1))
    elif kind == DIAG and a > 0:
      deps.append((t, PANEL, a, a - 1))
    elif kind == TRTRI:
      deps += [(t, DIAG, a, a), (t, PANEL, a, b)]
      if a - 1 > b:
        deps.append((t, TRTRI, a - 1, b))
    elif kind == ALPHA:
      deps += [(t, DIAG, nb[t] - 1, nb[t] - 1)] + [(t, TRTRI, nb[t] - 1, c) for c in range(nb[t] - 1)]
    elif kind == LAUUM:
      deps.append((t, ALPHA, 0, 0))
    for dep in deps:
      assert pos[dep] < k, (dep, (t, kind, a, b))
