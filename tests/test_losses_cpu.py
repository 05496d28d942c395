"""Host-side checks of the sparsity terms of the training loss: the product's plain
tensor formulas against the oracle's and against hand values.  No CUDA needed (the fused data term
is covered by the gpu tests)."""
import importlib.util
import os
import sys

import numpy as np
import torch

from oracle import pdenet_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_losses():
    """Import pde-net_b200/losses.py without importing the package's CUDA binding."""
    pkg = os.path.join(ROOT, "pde-net_b200")
    spec = importlib.util.spec_from_file_location("pdenet_b200_losses_only", os.path.join(pkg, "losses.py"),
                                                  submodule_search_locations=None)
    src = open(os.path.join(pkg, "losses.py")).read().replace("from . import ops", "ops = None")
    mod = importlib.util.module_from_spec(spec)
    exec(compile(src, spec.origin, "exec"), mod.__dict__)
    return mod


def test_huber_l1_known_values_and_oracle():
    L = _load_losses()
    p = torch.tensor([0.0, 5e-4, -5e-4, 1e-3, -2e-3, 0.5], dtype=torch.float64)
    thr = 1e-3
    want = 0 + 2 * (5e-4 ** 2 / (2 * thr)) + (1e-3 - thr / 2) + (2e-3 - thr / 2) + (0.5 - thr / 2)
    got = L.huber_l1([p[:3], p[3:]], thr)
    assert abs(got.item() - want) < 1e-15
    assert abs(O.huber_l1(p, thr).item() - want) < 1e-15


def test_huber_l1_gradient_is_continuous_and_bounded():
    L = _load_losses()
    thr = 1e-2
    p = torch.linspace(-3 * thr, 3 * thr, 601, dtype=torch.float64, requires_grad=True)
    L.huber_l1([p], thr).backward()
    g = p.grad.numpy()
    assert np.all(np.abs(g) <= 1 + 1e-12)
    assert np.max(np.abs(np.diff(g))) < 2 * (6 * thr / 600) / thr   # slope p/thr inside, +-1 outside
    inside = np.abs(p.detach().numpy()) < thr
    assert np.allclose(g[inside], p.detach().numpy()[inside] / thr)


def test_moment_sparsity_ignores_pinned_entries():
    L = _load_losses()

    class FD:
        def __init__(self, order):
            g = torch.Generator().manual_seed(sum(order))
            self.moment = torch.randn(5, 5, generator=g, dtype=torch.float64, requires_grad=True)
            self._pinned = O.pinned_mask(5, order, "moment")

    class Model:
        fds = [FD(o) for o in O.order_list(2)]

    m = Model()
    got = L.moment_sparsity(m, 1e-2)
    want = sum(O.huber_l1(fd.moment[~fd._pinned], 1e-2) for fd in m.fds)
    assert abs(got.item() - want.item()) < 1e-13
    got.backward()
    for fd in m.fds:
        assert torch.all(fd.moment.grad[fd._pinned] == 0)
        assert torch.any(fd.moment.grad[~fd._pinned] != 0)
