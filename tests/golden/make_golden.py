"""Generate the committed golden vectors from the CPU oracle.

    python tests/golden/make_golden.py

The reference mount has no source and no fixtures, so these vectors pin
the ORACLE (against silent drift) and give the CUDA path a second, frozen target; they are not
outputs of the upstream code (parity unpinned).
"""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pdenet_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def rollout_case(name, k, max_order, C, L, B, H, W, steps, dx, dt, seed):
    g = torch.Generator().manual_seed(seed)
    orders = O.order_list(max_order)
    N = len(orders)
    M = O.initial_moments(k, orders) + 1e-2 * torch.randn(N, k, k, generator=g, dtype=torch.float64)
    scales = O.filter_scales(orders, dx)
    sym = 0.1 * torch.randn(C, O.symnet_param_count(C * N, L), generator=g, dtype=torch.float64)
    u0 = torch.randn(B, C, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(steps, B, C, H, W, generator=g, dtype=torch.float64)
    Mc, sc, uc = M.clone().requires_grad_(True), sym.clone().requires_grad_(True), u0.clone().requires_grad_(True)
    st = torch.stack(O.rollout(uc, O.scaled_kernels(Mc, scales), sc, L, dt, steps))
    (st * w).sum().backward()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), k=k, max_order=max_order, C=C, L=L, steps=steps, dx=dx,
                        dt=dt, moments=M.numpy(), scales=scales.numpy(), symnet=sym.numpy(), u0=u0.numpy(),
                        w=w.numpy(), states=st.detach().numpy(), grad_u0=uc.grad.numpy(),
                        grad_moments=Mc.grad.numpy(), grad_symnet=sc.grad.numpy())


def linear_case(name, k, max_order, B, H, W, steps, dx, dt, seed):
    g = torch.Generator().manual_seed(seed)
    orders = [o for o in O.order_list(max_order) if o != (0, 0)]
    N = len(orders)
    M = O.initial_moments(k, orders) + 1e-2 * torch.randn(N, k, k, generator=g, dtype=torch.float64)
    Ks = O.scaled_kernels(M, O.filter_scales(orders, dx))
    K0 = O.m2k(O.initial_moments(k, [(0, 0)]) + 1e-3 * torch.randn(1, k, k, generator=g, dtype=torch.float64))[0]
    coe = torch.randn(N, H, W, generator=g, dtype=torch.float64)
    u0 = torch.randn(B, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(steps, B, H, W, generator=g, dtype=torch.float64)
    lv = [t.clone().requires_grad_(True) for t in (u0, K0, Ks, coe)]
    st = torch.stack(O.linear_rollout(*lv, dt, steps))
    (st * w).sum().backward()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), k=k, steps=steps, dt=dt, u0=u0.numpy(), K0=K0.numpy(),
                        Ks=Ks.numpy(), coe=coe.numpy(), w=w.numpy(), states=st.detach().numpy(),
                        grad_u0=lv[0].grad.numpy(), grad_K0=lv[1].grad.numpy(), grad_Ks=lv[2].grad.numpy(),
                        grad_coe=lv[3].grad.numpy())


def upwind_case(name, k, max_order, C, L, B, H, W, steps, dx, dt, seed):
    """scheme='upwind' (row f-1): one-sided filters (5e-2 moment noise) so that the twins differ from the filters."""
    g = torch.Generator().manual_seed(seed)
    orders = O.order_list(max_order)
    N = len(orders)
    M = O.initial_moments(k, orders) + 5e-2 * torch.randn(N, k, k, generator=g, dtype=torch.float64)
    scales = O.filter_scales(orders, dx)
    sym = 0.3 * torch.randn(C, O.symnet_param_count(C * N, L), generator=g, dtype=torch.float64)
    u0 = torch.randn(B, C, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(steps, B, C, H, W, generator=g, dtype=torch.float64)
    Mc, sc, uc = M.clone().requires_grad_(True), sym.clone().requires_grad_(True), u0.clone().requires_grad_(True)
    st = torch.stack(O.upwind_rollout(uc, O.scaled_kernels(Mc, scales), sc, L, dt, steps))
    (st * w).sum().backward()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), k=k, max_order=max_order, C=C, L=L, steps=steps, dx=dx,
                        dt=dt, moments=M.numpy(), scales=scales.numpy(), symnet=sym.numpy(), u0=u0.numpy(),
                        w=w.numpy(), states=st.detach().numpy(), grad_u0=uc.grad.numpy(),
                        grad_moments=Mc.grad.numpy(), grad_symnet=sc.grad.numpy())


if __name__ == "__main__":
    which = sys.argv[1:] or ["rollout", "linear", "upwind"]
    if "rollout" in which:
        rollout_case("rollout_k7_o2_c2_l2", 7, 2, 2, 2, 2, 18, 21, 3, 2 * math.pi / 21, 1e-2, 101)
        rollout_case("rollout_k5_o2_c1_l3", 5, 2, 1, 3, 2, 9, 14, 2, 0.3, 1e-2, 102)
    if "linear" in which:
        linear_case("linear_k5_o4", 5, 4, 2, 12, 13, 4, 0.2, 1e-3, 103)
    if "upwind" in which:
        upwind_case("upwind_k5_o2_c2_l2", 5, 2, 2, 2, 2, 13, 17, 3, 0.3, 1e-2, 104)
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))
