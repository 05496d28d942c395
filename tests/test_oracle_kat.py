"""Known-answer tests that pin the oracle.

The reference mount has no tests or golden vectors, so these are the pins.
"""
import math

import numpy as np
import pytest
import torch

from oracle import pdenet_oracle as O
from oracle import pdenet_oracle_np as ONP

torch.set_default_dtype(torch.float64)


def onehot(k, i, j):
    M = torch.zeros(k, k)
    M[i, j] = 1
    return M


# KAT-1 ------------------------------------------------------------------
@pytest.mark.parametrize("k,row", [
    (3, [-1 / 2, 0, 1 / 2]),
    (5, [1 / 12, -2 / 3, 0, 2 / 3, -1 / 12]),
    (7, [-1 / 60, 3 / 20, -3 / 4, 0, 3 / 4, -3 / 20, 1 / 60]),
])
def test_kat1_first_derivative_rows(k, row):
    K = O.m2k(onehot(k, 0, 1))
    c = (k - 1) // 2
    assert torch.allclose(K[c], torch.tensor(row), atol=1e-13)
    # all other rows vanish: the (0,1) filter is a pure d/dx1 stencil
    K[c] = 0
    assert K.abs().max() < 1e-13
    # and the transposed one-hot differentiates along axis 0
    Kt = O.m2k(onehot(k, 1, 0))
    assert torch.allclose(Kt[:, c], torch.tensor(row), atol=1e-13)


def test_kat1_second_derivative_k3():
    K = O.m2k(onehot(3, 0, 2))
    assert torch.allclose(K[1], torch.tensor([1.0, -2.0, 1.0]), atol=1e-13)


# KAT-2 ------------------------------------------------------------------
@pytest.mark.parametrize("k,cond", [(3, 4.74), (5, 32.7), (7, 293.8)])
def test_kat2_roundtrip_and_condition(k, cond):
    g = torch.Generator().manual_seed(k)
    M = torch.randn(4, k, k, generator=g)
    assert torch.allclose(O.k2m(O.m2k(M)), M, atol=1e-9)
    assert torch.allclose(O.m2k(O.k2m(M)), M, atol=1e-9)
    A = O.moment_axis_matrix(k)
    assert abs(torch.linalg.cond(A).item() - cond) / cond < 0.01


# KAT-3 ------------------------------------------------------------------
@pytest.mark.parametrize("k,tol", [
    (5, {"01": 5.2e-7, "02": 1.7e-7, "11": 1.8e-5}),
    (7, {"01": 4.4e-10, "02": 1.1e-10, "11": 5.7e-8}),
])
def test_kat3_truncation_error(k, tol):
    n = 100
    dx = 2 * math.pi / n
    y, x = torch.meshgrid(torch.arange(n) * dx, torch.arange(n) * dx, indexing="ij")
    # axis 1 (W) carries sin, axis 0 (H) carries cos 2y
    u = (torch.sin(x) * torch.cos(2 * y))[None, None]
    orders = [(0, 1), (0, 2), (1, 1)]
    Ks = O.scaled_kernels(O.initial_moments(k, orders), O.filter_scales(orders, dx))
    d = O.fd2d_apply(u, Ks)[0]
    exact = [torch.cos(x) * torch.cos(2 * y), -torch.sin(x) * torch.cos(2 * y),
             -2 * torch.cos(x) * torch.sin(2 * y)]
    for name, dn, ex in zip(["01", "02", "11"], d, exact):
        err = (dn - ex).abs().max().item()
        assert err < 1.5 * tol[name], (name, err)
        assert err > 0.3 * tol[name], (name, err)  # the KAT value, not merely a bound


# KAT-4 ------------------------------------------------------------------
burgers_symnet = O.burgers_symnet   # the hand-wired weights live in the oracle (also used by bench.py's parity leg)


def test_kat4_burgers_step_matches_independent_numpy():
    n, dt, nu = 32, 1e-3, 0.05
    dx = 2 * math.pi / n
    g = torch.Generator().manual_seed(0)
    u0 = torch.randn(1, 2, n, n, generator=g)
    orders = O.order_list(2)
    Ks = O.scaled_kernels(O.initial_moments(3, orders), O.filter_scales(orders, dx))
    sym = burgers_symnet(nu)
    out = O.step(u0, Ks, sym, 2, dt)
    un, vn = ONP.burgers_central_step(u0[0, 0].numpy(), u0[0, 1].numpy(), nu, dt, dx)
    assert np.abs(out[0, 0].numpy() - un).max() < 1e-13 * max(1.0, np.abs(un).max()) * 10
    assert np.abs(out[0, 1].numpy() - vn).max() < 1e-13 * max(1.0, np.abs(vn).max()) * 10


# KAT-5 ------------------------------------------------------------------
def random_model(k=5, P=2, C=2, L=2, dx=0.1, seed=0):
    g = torch.Generator().manual_seed(seed)
    orders = O.order_list(P)
    M = O.initial_moments(k, orders) + 1e-2 * torch.randn(len(orders), k, k, generator=g)
    Ks = O.scaled_kernels(M, O.filter_scales(orders, dx))
    sym = 0.1 * torch.randn(C, O.symnet_param_count(C * len(orders), L), generator=g)
    return Ks, sym


def test_kat5_shift_equivariance():
    Ks, sym = random_model()
    g = torch.Generator().manual_seed(5)
    u = torch.randn(2, 2, 12, 10, generator=g)
    a = O.step(torch.roll(u, (3, -2), (2, 3)), Ks, sym, 2, 1e-2)
    b = torch.roll(O.step(u, Ks, sym, 2, 1e-2), (3, -2), (2, 3))
    assert torch.equal(a, b) or (a - b).abs().max() < 1e-14


# KAT-6 ------------------------------------------------------------------
@pytest.mark.parametrize("k", [3, 5])
def test_kat6_gradcheck(k):
    Ks, sym = random_model(k=k, dx=0.5, seed=k)
    g = torch.Generator().manual_seed(6)
    u = torch.randn(1, 2, 8, 8, generator=g, requires_grad=True)
    Ks = Ks.clone().requires_grad_(True)
    sym = sym.clone().requires_grad_(True)

    def f(u, Ks, sym):
        return O.rollout(u, Ks, sym, 2, 1e-2, 2)[-1]

    assert torch.autograd.gradcheck(f, (u, Ks, sym), eps=1e-6, atol=1e-6, rtol=1e-5)


# KAT-7 ------------------------------------------------------------------
@pytest.mark.parametrize("order", [(0, 0), (0, 1), (1, 1), (2, 0)])
def test_kat7_constraint_masks(order):
    k = 7
    s = order[0] + order[1]
    mask = O.pinned_mask(k, order, "moment")
    assert int(mask.sum()) == (s + 1) * (s + 2) // 2
    assert int(O.pinned_mask(k, order, "free").sum()) == 0
    assert int(O.pinned_mask(k, order, "frozen").sum()) == k * k
    assert int(O.pinned_mask(k, order, 2).sum()) == (s + 2) * (s + 3) // 2
    assert mask[order]  # the one-hot entry itself is pinned


# cross-pin: torch autograd oracle == independent numpy adjoint -------------
@pytest.mark.parametrize("k,L,C", [(3, 2, 2), (5, 2, 2), (7, 1, 1), (5, 3, 2)])
def test_torch_oracle_matches_numpy_adjoint(k, L, C):
    P = 2
    Ks, sym = random_model(k=k, P=P, C=C, L=L, dx=0.3, seed=11 + k)
    g = torch.Generator().manual_seed(1)
    u0 = torch.randn(2, C, 9, 11, generator=g)
    steps, dt = 3, 1e-2
    u0t = u0.clone().requires_grad_(True)
    Kst = Ks.clone().requires_grad_(True)
    symt = sym.clone().requires_grad_(True)
    states = O.rollout(u0t, Kst, symt, L, dt, steps)
    w = torch.randn(states[-1].shape, generator=g)
    (states[-1] * w).sum().backward()
    st_np, lam0, dKs, dsym = ONP.rollout_fwd_bwd(u0.numpy(), Ks.numpy(), sym.numpy(), L, dt, steps, w.numpy())

    def rel(a, b):
        return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)

    for t in range(steps):
        assert rel(st_np[t], states[t].detach().numpy()) < 1e-12
    assert rel(lam0, u0t.grad.numpy()) < 1e-11
    assert rel(dKs, Kst.grad.numpy()) < 1e-11
    assert rel(dsym, symt.grad.numpy()) < 1e-11


def test_linear_step_matches_manual():
    k, N, H, W = 5, 15, 10, 12
    g = torch.Generator().manual_seed(3)
    orders = O.order_list(4)
    assert len(orders) == N
    M = O.initial_moments(k, orders) + 1e-2 * torch.randn(N, k, k, generator=g)
    Ks = O.scaled_kernels(M, O.filter_scales(orders, 0.2))
    K0 = O.m2k(onehot(k, 0, 0) + 1e-2 * torch.randn(k, k, generator=g))
    coe = torch.randn(N, H, W, generator=g)
    u = torch.randn(3, H, W, generator=g)
    out = O.linear_step(u, K0, Ks, coe, 1.5e-2)
    d = ONP.fd2d_apply(u[:, None].numpy(), torch.cat([K0[None], Ks]).numpy())
    ref = d[:, 0] + 1.5e-2 * (coe.numpy()[None] * d[:, 1:]).sum(1)
    assert np.abs(out.numpy() - ref).max() < 1e-11


# ----------------------------------------------------------------------------- f-2 (KAT-8)
@pytest.mark.parametrize("degree", [1, 2, 3, 4])
def test_kat8_lagrange1d_reproduces_polynomials_and_nodes(degree):
    """The piecewise interpolant of degree d is exact for any polynomial of degree <= d (also when
    continued outside the mesh), hits its node values, and is continuous across cells."""
    mesh, lo, hi = 5, -1.0, 2.0
    nodes = torch.linspace(lo, hi, mesh * degree + 1, dtype=torch.float64)

    def f(x):
        return 0.3 * x ** degree - x + 0.5

    x = torch.linspace(-1.5, 2.5, 401, dtype=torch.float64)
    assert (O.lagrange1d(x, f(nodes), lo, hi, mesh, degree) - f(x)).abs().max() < 1e-12
    g = torch.Generator().manual_seed(degree)
    coef = torch.randn(mesh * degree + 1, generator=g, dtype=torch.float64)
    assert (O.lagrange1d(nodes, coef, lo, hi, mesh, degree) - coef).abs().max() < 1e-13
    edges = torch.linspace(lo, hi, mesh + 1, dtype=torch.float64)[1:-1]
    left = O.lagrange1d(edges - 1e-9, coef, lo, hi, mesh, degree)
    right = O.lagrange1d(edges + 1e-9, coef, lo, hi, mesh, degree)
    assert (left - right).abs().max() < 1e-6


# ----------------------------------------------------------------------------- property tests
# (hypothesis; small random configurations -- the size-independent properties the domain offers)
from hypothesis import given, settings, strategies as st  # noqa: E402


@settings(max_examples=25, deadline=None, derandomize=True)
@given(k=st.sampled_from([3, 5, 7]), max_order=st.integers(1, 2), C=st.integers(1, 2), L=st.integers(0, 3),
       H=st.integers(1, 12), W=st.integers(1, 12), sy=st.integers(-15, 15), sx=st.integers(-15, 15),
       seed=st.integers(0, 10 ** 6))
def test_property_shift_equivariance_any_shape(k, max_order, C, L, H, W, sy, sx, seed):
    """Periodic boundaries: rolling the input rolls the output, for any grid (also grids smaller
    than the stencil, where taps wrap several times), any depth, any shift."""
    g = torch.Generator().manual_seed(seed)
    orders = O.order_list(max_order)
    N = len(orders)
    Ks = torch.randn(N, k, k, generator=g, dtype=torch.float64)
    sym = 0.3 * torch.randn(C, O.symnet_param_count(C * N, L), generator=g, dtype=torch.float64)
    u = torch.randn(2, C, H, W, generator=g, dtype=torch.float64)
    a = O.rollout(torch.roll(u, (sy, sx), (2, 3)), Ks, sym, L, 1e-2, 2)[-1]
    b = torch.roll(O.rollout(u, Ks, sym, L, 1e-2, 2)[-1], (sy, sx), (2, 3))
    assert torch.allclose(a, b, rtol=1e-12, atol=1e-12)


@settings(max_examples=25, deadline=None, derandomize=True)
@given(k=st.sampled_from([3, 5]), N=st.integers(1, 5), H=st.integers(2, 10), W=st.integers(2, 10),
       alpha=st.floats(-2, 2), beta=st.floats(-2, 2), seed=st.integers(0, 10 ** 6))
def test_property_linear_model_superposition(k, N, H, W, alpha, beta, seed):
    """Row a12 is linear in the state: rollout(alpha u + beta v) = alpha rollout(u) + beta rollout(v)."""
    g = torch.Generator().manual_seed(seed)
    K0 = torch.randn(k, k, generator=g, dtype=torch.float64) * 0.1
    Ks = torch.randn(N, k, k, generator=g, dtype=torch.float64)
    coe = torch.randn(N, H, W, generator=g, dtype=torch.float64)
    u, v = (torch.randn(2, H, W, generator=g, dtype=torch.float64) for _ in range(2))
    lhs = O.linear_rollout(alpha * u + beta * v, K0, Ks, coe, 1e-2, 3)[-1]
    rhs = alpha * O.linear_rollout(u, K0, Ks, coe, 1e-2, 3)[-1] + beta * O.linear_rollout(v, K0, Ks, coe, 1e-2, 3)[-1]
    assert torch.allclose(lhs, rhs, rtol=1e-11, atol=1e-11)


@settings(max_examples=20, deadline=None, derandomize=True)
@given(k=st.sampled_from([3, 5, 7]), seed=st.integers(0, 10 ** 6))
def test_property_m2k_k2m_inverse_and_linear(k, seed):
    g = torch.Generator().manual_seed(seed)
    M1, M2 = (torch.randn(3, k, k, generator=g, dtype=torch.float64) for _ in range(2))
    assert torch.allclose(O.k2m(O.m2k(M1)), M1, rtol=1e-8, atol=1e-9)
    assert torch.allclose(O.m2k(2.0 * M1 - 0.5 * M2), 2.0 * O.m2k(M1) - 0.5 * O.m2k(M2), rtol=1e-10, atol=1e-10)


@settings(max_examples=15, deadline=None, derandomize=True)
@given(k=st.sampled_from([3, 5]), C=st.integers(1, 2), L=st.integers(0, 2), H=st.integers(3, 9), W=st.integers(3, 9),
       steps=st.integers(1, 3), seed=st.integers(0, 10 ** 6))
def test_property_torch_and_numpy_oracles_agree(k, C, L, H, W, steps, seed):
    """Two independent restatements (autograd vs hand-derived adjoint) give the same states and the
    same input gradient on random small configurations."""
    g = torch.Generator().manual_seed(seed)
    orders = O.order_list(1)
    N = len(orders)
    Ks = torch.randn(N, k, k, generator=g, dtype=torch.float64)
    sym = 0.3 * torch.randn(C, O.symnet_param_count(C * N, L), generator=g, dtype=torch.float64)
    u0 = torch.randn(2, C, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(2, C, H, W, generator=g, dtype=torch.float64)
    uc = u0.clone().requires_grad_(True)
    last = O.rollout(uc, Ks, sym, L, 1e-2, steps)[-1]
    (last * w).sum().backward()
    out = ONP.rollout_fwd_bwd(u0.numpy(), Ks.numpy(), sym.numpy(), L, 1e-2, steps, w.numpy())
    states, gu = out[0], out[1]
    assert np.allclose(states[-1], last.detach().numpy(), rtol=1e-11, atol=1e-12)
    assert np.allclose(gu, uc.grad.numpy(), rtol=1e-9, atol=1e-11)


# ---- row f-1: pseudo-upwind scheme (restated from the paper's description; parity unpinned) ---------------
def _asym_model(k=5, C=2, L=2, dx=0.3, seed=0, noise=5e-2):
    """Bank with visibly one-sided first-order filters (moment noise) and a Burgers-like SymNet + noise."""
    g = torch.Generator().manual_seed(seed)
    orders = O.order_list(2)
    M = O.initial_moments(k, orders) + noise * torch.randn(len(orders), k, k, generator=g, dtype=torch.float64)
    Ks = O.scaled_kernels(M, O.filter_scales(orders, dx))
    sym = (O.burgers_symnet(0.05) if C == 2 else torch.zeros(C, O.symnet_param_count(C * 6, L), dtype=torch.float64))
    sym = sym + 0.1 * torch.randn(C, O.symnet_param_count(C * 6, L), generator=g, dtype=torch.float64)
    return Ks, sym


def test_upwind_twins_and_centred_filters():
    """A centred first-derivative filter is its own twin, so on one-hot moments the upwind scheme IS the central
    scheme; the twin of a one-sided filter is the same derivative with the opposite bias."""
    orders = O.order_list(2)
    Ks = O.scaled_kernels(O.initial_moments(5, orders), O.filter_scales(orders, 0.1))
    tw = O.upwind_twins(Ks)
    assert torch.allclose(tw[0], Ks[1], atol=1e-12) and torch.allclose(tw[1], Ks[2], atol=1e-12)
    _, sym = _asym_model()
    u = torch.randn(2, 2, 12, 14, generator=torch.Generator().manual_seed(1), dtype=torch.float64)
    assert torch.allclose(O.upwind_right_hand_items(u, Ks, sym, 2), O.right_hand_items(u, Ks, sym, 2), atol=1e-10)
    # backward difference (uses x-1, x) <-> forward difference (uses x, x+1): both approximate d/dx
    bwd = torch.zeros(3, 3, dtype=torch.float64)
    bwd[1, 0], bwd[1, 1] = -1.0, 1.0
    bank = torch.stack([torch.zeros(3, 3, dtype=torch.float64), bwd, bwd.t()])
    tw = O.upwind_twins(bank)
    fwd = torch.zeros(3, 3, dtype=torch.float64)
    fwd[1, 1], fwd[1, 2] = -1.0, 1.0
    assert torch.equal(tw[0], fwd) and torch.equal(tw[1], fwd.t())


def test_upwind_selects_by_sign_of_velocity():
    """Scalar advection u_t = a u_x with a constant-sign coefficient: the SymNet is linear, its derivative
    w.r.t. u_x is `a` everywhere, so a >= 0 uses the filter as given and a < 0 its twin -- at every point."""
    H = W = 16
    dx = 2 * math.pi / W
    bwd = torch.zeros(3, 3, dtype=torch.float64)
    bwd[1, 0], bwd[1, 1] = -1.0 / dx, 1.0 / dx          # one-sided: (u(x) - u(x-dx)) / dx
    ident = torch.zeros(3, 3, dtype=torch.float64)
    ident[1, 1] = 1.0
    Ks = torch.stack([ident, bwd, bwd.t()])             # max_order 1 bank: (0,0), (0,1), (1,0)
    u = torch.sin(torch.arange(W, dtype=torch.float64) * dx)[None, None, None, :].expand(1, 1, H, W).contiguous()
    for a in (0.7, -0.7):
        sym = torch.zeros(1, O.symnet_param_count(3, 0), dtype=torch.float64)
        sym[0, 1] = a                                   # out = a * u_x
        F = O.upwind_right_hand_items(u, Ks, sym, 0)
        ux_bwd = (u - torch.roll(u, 1, 3)) / dx
        ux_fwd = (torch.roll(u, -1, 3) - u) / dx
        assert torch.allclose(F, a * (ux_bwd if a >= 0 else ux_fwd), atol=1e-13)


@pytest.mark.parametrize("k,C,L", [(5, 2, 2), (3, 1, 1), (7, 2, 2), (5, 1, 3)])
def test_upwind_torch_oracle_matches_numpy_adjoint(k, C, L):
    """The torch restatement (autograd through the detached selection) and the independent numpy restatement
    (closed-form velocity, hand-derived adjoint, twins folded back) agree on states and every gradient."""
    g = torch.Generator().manual_seed(3 + k)
    orders = O.order_list(2)
    M = O.initial_moments(k, orders) + 5e-2 * torch.randn(len(orders), k, k, generator=g, dtype=torch.float64)
    Ks = O.scaled_kernels(M, O.filter_scales(orders, 0.3))
    sym = 0.3 * torch.randn(C, O.symnet_param_count(C * 6, L), generator=g, dtype=torch.float64)
    u0 = torch.randn(2, C, 9, 11, generator=g, dtype=torch.float64)
    w = torch.randn(2, C, 9, 11, generator=g, dtype=torch.float64)
    steps, dt = 2, 1e-2
    lv = [t.clone().requires_grad_(True) for t in (u0, Ks, sym)]
    out = O.upwind_rollout(*lv, L, dt, steps)[-1]
    (out * w).sum().backward()
    st, lam, dKs, dsym = ONP.upwind_rollout_fwd_bwd(u0.numpy(), Ks.numpy(), sym.numpy(), L, dt, steps, w.numpy())

    def rel(a, b):
        return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)
    assert rel(out.detach().numpy(), st[-1]) < 1e-13
    assert rel(lv[0].grad.numpy(), lam) < 1e-12
    assert rel(lv[1].grad.numpy(), dKs) < 1e-12
    assert rel(lv[2].grad.numpy(), dsym) < 1e-12
    # and the scheme really differs from the central one for these one-sided filters
    assert (out.detach() - O.rollout(u0, Ks, sym, L, dt, steps)[-1]).abs().max() > 1e-6


def test_linear_model_combined_stencil_and_tap_correlation_identities():
    """The algebra the small-grid CUDA path of the PDE-Net 1 linear model rests on (DESIGN.md section 3.9,
    csrc/linear_cluster_kernels.cuh), checked on the CPU against the oracle's step and its autograd:
    one step is ONE variable-coefficient stencil  u+(p) = sum_ab w_ab(p) u(p + ab - r),
    w_ab(p) = K0[ab] + dt sum_n c_n(p) K_n[ab];  and with the tap correlations
    R_ab(p) = sum_t lam_{t+1}(p) u_t(p + ab - r)  (lam = adjoint of the rollout) every parameter gradient is
    dK0[ab] = sum_p R_ab,  dK_n[ab] = dt sum_p c_n R_ab,  dc_n(p) = dt sum_ab K_n[ab] R_ab(p)."""
    g = torch.Generator().manual_seed(5)
    k, N, B, H, W, steps, dt = 5, 4, 2, 9, 11, 3, 0.013
    r = (k - 1) // 2
    K0 = (torch.eye(k * k, dtype=torch.float64)[k * k // 2].reshape(k, k)
          + 0.01 * torch.randn(k, k, generator=g, dtype=torch.float64)).requires_grad_(True)
    Ks = torch.randn(N, k, k, generator=g, dtype=torch.float64).requires_grad_(True)
    coe = torch.randn(N, H, W, generator=g, dtype=torch.float64).requires_grad_(True)
    u0 = torch.randn(B, H, W, generator=g, dtype=torch.float64).requires_grad_(True)
    wts = torch.randn(steps, B, H, W, generator=g, dtype=torch.float64)
    states = O.linear_rollout(u0, K0, Ks, coe, dt, steps)
    (torch.stack(states) * wts).sum().backward()

    def shifted(u, a, b):  # u(p + (a, b) - r), periodic
        return torch.roll(u, shifts=(r - a, r - b), dims=(-2, -1))

    with torch.no_grad():
        w = K0.reshape(k * k)[:, None, None] + dt * torch.einsum("nhw,ne->ehw", coe, Ks.reshape(N, k * k))  # [kk,H,W]
        # forward: the combined stencil reproduces the oracle's states
        u = u0.detach()
        us = [u]
        for t in range(steps):
            u = sum(w[a * k + b] * shifted(u, a, b) for a in range(k) for b in range(k))
            assert (u - states[t].detach()).abs().max() < 1e-12
            us.append(u)
        # reverse: lam_T = seed; lam_t(q) = seed_t(q) + sum_ab (w_ab lam_{t+1})(q - ab + r)
        lam = wts[steps - 1].clone()
        R = torch.zeros(k * k, B, H, W, dtype=torch.float64)
        for t in range(steps - 1, -1, -1):
            for a in range(k):
                for b in range(k):
                    R[a * k + b] += lam * shifted(us[t], a, b)
            nxt = sum(torch.roll(w[a * k + b] * lam, shifts=(a - r, b - r), dims=(-2, -1))
                      for a in range(k) for b in range(k))
            lam = nxt + (wts[t - 1] if t > 0 else 0)
        assert (lam - u0.grad).abs().max() < 1e-11 * u0.grad.abs().max()
        Rsum = R.sum(dim=1)                                                   # [kk,H,W], summed over the batch
        dK0 = Rsum.sum(dim=(1, 2)).reshape(k, k)
        dKs = dt * torch.einsum("nhw,ehw->ne", coe, Rsum).reshape(N, k, k)
        dcoe = dt * torch.einsum("ne,ehw->nhw", Ks.reshape(N, k * k), Rsum)
    for got, ref in ((dK0, K0.grad), (dKs, Ks.grad), (dcoe, coe.grad)):
        assert (got - ref).abs().max() < 1e-11 * ref.abs().max()
