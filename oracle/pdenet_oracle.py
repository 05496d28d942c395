"""CPU oracle for the PDE-Net / PDE-Net 2.0 time-stepping hot path.

*** TEST INFRASTRUCTURE ONLY -- PARITY UNPINNED ***

This file is the checker, never the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it.  Nothing under ``pde-net_b200/`` imports it.

Provenance.  The mounted reference (``/root/reference``) is the upstream
``master`` branch: ``README.md`` (31 lines), ``.gitignore`` (1 line) and
``figures/pdenet.jpg``.  It contains no source (README.md:8 points at the absent
``PDE-Net`` / ``PDE-Net-2.0`` branches) and the numerical library it relies on,
``ZichaoLong/aTEAM``, is un-vendored and un-pinned (``.gitignore:1``).  There are no
reference tests, fixtures or golden vectors.  This oracle therefore *restates* the
algorithm from the mathematics of the two papers named at README.md:5-6
(arXiv 1710.09668, arXiv 1812.04426) in the op sequence the upstream code is
built from -- wrap-around pad -> ``F.conv2d`` -> ``F.linear``/mul/cat -> Euler,
with autograd for the reverse pass -- and is pinned by the known-answer tests in
``tests/test_oracle_kat.py``.  Parity with the
upstream code itself is **unpinned**.

Each function cites the SURVEY.md section-8(a) row it restates; the upstream
symbol names in those rows are recollections, not verifiable file:line anchors.

Conventions fixed here (and mirrored by the CUDA path):
  * state ``u``: ``[B, C, H, W]``, axis 0 of a filter acts along H, axis 1 along W;
  * derivative order ``(o0, o1)``: ``o0`` along H (axis 0), ``o1`` along W (axis 1);
  * filter bank order: total order ascending, then ``o0`` ascending:
    ``(0,0), (0,1), (1,0), (0,2), (1,1), (2,0), ...``;
  * feature index into SymNet: ``f = c * N + n`` (channel-major);
  * SymNet flat layout per output channel: for i < L: ``W_i[2, m+i]`` row-major,
    ``b_i[2]``; then ``w_f[m+L]``, ``b_f[1]``.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import torch
import torch.nn.functional as F

__all__ = [
    "order_list", "moment_axis_matrix", "m2k", "k2m", "initial_moments", "pinned_mask",
    "filter_scales", "scaled_kernels", "periodic_pad", "fd2d_apply", "symnet_param_count",
    "symnet_unpack", "symnet_apply", "right_hand_items", "step", "rollout",
    "linear_step", "linear_rollout", "burgers_symnet", "stable_burgers_model",
    "upwind_twins", "upwind_right_hand_items", "upwind_step", "upwind_rollout",
]


# ----------------------------------------------------------------------------- a1
def moment_axis_matrix(k: int, dtype=torch.float64) -> torch.Tensor:
    """SURVEY 8(a) row a1 (upstream ``MK.py::_MK.__init__``).

    ``A[i, j] = (j - (k-1)//2)**i / i!`` for one axis of length ``k``.
    """
    c = (k - 1) // 2
    A = torch.zeros(k, k, dtype=torch.float64)
    for i in range(k):
        for j in range(k):
            A[i, j] = float((j - c) ** i) / math.factorial(i)
    return A.to(dtype)


# ----------------------------------------------------------------------------- a2
def m2k(moment: torch.Tensor) -> torch.Tensor:
    """Row a2 (``M2K.forward``): ``K = A0^-1 . M . A1^-T`` over leading dims."""
    k0, k1 = moment.shape[-2:]
    P0 = torch.linalg.inv(moment_axis_matrix(k0)).to(moment.dtype)
    P1 = torch.linalg.inv(moment_axis_matrix(k1)).to(moment.dtype)
    return P0 @ moment @ P1.transpose(0, 1)


# ----------------------------------------------------------------------------- a3
def k2m(kernel: torch.Tensor) -> torch.Tensor:
    """Row a3 (``K2M.forward``): ``M = A0 . K . A1^T``; exact inverse of :func:`m2k`."""
    k0, k1 = kernel.shape[-2:]
    A0 = moment_axis_matrix(k0, kernel.dtype)
    A1 = moment_axis_matrix(k1, kernel.dtype)
    return A0 @ kernel @ A1.transpose(0, 1)


# ----------------------------------------------------------------------------- a4
def order_list(max_order: int) -> List[Tuple[int, int]]:
    """Filter bank order: total order ascending, then order along H ascending."""
    return [(j, o - j) for o in range(max_order + 1) for j in range(o + 1)]


def initial_moments(k: int, orders: Sequence[Tuple[int, int]], dtype=torch.float64) -> torch.Tensor:
    """Row a4: one-hot moment matrix at each filter's order."""
    M = torch.zeros(len(orders), k, k, dtype=dtype)
    for n, (o0, o1) in enumerate(orders):
        M[n, o0, o1] = 1
    return M


def pinned_mask(k: int, order: Tuple[int, int], constraint="moment") -> torch.Tensor:
    """Row a4 sum-rule constraint.  ``True`` where the moment entry is pinned.

    ``'free'`` (or 0) pins nothing; ``'moment'`` is accuracy 1; an int ``acc`` pins all
    entries of total order ``i + j < |order| + acc``; ``'frozen'`` pins everything.
    """
    mask = torch.zeros(k, k, dtype=torch.bool)
    if constraint == "free" or constraint == 0:
        return mask
    if constraint == "frozen":
        return ~mask
    acc = 1 if constraint == "moment" else int(constraint)
    s = order[0] + order[1] + acc
    for i in range(k):
        for j in range(k):
            if i + j < s:
                mask[i, j] = True
    return mask


def filter_scales(orders: Sequence[Tuple[int, int]], dx, dtype=torch.float64) -> torch.Tensor:
    """Row a4: ``scale_n = 1 / (dx0**o0 * dx1**o1)``."""
    dx0, dx1 = (dx, dx) if not isinstance(dx, (tuple, list)) else dx
    return torch.tensor([1.0 / (dx0 ** o0 * dx1 ** o1) for (o0, o1) in orders], dtype=dtype)


def scaled_kernels(moments: torch.Tensor, scales: torch.Tensor) -> torch.Tensor:
    """``Ks[n] = scale_n * M2K(moment_n)`` -- what the stencil actually applies."""
    return m2k(moments) * scales.to(moments.dtype)[:, None, None]


# ----------------------------------------------------------------------------- a5
def periodic_pad(u: torch.Tensor, k: int) -> torch.Tensor:
    """Row a5 (``periodicpad``): wrap-around pad ``(r, k-1-r)`` on both spatial axes."""
    r0, r1 = (k - 1) // 2, k - 1 - (k - 1) // 2
    H, W = u.shape[-2:]
    iy = torch.arange(-r0, H + r1) % H
    ix = torch.arange(-r0, W + r1) % W
    return u[..., iy, :][..., ix]


# ----------------------------------------------------------------------------- a6
def fd2d_apply(u: torch.Tensor, Ks: torch.Tensor) -> torch.Tensor:
    """Row a6 (``_FDNd.forward``): cross-correlate every channel with every filter.

    ``u [B,C,H,W]``, ``Ks [N,k,k]`` -> ``[B, C*N, H, W]`` with feature ``c*N + n``.
    PyTorch ``conv2d`` semantics: no kernel flip.
    """
    B, C, H, W = u.shape
    N, k, _ = Ks.shape
    up = periodic_pad(u.reshape(B * C, 1, H, W), k)
    d = F.conv2d(up, Ks[:, None])
    return d.reshape(B, C * N, H, W)


# ----------------------------------------------------------------------------- a7
def symnet_param_count(m: int, L: int) -> int:
    return sum(2 * (m + i) + 2 for i in range(L)) + (m + L) + 1


def symnet_unpack(flat: torch.Tensor, m: int, L: int):
    """Split one channel's flat parameter vector into ``[(W_i, b_i)...], (w_f, b_f)``."""
    out, o = [], 0
    for i in range(L):
        W = flat[o:o + 2 * (m + i)].reshape(2, m + i); o += 2 * (m + i)
        b = flat[o:o + 2]; o += 2
        out.append((W, b))
    wf = flat[o:o + m + L]; o += m + L
    bf = flat[o:o + 1]; o += 1
    assert o == flat.numel()
    return out, (wf, bf)


def symnet_apply(feat: torch.Tensor, flat: torch.Tensor, L: int) -> torch.Tensor:
    """Row a7 (``expr.poly.forward``): ``feat [..., m]`` -> ``[...]``."""
    m = feat.shape[-1]
    hidden, (wf, bf) = symnet_unpack(flat, m, L)
    h = feat
    for (W, b) in hidden:
        o = F.linear(h, W, b)
        h = torch.cat([h, o[..., :1] * o[..., 1:]], dim=-1)
    return F.linear(h, wf[None], bf)[..., 0]


# ----------------------------------------------------------------------------- a8
def right_hand_items(u: torch.Tensor, Ks: torch.Tensor, symnet: torch.Tensor, L: int) -> torch.Tensor:
    """Row a8: ``F[b,c,y,x] = SymNet_c(all derivative features at (y,x))``.

    ``symnet`` is ``[C, P]`` (one flat parameter vector per output channel).
    """
    d = fd2d_apply(u, Ks).permute(0, 2, 3, 1)  # channels-last features
    return torch.stack([symnet_apply(d, symnet[c], L) for c in range(u.shape[1])], dim=1)


# ----------------------------------------------------------------------------- a9
def step(u, Ks, symnet, L, dt):
    """Row a9 (``predict``): forward Euler ``u + dt * F(u)``."""
    return u + dt * right_hand_items(u, Ks, symnet, L)


# ----------------------------------------------------------------------------- a10
def rollout(u0, Ks, symnet, L, dt, steps: int):
    """Row a10 (``multistep``): list ``[u_1, ..., u_steps]``."""
    states, u = [], u0
    for _ in range(steps):
        u = step(u, Ks, symnet, L, dt)
        states.append(u)
    return states


# ----------------------------------------------------------------------------- f-1
def upwind_twins(Ks: torch.Tensor) -> torch.Tensor:
    """Row f-1.  The flipped twins of the two first-order filters of a bank in bank order (index 1 =
    order (0,1), derivative along W; index 2 = order (1,0), along H): mirror along the derivative axis
    and negate, ``K1-[a,b] = -K1[a,k-1-b]``, ``K2-[a,b] = -K2[k-1-a,b]`` -- the same derivative with
    the opposite one-sided bias (a centred filter is its own twin).  Returns ``[2, k, k]``."""
    return torch.stack([-torch.flip(Ks[1], dims=(1,)), -torch.flip(Ks[2], dims=(0,))])


def upwind_right_hand_items(u: torch.Tensor, Ks: torch.Tensor, symnet: torch.Tensor, L: int) -> torch.Tensor:
    """Row f-1: ``RightHandItems`` under ``scheme='upwind'`` (pseudo-upwind).

    [PAPER, UNPINNED]  Restated from the description of the pseudo-upwind scheme in PDE-Net 2.0
    (arXiv 1812.04426, README.md:6): the first-order derivative inputs of the SymNet are taken from a
    one-sided ("upwind") version of their filter, the side being chosen point by point from the sign of
    the local convection velocity, which the network itself supplies as the derivative of its output
    with respect to that input.  The upstream implementation (branch ``PDE-Net-2.0``, ``polypde.py``)
    is not in the mount; the conventions below are this repository's explicit choices:

    * for output channel ``co`` and every first-order input ``f = c*N + n`` (``n = 1``: d/dx along W,
      ``n = 2``: d/dy along H) the velocity is ``a = d SymNet_co / d x_f`` evaluated at the CENTRAL
      features (all filters as given);
    * ``a >= 0``: the input stays ``K_n * u_c``;  ``a < 0``: it becomes ``K_n- * u_c`` with the flipped
      twin of :func:`upwind_twins`;
    * every SymNet sees its own selection; the selection is piecewise constant and is treated as a
      constant by the reverse pass (``detach``).
    """
    B, C, H, W = u.shape
    N = Ks.shape[0]
    d = fd2d_apply(u, Ks).permute(0, 2, 3, 1)                       # [B,H,W,C*N] central features
    dt = fd2d_apply(u, upwind_twins(Ks)).permute(0, 2, 3, 1)        # [B,H,W,C*2] twins, index 2c + dir
    out = []
    for co in range(C):
        with torch.enable_grad():
            x = d.detach().requires_grad_(True)
            a, = torch.autograd.grad(symnet_apply(x, symnet[co].detach(), L).sum(), x)
        cols = []
        for f in range(C * N):
            c, n = divmod(f, N)
            if n in (1, 2):
                cols.append(torch.where(a[..., f] < 0, dt[..., 2 * c + (n - 1)], d[..., f]))
            else:
                cols.append(d[..., f])
        out.append(symnet_apply(torch.stack(cols, dim=-1), symnet[co], L))
    return torch.stack(out, dim=1)


def upwind_step(u, Ks, symnet, L, dt):
    return u + dt * upwind_right_hand_items(u, Ks, symnet, L)


def upwind_rollout(u0, Ks, symnet, L, dt, steps: int):
    states, u = [], u0
    for _ in range(steps):
        u = upwind_step(u, Ks, symnet, L, dt)
        states.append(u)
    return states


# ----------------------------------------------------------------------------- synthetic models
def burgers_symnet(nu: float, N: int = 6, C: int = 2, L: int = 2) -> torch.Tensor:
    """Hand-set SymNet weights (KAT-4): channel c -> ``-u w_x - v w_y + nu (w_xx + w_yy)``, w = u or v.

    Feature index f = c*N + n with orders (0,0),(0,1),(1,0),(0,2),(1,1),(2,0);
    (0,1) is d/dx (axis W), (1,0) is d/dy (axis H).  Two product units suffice (L = 2).
    """
    assert N >= 6 and C == 2 and L == 2
    m = C * N
    sym = torch.zeros(C, symnet_param_count(m, L), dtype=torch.float64)
    iu, iv = 0, N
    for c in range(C):
        base = c * N
        W0 = torch.zeros(2, m, dtype=torch.float64)
        W0[0, iu] = 1; W0[1, base + 1] = 1        # p0 = u,  q0 = w_x
        W1 = torch.zeros(2, m + 1, dtype=torch.float64)
        W1[0, iv] = 1; W1[1, base + 2] = 1        # p1 = v,  q1 = w_y
        wf = torch.zeros(m + L, dtype=torch.float64)
        wf[m] = -1; wf[m + 1] = -1
        wf[base + 3] = nu; wf[base + 5] = nu
        z2 = torch.zeros(2, dtype=torch.float64)
        sym[c] = torch.cat([W0.flatten(), z2, W1.flatten(), z2, wf, torch.zeros(1, dtype=torch.float64)])
    return sym


def stable_burgers_model(k: int, dx: float, nu: float = 0.05, noise: float = 1e-2, moment_noise: float = 1e-3,
                         seed: int = 0):
    """Synthetic model for parity / benchmark runs on fine grids: viscous Burgers (KAT-4 weights) with
    ``noise * N(0,1)`` added to EVERY SymNet entry and ``moment_noise * N(0,1)`` to every non-pinned moment
    entry (1e-2 there puts an anti-diffusive grid-scale response of ~0.85 per step into the 1/dx^2-scaled
    filters at 128^2, dt = 1e-3, which single precision does not survive for 15 steps), so that all
    parameter gradients are non-trivial while forward Euler stays stable with margin at ``dt = dx^2``,
    ``nu = 0.05`` (diffusion number ``nu dt ||K_xx||_1 = 0.05 * 6.04 = 0.3`` per direction).  (A fully random SymNet -- N(0, 0.1^2) on the
    second-derivative inputs -- is anti-diffusive half of the time and blows up within a few steps on
    128^2 and finer grids.)  Returns ``(orders, moments [6,k,k], scales [6], symnet [2,69])``, all fp64."""
    orders = order_list(2)
    g = torch.Generator().manual_seed(seed)
    M = initial_moments(k, orders)
    for n, o in enumerate(orders):
        free = ~pinned_mask(k, o, "moment")
        M[n] += moment_noise * torch.randn(k, k, generator=g, dtype=torch.float64) * free
    sym = burgers_symnet(nu) + noise * torch.randn(2, symnet_param_count(12, 2), generator=g, dtype=torch.float64)
    return orders, M, filter_scales(orders, dx), sym


# ----------------------------------------------------------------------------- a12
def linear_step(u, K0, Ks, coe, dt):
    """Row a12 (PDE-Net 1 ``VariantCoeLinear2d``): ``D0*u + dt * sum_n c_n . (D_n*u)``.

    ``u [B,H,W]``, ``K0 [k,k]`` (learnable order-(0,0) filter), ``Ks [N,k,k]``,
    ``coe [N,H,W]``.
    """
    B, H, W = u.shape
    allK = torch.cat([K0[None], Ks], dim=0)
    d = fd2d_apply(u[:, None], allK)  # [B, 1+N, H, W]
    return d[:, 0] + dt * (coe[None] * d[:, 1:]).sum(dim=1)


def linear_rollout(u0, K0, Ks, coe, dt, steps: int):
    states, u = [], u0
    for _ in range(steps):
        u = linear_step(u, K0, Ks, coe, dt)
        states.append(u)
    return states


# ----------------------------------------------------------------------------- f-3
def huber_l1(p: torch.Tensor, threshold: float) -> torch.Tensor:
    """Huber-smoothed L1 norm used by the training loss's sparsity terms (upstream
    ``setups.py`` [RECALL]): ``|p| - threshold/2`` outside ``[-threshold, threshold]``,
    ``p^2 / (2 threshold)`` inside."""
    a = p.abs()
    return torch.where(a >= threshold, a - threshold / 2, p * p / (2 * threshold)).sum()


def training_loss(u0, moments, scales, pinned, symnet, L, dt, observed, block_weights=None, sparsity=0.0,
                  momentsparsity=0.0, symnet_threshold=1e-3, moment_threshold=1e-2):
    """Loss of one optimiser evaluation (upstream ``train.py`` closure [RECALL]):
    ``sum_t w_t mean((u_t - obs_t)^2) + blocks * (sparsity * H(symnet) + momentsparsity * H(free moments))``."""
    blocks = observed.shape[0]
    states = torch.stack(rollout(u0, scaled_kernels(moments, scales), symnet, L, dt, blocks))
    w = [1.0] * blocks if block_weights is None else list(block_weights)
    loss = sum(w[t] * ((states[t] - observed[t]) ** 2).mean() for t in range(blocks))
    loss = loss + blocks * sparsity * huber_l1(symnet, symnet_threshold)
    loss = loss + blocks * momentsparsity * huber_l1(moments[~pinned], moment_threshold)
    return loss


# ----------------------------------------------------------------------------- f-2
def lagrange1d(x: torch.Tensor, coef: torch.Tensor, lo: float, hi: float, mesh: int, degree: int) -> torch.Tensor:
    """Piecewise Lagrange interpolant of one variable (upstream ``aTEAM`` ``LagrangeInterp``, m = 1
    [RECALL]): ``mesh`` uniform cells on [lo, hi], ``degree + 1`` equispaced nodes per cell sharing
    the end nodes, ``coef[mesh*degree + 1]`` the node values; outside [lo, hi] the first / last
    cell's polynomial is continued."""
    s = (x - lo) * (mesh / (hi - lo))
    c = torch.clamp(torch.floor(s.detach()), 0, mesh - 1).long()
    t = s - c.to(s.dtype)
    y = torch.zeros_like(x)
    for j in range(degree + 1):
        basis = torch.ones_like(x)
        for m in range(degree + 1):
            if m != j:
                basis = basis * (t - m / degree) / ((j - m) / degree)
        y = y + coef[c * degree + j] * basis
    return y


def single_nonlinear_step(u, K0, Ks, coe, fcoef, lo, hi, mesh, degree, dt):
    """Row f-2 (PDE-Net 1 ``SingleNonLinear2d`` [RECALL]): ``D0*u + dt*(sum_n c_n (D_n*u) + f(u))``
    with constant coefficients ``coe[N]`` and the interpolated source term ``f``."""
    B, H, W = u.shape
    fields = coe[:, None, None].expand(-1, H, W)
    return linear_step(u, K0, Ks, fields, dt) + dt * lagrange1d(u, fcoef, lo, hi, mesh, degree)
