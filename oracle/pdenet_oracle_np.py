"""Second, independent CPU restatement (numpy, shift-and-add, hand-derived adjoint).

*** TEST INFRASTRUCTURE ONLY -- PARITY UNPINNED (see pdenet_oracle.py header) ***

Where ``pdenet_oracle.py`` restates the hot path with the PyTorch op sequence
(pad -> conv2d -> linear -> autograd), this file restates the same mathematics
 with ``np.roll`` sums and an explicitly derived
reverse pass.  The two share no code; ``tests/test_oracle_kat.py`` requires them to
agree to round-off, which pins the torch oracle's autograd against the adjoint
formulas the CUDA reverse kernels implement:

    g        = J_SymNet^T (dt * lam)                      (per point)
    lam_t    = lam_{t+1} + sum_n corr^T(Ks_n, g_{c,n})     (data gradient)
    dKs_n[a,b] = sum_{b,c,y,x} g_{c,n}(y,x) u_c(y+a-r, x+b-r)
"""
from __future__ import annotations

import numpy as np


def fd2d_apply(u, Ks):
    """d[b, c*N+n, y, x] = sum_{a,b} Ks[n,a,b] * u[b,c,(y+a-r)%H,(x+b-r)%W]."""
    B, C, H, W = u.shape
    N, k, _ = Ks.shape
    r = (k - 1) // 2
    d = np.zeros((B, C, N, H, W), dtype=u.dtype)
    for a in range(k):
        for b in range(k):
            sh = np.roll(u, shift=(-(a - r), -(b - r)), axis=(2, 3))
            d += Ks[None, None, :, a, b, None, None] * sh[:, :, None]
    return d.reshape(B, C * N, H, W)


def _unpack(flat, m, L):
    out, o = [], 0
    for i in range(L):
        Wt = flat[o:o + 2 * (m + i)].reshape(2, m + i); o += 2 * (m + i)
        b = flat[o:o + 2]; o += 2
        out.append((Wt, b))
    wf = flat[o:o + m + L]; o += m + L
    bf = flat[o]; o += 1
    return out, wf, bf


def symnet_fwd(d, flat, L):
    """d [B,m,H,W] -> (out [B,H,W], h list, (p,q) list)."""
    m = d.shape[1]
    hidden, wf, bf = _unpack(flat, m, L)
    h = [d[:, j] for j in range(m)]
    pq = []
    for (Wt, b) in hidden:
        p = sum(Wt[0, j] * h[j] for j in range(len(h))) + b[0]
        q = sum(Wt[1, j] * h[j] for j in range(len(h))) + b[1]
        pq.append((p, q))
        h.append(p * q)
    out = sum(wf[j] * h[j] for j in range(len(h))) + bf
    return out, h, pq


def step_fwd(u, Ks, symnet, L, dt):
    d = fd2d_apply(u, Ks)
    F = np.stack([symnet_fwd(d, symnet[c], L)[0] for c in range(u.shape[1])], axis=1)
    return u + dt * F


def _symnet_adjoint(x, flat, L, Lam):
    """Adjoint of one SymNet evaluated at inputs ``x [B,m,H,W]`` with output seed ``Lam [B,H,W]``:
    returns (list of m input adjoints, flat parameter gradient in the parameter layout)."""
    m = x.shape[1]
    hidden, wf, bf = _unpack(flat, m, L)
    _, h, pq = symnet_fwd(x, flat, L)
    hbar = [Lam * wf[j] for j in range(m + L)]
    gflat = np.zeros_like(flat)
    off_f = sum(2 * (m + i) + 2 for i in range(L))
    for j in range(m + L):
        gflat[off_f + j] = np.sum(Lam * h[j])
    gflat[off_f + m + L] = np.sum(Lam)
    for i in reversed(range(L)):
        Wt, b = hidden[i]
        p, q = pq[i]
        eta = hbar[m + i]
        pbar, qbar = eta * q, eta * p
        o = sum(2 * (m + ii) + 2 for ii in range(i))
        for j in range(m + i):
            gflat[o + j] = np.sum(pbar * h[j])
            gflat[o + (m + i) + j] = np.sum(qbar * h[j])
            hbar[j] = hbar[j] + pbar * Wt[0, j] + qbar * Wt[1, j]
        gflat[o + 2 * (m + i)] = np.sum(pbar)
        gflat[o + 2 * (m + i) + 1] = np.sum(qbar)
    return hbar[:m], gflat


def _bank_adjoint(u, lam, Ks, g):
    """Data and filter gradient of the stencil bank: ``g [B,C*N,H,W]`` are the feature adjoints."""
    B, C, H, W = u.shape
    N, k, _ = Ks.shape
    r = (k - 1) // 2
    g5 = g.reshape(B, C, N, H, W)
    lam_prev = lam.copy()
    dKs = np.zeros_like(Ks)
    for a in range(k):
        for b in range(k):
            # corr^T: g at (y,x) contributes Ks[n,a,b]*g to lam at (y+a-r, x+b-r)
            contrib = np.einsum("n,bcnyx->bcyx", Ks[:, a, b], g5)
            lam_prev += np.roll(contrib, shift=(a - r, b - r), axis=(2, 3))
            sh = np.roll(u, shift=(-(a - r), -(b - r)), axis=(2, 3))
            dKs[:, a, b] = np.einsum("bcnyx,bcyx->n", g5, sh)
    return lam_prev, dKs


def step_bwd(u, lam, Ks, symnet, L, dt):
    """Adjoint of one Euler step.  Returns (lam_prev, dKs, dsymnet)."""
    B, C, H, W = u.shape
    N = Ks.shape[0]
    m = C * N
    d = fd2d_apply(u, Ks)
    g = np.zeros_like(d)
    dsym = np.zeros_like(symnet)
    for c in range(C):
        hbar, dsym[c] = _symnet_adjoint(d, symnet[c], L, dt * lam[:, c])
        for j in range(m):
            g[:, j] += hbar[j]
    lam_prev, dKs = _bank_adjoint(u, lam, Ks, g)
    return lam_prev, dKs, dsym


# ---- row f-1: pseudo-upwind (independent restatement of pdenet_oracle.upwind_right_hand_items) -------
def upwind_bank(Ks):
    """[N+2,k,k]: the bank followed by the flipped twins of filters 1 (along W) and 2 (along H)."""
    return np.concatenate([Ks, -Ks[1:2, :, ::-1], -Ks[2:3, ::-1, :]], axis=0)


def _upwind_inputs(dall, C, N, flat, L):
    """Select the SymNet inputs of ONE output channel.  ``dall [B, C*(N+2), H, W]`` are the features of
    the extended bank (channel-major).  Returns (x [B,C*N,H,W], masks {(c,dir): bool [B,H,W]})."""
    B, _, H, W = dall.shape
    d5 = dall.reshape(B, C, N + 2, H, W)
    d = d5[:, :, :N].reshape(B, C * N, H, W)
    vel, _ = _symnet_adjoint(d, flat, L, np.ones((B, H, W), dtype=dall.dtype))   # d out / d input
    x = d.copy()
    masks = {}
    for c in range(C):
        for dr in range(2):
            mk = vel[c * N + 1 + dr] < 0
            masks[(c, dr)] = mk
            x[:, c * N + 1 + dr] = np.where(mk, d5[:, c, N + dr], d5[:, c, 1 + dr])
    return x, masks


def upwind_step_fwd(u, Ks, symnet, L, dt):
    C, N = u.shape[1], Ks.shape[0]
    dall = fd2d_apply(u, upwind_bank(Ks))
    F = np.stack([symnet_fwd(_upwind_inputs(dall, C, N, symnet[co], L)[0], symnet[co], L)[0] for co in range(C)],
                 axis=1)
    return u + dt * F


def upwind_step_bwd(u, lam, Ks, symnet, L, dt):
    """Adjoint of one pseudo-upwind Euler step (selection held constant): (lam_prev, dKs, dsymnet)."""
    B, C, H, W = u.shape
    N, k, _ = Ks.shape
    bank = upwind_bank(Ks)
    dall = fd2d_apply(u, bank)
    g = np.zeros((B, C, N + 2, H, W), dtype=u.dtype)
    dsym = np.zeros_like(symnet)
    for co in range(C):
        x, masks = _upwind_inputs(dall, C, N, symnet[co], L)
        hbar, dsym[co] = _symnet_adjoint(x, symnet[co], L, dt * lam[:, co])
        for f in range(C * N):
            c, n = divmod(f, N)
            if n in (1, 2):
                mk = masks[(c, n - 1)]
                g[:, c, n] += np.where(mk, 0.0, hbar[f])
                g[:, c, N + n - 1] += np.where(mk, hbar[f], 0.0)
            else:
                g[:, c, n] += hbar[f]
    lam_prev, dbank = _bank_adjoint(u, lam, bank, g.reshape(B, C * (N + 2), H, W))
    dKs = dbank[:N].copy()
    dKs[1] -= dbank[N][:, ::-1]
    dKs[2] -= dbank[N + 1][::-1, :]
    return lam_prev, dKs, dsym


def upwind_rollout_fwd_bwd(u0, Ks, symnet, L, dt, steps, grad_last):
    states = [u0]
    for _ in range(steps):
        states.append(upwind_step_fwd(states[-1], Ks, symnet, L, dt))
    lam = grad_last.copy()
    dKs = np.zeros_like(Ks)
    dsym = np.zeros_like(symnet)
    for t in reversed(range(steps)):
        lam, dk, ds = upwind_step_bwd(states[t], lam, Ks, symnet, L, dt)
        dKs += dk
        dsym += ds
    return states[1:], lam, dKs, dsym


def rollout_fwd_bwd(u0, Ks, symnet, L, dt, steps, grad_last):
    """Full rollout + adjoint with the loss gradient given on the last state only."""
    states = [u0]
    for _ in range(steps):
        states.append(step_fwd(states[-1], Ks, symnet, L, dt))
    lam = grad_last.copy()
    dKs = np.zeros_like(Ks)
    dsym = np.zeros_like(symnet)
    for t in reversed(range(steps)):
        lam, dk, ds = step_bwd(states[t], lam, Ks, symnet, L, dt)
        dKs += dk
        dsym += ds
    return states[1:], lam, dKs, dsym


def burgers_central_step(u, v, nu, dt, dx):
    """Independent central-difference Burgers Euler step (KAT-4 target).

    Uses the classical 2nd-order 3-point formulas -- valid target only for k=3
    one-hot filters.  ``u,v [H,W]``; axis 0 is y (H), axis 1 is x (W).
    """
    def d_ax(f, ax):
        return (np.roll(f, -1, ax) - np.roll(f, 1, ax)) / (2 * dx)

    def d2_ax(f, ax):
        return (np.roll(f, -1, ax) - 2 * f + np.roll(f, 1, ax)) / dx ** 2

    ux, uy = d_ax(u, 1), d_ax(u, 0)
    vx, vy = d_ax(v, 1), d_ax(v, 0)
    lap_u = d2_ax(u, 0) + d2_ax(u, 1)
    lap_v = d2_ax(v, 0) + d2_ax(v, 1)
    un = u + dt * (-u * ux - v * uy + nu * lap_u)
    vn = v + dt * (-u * vx - v * vy + nu * lap_v)
    return un, vn
