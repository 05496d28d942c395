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


def step_bwd(u, lam, Ks, symnet, L, dt):
    """Adjoint of one Euler step.  Returns (lam_prev, dKs, dsymnet)."""
    B, C, H, W = u.shape
    N, k, _ = Ks.shape
    r = (k - 1) // 2
    m = C * N
    d = fd2d_apply(u, Ks)
    g = np.zeros_like(d)
    dsym = np.zeros_like(symnet)
    for c in range(C):
        flat = symnet[c]
        hidden, wf, bf = _unpack(flat, m, L)
        _, h, pq = symnet_fwd(d, flat, L)
        Lam = dt * lam[:, c]
        hbar = [Lam * wf[j] for j in range(m + L)]
        # flat gradient, same layout as the parameters
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
        dsym[c] = gflat
        for j in range(m):
            g[:, j] += hbar[j]
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
    return lam_prev, dKs, dsym


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
