"""torch.autograd bridge over the C-ABI: the only place Python touches the CUDA hot path.

Each function takes CUDA tensors, hands raw device pointers + the current stream to
libpdenet_b200.so, and wires the reverse kernels into autograd.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import os

import torch

from . import _lib


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _req(t: torch.Tensor, name: str, dtype=None) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError("pdenet_b200: `%s` must be a CUDA tensor (the hot path has no CPU fallback)" % name)
    if dtype is not None and t.dtype != dtype:
        raise TypeError("pdenet_b200: `%s` has dtype %s, expected %s" % (name, t.dtype, dtype))
    return t.contiguous()


def _scale_array(scale: Optional[Sequence[float]], N: int):
    if scale is None:
        return None
    if len(scale) != N:
        raise ValueError("scale must have one entry per filter")
    return (C.c_double * N)(*[float(s) for s in scale])


def symnet_param_count(m: int, L: int) -> int:
    return sum(2 * (m + i) + 2 for i in range(L)) + (m + L) + 1


# ------------------------------------------------------------------------------ rows a1-a3
class _M2K(torch.autograd.Function):
    @staticmethod
    def forward(ctx, moments, scale, mask):
        moments = _req(moments, "moments")
        N, k, _ = moments.shape
        ctx.scale, ctx.mask = scale, mask
        T = _lib.torch_ops()
        if T is not None:
            _scale_array(scale, N)   # length check
            return _lib.call_op("pdenet_m2k", T.m2k, moments, [] if scale is None else [float(s) for s in scale], None)
        out = torch.empty_like(moments)
        _lib.check(_lib.fn("pdenet_m2k", moments.dtype)(_ptr(moments), _scale_array(scale, N), _ptr(out), N, k,
                                                        _stream()), "pdenet_m2k")
        return out

    @staticmethod
    def backward(ctx, gk):
        gk = _req(gk, "grad_kernels")
        N, k, _ = gk.shape
        T = _lib.torch_ops()
        if T is not None:
            sc = [] if ctx.scale is None else [float(s) for s in ctx.scale]
            return _lib.call_op("pdenet_m2k_bwd", T.m2k_bwd, gk, sc, ctx.mask), None, None
        gm = torch.empty_like(gk)
        _lib.check(_lib.fn("pdenet_m2k_bwd", gk.dtype)(_ptr(gk), _scale_array(ctx.scale, N), _ptr(ctx.mask),
                                                       _ptr(gm), N, k, _stream()), "pdenet_m2k_bwd")
        return gm, None, None


def m2k(moments: torch.Tensor, scale: Optional[Sequence[float]] = None,
        pinned_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """kernels[n] = scale[n] * A^-1 moments[n] A^-T.  ``pinned_mask``
    (uint8 [N,k,k] on the device) zeroes the gradient of pinned moment entries (row a4)."""
    lead = moments.shape[:-2]
    k = moments.shape[-1]
    flat = moments.reshape(-1, k, k)
    if pinned_mask is not None:
        pinned_mask = _req(pinned_mask.reshape(-1, k, k).to(torch.uint8), "pinned_mask")
    return _M2K.apply(flat, scale, pinned_mask).reshape(*lead, k, k)


def k2m(kernels: torch.Tensor, scale: Optional[Sequence[float]] = None) -> torch.Tensor:
    """moments[n] = A (kernels[n] / scale[n]) A^T (row a3; inspection only, no autograd)."""
    lead = kernels.shape[:-2]
    k = kernels.shape[-1]
    flat = _req(kernels.detach().reshape(-1, k, k), "kernels")
    out = torch.empty_like(flat)
    _lib.check(_lib.fn("pdenet_k2m", flat.dtype)(_ptr(flat), _scale_array(scale, flat.shape[0]), _ptr(out),
                                                 flat.shape[0], k, _stream()), "pdenet_k2m")
    return out.reshape(*lead, k, k)


# ------------------------------------------------------------------------------ rows a5-a6
def fd2d_features(u: torch.Tensor, kernels: torch.Tensor) -> torch.Tensor:
    """[B,C,H,W] x [N,k,k] -> [B,C*N,H,W] periodic cross-correlation (no autograd; the
    differentiable path is :func:`rollout`)."""
    u = _req(u.detach(), "u")
    kernels = _req(kernels.detach(), "kernels", u.dtype)
    B, Cc, H, W = u.shape
    N, k, _ = kernels.shape
    out = torch.empty(B, Cc * N, H, W, dtype=u.dtype, device=u.device)
    _lib.check(_lib.fn("pdenet_fd2d_fwd", u.dtype)(_ptr(u), _ptr(kernels), _ptr(out), B, Cc, H, W, k, N,
                                                   _stream()), "pdenet_fd2d_fwd")
    return out


class _FD2d(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u, kernels):
        ctx.save_for_backward(u, kernels)
        return fd2d_features(u, kernels)

    @staticmethod
    def backward(ctx, gf):
        u, kernels = ctx.saved_tensors
        u, kernels = _req(u, "u"), _req(kernels, "kernels", u.dtype)
        gf = _req(gf, "grad_feats", u.dtype)
        B, Cc, H, W = u.shape
        N, k, _ = kernels.shape
        nbytes = _lib.lib().pdenet_fd2d_bwd_workspace_bytes(k, N)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=u.device)
        gu = torch.empty_like(u) if ctx.needs_input_grad[0] else None
        gk = torch.empty_like(kernels) if ctx.needs_input_grad[1] else None
        _lib.check(_lib.fn("pdenet_fd2d_bwd", u.dtype)(_ptr(u), _ptr(gf), _ptr(kernels), _ptr(gu), _ptr(gk),
                                                       _ptr(ws), nbytes, B, Cc, H, W, k, N, _stream()),
                   "pdenet_fd2d_bwd")
        return gu, gk


def fd2d(u: torch.Tensor, kernels: torch.Tensor) -> torch.Tensor:
    """Differentiable periodic FD2d bank: [B,C,H,W] x [N,k,k] -> [B,C*N,H,W] (rows a5-a6, a11)."""
    return _FD2d.apply(u, kernels)


def symnet_eval(feat: torch.Tensor, flat: torch.Tensor, L: int) -> torch.Tensor:
    """Standalone SymNet evaluation ``feat [..., m] -> [...]`` with the flat C-ABI parameter
    layout.  Cold path (inspection / unit use of ``expr.poly``): plain tensor ops on the
    caller's device; the time-stepping path evaluates SymNet inside the fused rollout kernel."""
    m = feat.shape[-1]
    o, h = 0, feat
    for i in range(L):
        mi = m + i
        Wt = flat[o:o + 2 * mi].reshape(2, mi)
        b = flat[o + 2 * mi:o + 2 * mi + 2]
        o += 2 * mi + 2
        pq = h @ Wt.t() + b
        h = torch.cat([h, pq[..., :1] * pq[..., 1:]], dim=-1)
    return h @ flat[o:o + m + L] + flat[o + m + L]


# ------------------------------------------------------------------------------ rows a7-a11
#: HBM budget for keeping the derivative features of a rollout alive between the forward and the
#: reverse pass (``keep_features=None``): below it the reverse pass reads them back instead of
#: recomputing the stencils (cfg4: 96 B per grid point and step in fp64).
FEATURE_CACHE_BYTES = int(float(os.environ.get("PDENET_FEATURE_CACHE_GB", "32")) * 2 ** 30)


_F_KEEP, _F_UPWIND, _F_LAST = 1, 2, 4   # flags of the dispatcher ops (csrc/torch_shim.cpp)


def _rollout_fwd_ctypes(u0, kernels, symnet, L, dt, steps, keep, upwind):
    B, Cc, H, W = u0.shape
    N, k, _ = kernels.shape
    states = torch.empty(steps, B, Cc, H, W, dtype=u0.dtype, device=u0.device)
    feats = None
    if upwind:
        _lib.check(_lib.fn("pdenet_rollout_fwd_upwind", u0.dtype)(
            _ptr(u0), _ptr(kernels), _ptr(symnet), _ptr(states), B, Cc, H, W, k, N, L, steps, float(dt),
            _stream()), "pdenet_rollout_fwd_upwind")
    elif keep:
        feats = torch.empty(steps, B, Cc * N, H, W, dtype=u0.dtype, device=u0.device)
        _lib.check(_lib.fn("pdenet_rollout_fwd_keep", u0.dtype)(
            _ptr(u0), _ptr(kernels), _ptr(symnet), _ptr(states), _ptr(feats), B, Cc, H, W, k, N, L, steps,
            float(dt), _stream()), "pdenet_rollout_fwd_keep")
    else:
        _lib.check(_lib.fn("pdenet_rollout_fwd", u0.dtype)(
            _ptr(u0), _ptr(kernels), _ptr(symnet), _ptr(states), B, Cc, H, W, k, N, L, steps, float(dt),
            _stream()), "pdenet_rollout_fwd")
    return states, feats


def _rollout_fwd(u0, kernels, symnet, L, dt, steps, keep, upwind):
    """-> (states, feats or None) through the dispatcher op (TORCH_LIBRARY shim) or the ctypes binding."""
    T = _lib.torch_ops()
    if T is None:
        return _rollout_fwd_ctypes(u0, kernels, symnet, L, dt, steps, keep, upwind)
    flags = (_F_KEEP if keep and not upwind else 0) | (_F_UPWIND if upwind else 0)
    states, feats = _lib.call_op("pdenet_rollout_fwd", T.rollout_fwd, u0, kernels, symnet, L, float(dt), steps, flags)
    return states, (feats if flags & _F_KEEP else None)


def _rollout_bwd(u0, states, feats, grad, kernels, symnet, L, dt, steps, last_only, batch_chunk, need, upwind):
    B, Cc, H, W = u0.shape
    N, k, _ = kernels.shape
    need_u, need_k, need_s = need
    T = _lib.torch_ops()
    if T is not None:
        flags = (_F_KEEP if feats is not None else 0) | (_F_UPWIND if upwind else 0) | (_F_LAST if last_only else 0)
        empty = u0.new_empty(0)
        gu, gk, gs = _lib.call_op("pdenet_rollout_bwd", T.rollout_bwd, u0, states if states is not None else empty,
                                  feats if feats is not None else empty, grad, kernels, symnet, L, float(dt), steps,
                                  batch_chunk, bool(need_u), bool(need_k), bool(need_s), flags)
        return (gu if need_u else None), (gk if need_k else None), (gs if need_s else None)
    es = u0.element_size()
    wsfn = _lib.lib().pdenet_rollout_bwd_upwind_workspace_bytes if upwind \
        else _lib.lib().pdenet_rollout_bwd_workspace_bytes
    nbytes = wsfn(es, B, Cc, H, W, k, N, L, batch_chunk)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=u0.device)
    gu = torch.empty_like(u0) if need_u else None
    gk = torch.empty_like(kernels) if need_k else None
    gs = torch.empty_like(symnet) if need_s else None
    mode = _lib.GRAD_LAST if last_only else _lib.GRAD_ALL
    if upwind:
        _lib.check(_lib.fn("pdenet_rollout_bwd_upwind", u0.dtype)(
            _ptr(u0), _ptr(states), _ptr(grad), mode, _ptr(kernels), _ptr(symnet), _ptr(gu), _ptr(gk), _ptr(gs),
            _ptr(ws), nbytes, batch_chunk, B, Cc, H, W, k, N, L, steps, dt, _stream()),
            "pdenet_rollout_bwd_upwind")
    elif feats is not None:
        _lib.check(_lib.fn("pdenet_rollout_bwd_kept", u0.dtype)(
            _ptr(u0), _ptr(states), _ptr(feats), _ptr(grad), mode, _ptr(kernels), _ptr(symnet), _ptr(gu),
            _ptr(gk), _ptr(gs), _ptr(ws), nbytes, batch_chunk, B, Cc, H, W, k, N, L, steps, dt, _stream()),
            "pdenet_rollout_bwd_kept")
    else:
        _lib.check(_lib.fn("pdenet_rollout_bwd", u0.dtype)(
            _ptr(u0), _ptr(states), _ptr(grad), mode, _ptr(kernels), _ptr(symnet), _ptr(gu), _ptr(gk), _ptr(gs),
            _ptr(ws), nbytes, batch_chunk, B, Cc, H, W, k, N, L, steps, dt, _stream()), "pdenet_rollout_bwd")
    return gu, gk, gs


class _Rollout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u0, kernels, symnet, L, dt, steps, last_only, batch_chunk, keep_features, upwind=False):
        u0 = _req(u0, "u0")
        kernels = _req(kernels, "kernels", u0.dtype)
        symnet = _req(symnet, "symnet", u0.dtype)
        B, Cc, H, W = u0.shape
        N, k, _ = kernels.shape
        if tuple(symnet.shape) != (Cc, symnet_param_count(Cc * N, L)):
            raise ValueError("symnet must be [C, Psym] = [%d, %d], got %s"
                             % (Cc, symnet_param_count(Cc * N, L), tuple(symnet.shape)))
        ctx.upwind = bool(upwind)
        auto = False
        if upwind:   # pseudo-upwind scheme (row f-1): features are always recomputed by the reverse pass
            keep_features = False
        else:
            feat_bytes = steps * B * Cc * N * H * W * u0.element_size()
            auto = keep_features is None
            if auto:
                # a few banks (3 channels, deep SymNet, fp64) only fit the reverse pass with kept features
                needs = _lib.lib().pdenet_rollout_bwd_needs_features(u0.element_size(), k, N, Cc, L) == 1
                keep_features = needs or feat_bytes <= FEATURE_CACHE_BYTES
                auto = not needs
            keep_features = bool(keep_features) and any(ctx.needs_input_grad[:3])
        try:
            states, feats = _rollout_fwd(u0, kernels, symnet, L, dt, steps, keep_features, ctx.upwind)
        except torch.OutOfMemoryError:
            if not (auto and keep_features):
                raise
            # no room for the feature buffer: the reverse pass recomputes the features instead
            states, feats = _rollout_fwd(u0, kernels, symnet, L, dt, steps, False, False)
        ctx.feats = feats  # an intermediate, never an output: no reference cycle
        ctx.cfg = (L, float(dt), steps, last_only, batch_chunk)
        if last_only:  # `states` is an intermediate: keep it on ctx (no cycle, it is not an output)
            ctx.save_for_backward(u0, kernels, symnet)
            ctx.states = states
            return states[-1]
        ctx.save_for_backward(u0, kernels, symnet, states)
        return states

    @staticmethod
    def backward(ctx, grad):
        if ctx.cfg[3]:
            (u0, kernels, symnet), states = ctx.saved_tensors, ctx.states
        else:
            u0, kernels, symnet, states = ctx.saved_tensors
        L, dt, steps, last_only, batch_chunk = ctx.cfg
        grad = _req(grad, "grad_states", u0.dtype)
        gu, gk, gs = _rollout_bwd(u0, states, ctx.feats, grad, kernels, symnet, L, dt, steps, last_only, batch_chunk,
                                  ctx.needs_input_grad[:3], ctx.upwind)
        return gu, gk, gs, None, None, None, None, None, None, None


def _scheme_is_upwind(scheme: str) -> bool:
    if scheme not in ("central", "upwind"):
        raise ValueError("scheme must be 'central' or 'upwind', got %r" % (scheme,))
    return scheme == "upwind"


def rollout(u0: torch.Tensor, kernels: torch.Tensor, symnet: torch.Tensor, L: int, dt: float, steps: int,
            last_only: bool = False, batch_chunk: int = 0, keep_features=None, scheme: str = "central") -> torch.Tensor:
    """Fused forward-Euler rollout ``u <- u + dt*SymNet(FD2d(u))`` (rows a5-a10), differentiable
    w.r.t. ``u0``, the scaled ``kernels`` [N,k,k] and the flat ``symnet`` [C,Psym] (row a11).

    ``keep_features``: True keeps the C*N derivative features of every step for the reverse pass
    (as the reference's autograd does), False recomputes them there, None decides by
    ``FEATURE_CACHE_BYTES``.  The results are identical bit for bit.

    ``scheme='upwind'``: the pseudo-upwind scheme (row f-1; ``include/pdenet_b200.h``): per output
    channel and grid point the first-order inputs of the SymNet come from the filter as given or from
    its flipped twin, by the sign of the SymNet's own derivative w.r.t. that input.

    Returns the states after every step ``[steps,B,C,H,W]`` (or only the last one)."""
    if steps < 1:
        raise ValueError("steps must be >= 1")
    return _Rollout.apply(u0, kernels, symnet, int(L), float(dt), int(steps), bool(last_only), int(batch_chunk),
                          keep_features, _scheme_is_upwind(scheme))


# ------------------------------------------------------------------------------ row a8
class _Rhs(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u, kernels, symnet, L, batch_chunk, upwind=False):
        u = _req(u, "u")
        kernels = _req(kernels, "kernels", u.dtype)
        symnet = _req(symnet, "symnet", u.dtype)
        B, Cc, H, W = u.shape
        N, k, _ = kernels.shape
        if tuple(symnet.shape) != (Cc, symnet_param_count(Cc * N, L)):
            raise ValueError("symnet must be [C, Psym] = [%d, %d], got %s"
                             % (Cc, symnet_param_count(Cc * N, L), tuple(symnet.shape)))
        name = "pdenet_rhs_upwind_fwd" if upwind else "pdenet_rhs_fwd"
        T = _lib.torch_ops()
        if T is not None:
            out = _lib.call_op(name, T.rhs_fwd, u, kernels, symnet, L, _F_UPWIND if upwind else 0)
        else:
            out = torch.empty_like(u)
            _lib.check(_lib.fn(name, u.dtype)(_ptr(u), _ptr(kernels), _ptr(symnet), _ptr(out), B, Cc, H, W, k, N, L,
                                              _stream()), name)
        ctx.save_for_backward(u, kernels, symnet)
        ctx.cfg = (L, batch_chunk, bool(upwind))
        return out

    @staticmethod
    def backward(ctx, grad):
        u, kernels, symnet = ctx.saved_tensors
        L, batch_chunk, upwind = ctx.cfg
        B, Cc, H, W = u.shape
        N, k, _ = kernels.shape
        grad = _req(grad, "grad_out", u.dtype)
        name = "pdenet_rhs_upwind_bwd" if upwind else "pdenet_rhs_bwd"
        T = _lib.torch_ops()
        if T is not None:
            nu, nk, ns = ctx.needs_input_grad[:3]
            gu, gk, gs = _lib.call_op(name, T.rhs_bwd, u, grad, kernels, symnet, L, batch_chunk, bool(nu), bool(nk),
                                      bool(ns), _F_UPWIND if upwind else 0)
            return (gu if nu else None), (gk if nk else None), (gs if ns else None), None, None, None
        wsfn = _lib.lib().pdenet_rollout_bwd_upwind_workspace_bytes if upwind \
            else _lib.lib().pdenet_rollout_bwd_workspace_bytes
        nbytes = wsfn(u.element_size(), B, Cc, H, W, k, N, L, batch_chunk)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=u.device)
        need_u, need_k, need_s = ctx.needs_input_grad[:3]
        gu = torch.empty_like(u) if need_u else None
        gk = torch.empty_like(kernels) if need_k else None
        gs = torch.empty_like(symnet) if need_s else None
        _lib.check(_lib.fn(name, u.dtype)(
            _ptr(u), _ptr(grad), _ptr(kernels), _ptr(symnet), _ptr(gu), _ptr(gk), _ptr(gs), _ptr(ws), nbytes,
            batch_chunk, B, Cc, H, W, k, N, L, _stream()), name)
        return gu, gk, gs, None, None, None


def rhs(u: torch.Tensor, kernels: torch.Tensor, symnet: torch.Tensor, L: int, batch_chunk: int = 0,
        scheme: str = "central") -> torch.Tensor:
    """The right-hand side ``F(u) = SymNet(FD2d(u))`` [B,C,H,W] itself (row a8, upstream
    ``POLYPDE2D.RightHandItems``): the fused step kernel without the Euler update, differentiable
    w.r.t. ``u``, the scaled ``kernels`` and the flat ``symnet``; ``scheme`` as in :func:`rollout`."""
    return _Rhs.apply(u, kernels, symnet, int(L), int(batch_chunk), _scheme_is_upwind(scheme))


# ------------------------------------------------------------------------------ row a12
class _LinearRollout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u0, id_kernel, kernels, coe, dt, steps, last_only):
        u0 = _req(u0, "u0")
        id_kernel = _req(id_kernel, "id_kernel", u0.dtype)
        kernels = _req(kernels, "kernels", u0.dtype)
        coe = _req(coe, "coe", u0.dtype)
        B, H, W = u0.shape
        N, k, _ = kernels.shape
        if tuple(coe.shape) != (N, H, W):
            raise ValueError("coe must be [N,H,W]")
        T = _lib.torch_ops()
        if T is not None:
            states = _lib.call_op("pdenet_linear_rollout_fwd", T.linear_rollout_fwd, u0, id_kernel, kernels, coe,
                                  float(dt), steps)
        else:
            states = torch.empty(steps, B, H, W, dtype=u0.dtype, device=u0.device)
            _lib.check(_lib.fn("pdenet_linear_rollout_fwd", u0.dtype)(
                _ptr(u0), _ptr(id_kernel), _ptr(kernels), _ptr(coe), _ptr(states), B, H, W, k, N, steps, float(dt),
                _stream()), "pdenet_linear_rollout_fwd")
        ctx.cfg = (float(dt), steps, last_only)
        if last_only:
            ctx.save_for_backward(u0, id_kernel, kernels, coe)
            ctx.states = states
            return states[-1]
        ctx.save_for_backward(u0, id_kernel, kernels, coe, states)
        return states

    @staticmethod
    def backward(ctx, grad):
        if ctx.cfg[2]:
            (u0, id_kernel, kernels, coe), states = ctx.saved_tensors, ctx.states
        else:
            u0, id_kernel, kernels, coe, states = ctx.saved_tensors
        dt, steps, last_only = ctx.cfg
        B, H, W = u0.shape
        N, k, _ = kernels.shape
        grad = _req(grad, "grad_states", u0.dtype)
        T = _lib.torch_ops()
        if T is not None:
            n = [bool(x) for x in ctx.needs_input_grad[:4]]
            out = _lib.call_op("pdenet_linear_rollout_bwd", T.linear_rollout_bwd, u0, states, grad, id_kernel, kernels,
                               coe, dt, steps, bool(last_only), *n)
            return tuple(g if need else None for g, need in zip(out, n)) + (None, None, None)
        nbytes = _lib.lib().pdenet_linear_rollout_bwd_workspace_bytes(u0.element_size(), B, H, W, k, N)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=u0.device)
        need = ctx.needs_input_grad
        gu = torch.empty_like(u0) if need[0] else None
        g0 = torch.empty_like(id_kernel) if need[1] else None
        gk = torch.empty_like(kernels) if need[2] else None
        gc = torch.empty_like(coe) if need[3] else None
        _lib.check(_lib.fn("pdenet_linear_rollout_bwd", u0.dtype)(
            _ptr(u0), _ptr(states), _ptr(grad), _lib.GRAD_LAST if last_only else _lib.GRAD_ALL, _ptr(id_kernel),
            _ptr(kernels), _ptr(coe), _ptr(gu), _ptr(g0), _ptr(gk), _ptr(gc), _ptr(ws), nbytes, B, H, W, k, N,
            steps, dt, _stream()), "pdenet_linear_rollout_bwd")
        return gu, g0, gk, gc, None, None, None


def linear_rollout(u0, id_kernel, kernels, coe, dt: float, steps: int, last_only: bool = False):
    """PDE-Net 1 rollout ``u <- K0*u + dt*sum_n coe_n (K_n*u)`` (row a12); u0 [B,H,W]."""
    if steps < 1:
        raise ValueError("steps must be >= 1")
    return _LinearRollout.apply(u0, id_kernel, kernels, coe, float(dt), int(steps), bool(last_only))


def slab_step(u_ext: torch.Tensor, kernels: torch.Tensor, symnet: torch.Tensor, L: int, dt: float, halo: int,
              out: Optional[torch.Tensor] = None, out_row_offset: int = 0, peer_up: int = 0, peer_dn: int = 0,
              mirror_rows: int = 0, feats: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One forward step on a slab [B,C,Hs+2*halo,W] whose ghost rows are already filled
    (slab domain decomposition).  Writes rows [out_row_offset, out_row_offset+Hs)
    of ``out`` (default: a fresh [B,C,Hs,W]).  With ``mirror_rows > 0`` the kernel also stores its
    first / last ``mirror_rows`` rows into the ring neighbours' output buffers (``peer_up`` /
    ``peer_dn``: raw device pointers of the neighbours' buffers mapped into this process) -- the
    halo exchange fused into the step over NVLink peer memory.  ``feats`` [B,C*N,Hs,W] (optional)
    receives the derivative features of the owned rows for :class:`SlabBackward`."""
    u_ext = _req(u_ext.detach(), "u_ext")
    kernels = _req(kernels.detach(), "kernels", u_ext.dtype)
    symnet = _req(symnet.detach(), "symnet", u_ext.dtype)
    B, Cc, He, W = u_ext.shape
    Hs = He - 2 * halo
    N, k, _ = kernels.shape
    if out is None:
        out = torch.empty(B, Cc, Hs, W, dtype=u_ext.dtype, device=u_ext.device)
    if not out.is_contiguous():
        raise ValueError("out must be contiguous")
    if feats is not None:
        feats = _req(feats, "feats", u_ext.dtype)
        if tuple(feats.shape) != (B, Cc * N, Hs, W):
            raise ValueError("feats must be [B, C*N, Hs, W]")
        _lib.check(_lib.fn("pdenet_slab_step_fwd_keep", u_ext.dtype)(
            _ptr(u_ext), _ptr(kernels), _ptr(symnet), _ptr(out), _ptr(feats), C.c_void_p(peer_up or None),
            C.c_void_p(peer_dn or None), mirror_rows, B, Cc, Hs, W, halo, out.shape[2], out_row_offset, k, N, L,
            float(dt), _stream()), "pdenet_slab_step_fwd_keep")
        return out
    _lib.check(_lib.fn("pdenet_slab_step_fwd_p2p", u_ext.dtype)(
        _ptr(u_ext), _ptr(kernels), _ptr(symnet), _ptr(out), C.c_void_p(peer_up or None), C.c_void_p(peer_dn or None),
        mirror_rows, B, Cc, Hs, W, halo, out.shape[2], out_row_offset, k, N, L, float(dt), _stream()),
        "pdenet_slab_step_fwd_p2p")
    return out


# ------------------------------------------------------------------------------ row f-2
class _Lagrange1d(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, coef, lo, hi, mesh, degree):
        x = _req(x, "x")
        coef = _req(coef, "coef", x.dtype)
        if coef.numel() != mesh * degree + 1:
            raise ValueError("coef must hold mesh*degree + 1 = %d node values" % (mesh * degree + 1))
        y = torch.empty_like(x)
        if x.numel():
            _lib.check(_lib.fn("pdenet_lagrange1d_fwd", x.dtype)(_ptr(x), _ptr(coef), _ptr(y), x.numel(), lo, hi, mesh,
                                                                 degree, _stream()), "pdenet_lagrange1d_fwd")
        ctx.save_for_backward(x, coef)
        ctx.cfg = (lo, hi, mesh, degree)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, coef = ctx.saved_tensors
        lo, hi, mesh, degree = ctx.cfg
        gy = _req(gy, "grad", x.dtype)
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gc = torch.zeros_like(coef)
        if x.numel():
            nbytes = _lib.lib().pdenet_lagrange1d_bwd_workspace_bytes(x.numel(), mesh, degree)
            ws = torch.empty(max(nbytes, 8), dtype=torch.uint8, device=x.device)
            _lib.check(_lib.fn("pdenet_lagrange1d_bwd", x.dtype)(
                _ptr(x), _ptr(gy), _ptr(coef), _ptr(gx), _ptr(gc), x.numel(), lo, hi, mesh, degree, _ptr(ws), nbytes,
                _stream()), "pdenet_lagrange1d_bwd")
        elif gx is not None:
            gx.zero_()
        return gx, gc, None, None, None, None


def lagrange1d(x: torch.Tensor, coef: torch.Tensor, lo: float, hi: float, mesh: int, degree: int) -> torch.Tensor:
    """Piecewise Lagrange interpolant of degree ``degree`` on ``mesh`` uniform cells of [lo, hi],
    evaluated at every entry of ``x`` (any shape); ``coef`` [mesh*degree + 1] are the node values.
    The nonlinear source term of PDE-Net 1's ``SingleNonLinear2d``."""
    return _Lagrange1d.apply(x, coef, float(lo), float(hi), int(mesh), int(degree))


class _SingleNonlinearRollout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u0, id_kernel, kernels, coe, fcoef, dt, steps, lo, hi, mesh, degree):
        u0 = _req(u0, "u0")
        id_kernel, kernels = _req(id_kernel, "id_kernel", u0.dtype), _req(kernels, "kernels", u0.dtype)
        coe, fcoef = _req(coe, "coe", u0.dtype), _req(fcoef, "fcoef", u0.dtype)
        B, H, W = u0.shape
        N, k, _ = kernels.shape
        if coe.numel() != N or fcoef.numel() != mesh * degree + 1:
            raise ValueError("coe must hold N constants and fcoef mesh*degree + 1 node values")
        states = torch.empty(steps, B, H, W, dtype=u0.dtype, device=u0.device)
        _lib.check(_lib.fn("pdenet_single_nonlinear_rollout_fwd", u0.dtype)(
            _ptr(u0), _ptr(id_kernel), _ptr(kernels), _ptr(coe), _ptr(fcoef), _ptr(states), B, H, W, k, N, steps,
            float(dt), lo, hi, mesh, degree, _stream()), "pdenet_single_nonlinear_rollout_fwd")
        ctx.save_for_backward(u0, id_kernel, kernels, coe, fcoef)
        ctx.states = states
        ctx.cfg = (float(dt), steps, lo, hi, mesh, degree)
        return states[-1]

    @staticmethod
    def backward(ctx, grad):
        u0, id_kernel, kernels, coe, fcoef = ctx.saved_tensors
        dt, steps, lo, hi, mesh, degree = ctx.cfg
        B, H, W = u0.shape
        N, k, _ = kernels.shape
        grad = _req(grad, "grad_state", u0.dtype)
        nbytes = _lib.lib().pdenet_single_nonlinear_rollout_bwd_workspace_bytes(u0.element_size(), B, H, W, k, N, steps,
                                                                                mesh, degree)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=u0.device)
        need = ctx.needs_input_grad
        outs = [torch.empty_like(t) if n else None for t, n in zip((u0, id_kernel, kernels, coe, fcoef), need[:5])]
        _lib.check(_lib.fn("pdenet_single_nonlinear_rollout_bwd", u0.dtype)(
            _ptr(u0), _ptr(ctx.states), _ptr(grad), _lib.GRAD_LAST, _ptr(id_kernel), _ptr(kernels), _ptr(coe), _ptr(fcoef),
            *[_ptr(o) for o in outs], _ptr(ws), nbytes, B, H, W, k, N, steps, dt, lo, hi, mesh, degree, _stream()),
            "pdenet_single_nonlinear_rollout_bwd")
        return tuple(outs) + (None,) * 6


def single_nonlinear_fits(u0: torch.Tensor, kernels: torch.Tensor) -> bool:
    """True if the fused SingleNonLinear2d rollout (cluster-resident small-grid path) serves this shape."""
    B, H, W = u0.shape
    N, k, _ = kernels.shape
    return _lib.lib().pdenet_single_nonlinear_fits(u0.element_size(), B, H, W, k, N) == 1


def single_nonlinear_rollout(u0, id_kernel, kernels, coe, fcoef, dt: float, steps: int, lo: float, hi: float, mesh: int,
                             degree: int) -> torch.Tensor:
    """PDE-Net 1 ``SingleNonLinear2d`` rollout ``u <- K0*u + dt*(sum_n coe_n (K_n*u) + f(u))`` as ONE fused op (row
    f-2): final state after ``steps`` blocks, differentiable w.r.t. everything.  Small grids only
    (:func:`single_nonlinear_fits`); ``coe`` [N] constants, ``fcoef`` the node values of ``f``."""
    if steps < 1:
        raise ValueError("steps must be >= 1")
    return _SingleNonlinearRollout.apply(u0, id_kernel, kernels, coe, fcoef, float(dt), int(steps), float(lo), float(hi),
                                         int(mesh), int(degree))


def mse_seed(states: torch.Tensor, observed: torch.Tensor, weights: Optional[Sequence[float]] = None):
    """``(loss, seed)`` without an autograd node: ``loss = sum_t w_t * mean((states[t]-observed[t])**2)`` (a 0-d
    tensor of the states' dtype) and ``seed = d loss / d states`` from ONE pass over the states
    (``pdenet_mse_seed_*``).  A caller that drives the reverse rollout itself (``states.backward(seed)``,
    :class:`losses.GraphedTrainingLoss`) needs neither the ones-filled root gradient nor the ``seed * 1`` pass that
    going through :func:`mse_loss` and ``loss.backward()`` costs."""
    states = _req(states.detach(), "states")
    observed = _req(observed.detach(), "observed", states.dtype)
    if states.shape != observed.shape:
        raise ValueError("states and observed must have the same shape [steps, ...]")
    steps = states.shape[0]
    n = states[0].numel()
    grad = torch.empty_like(states)
    loss = torch.empty((), dtype=torch.float64, device=states.device)
    ws = torch.empty(64 * 512, dtype=torch.float64, device=states.device)
    wh = None if weights is None else (C.c_double * steps)(*[float(w) for w in weights])
    _lib.check(_lib.fn("pdenet_mse_seed", states.dtype)(_ptr(states), _ptr(observed), wh, _ptr(grad), _ptr(loss), n,
                                                        steps, _ptr(ws), ws.numel() * 8, _stream()),
               "pdenet_mse_seed")
    return loss.to(states.dtype), grad


class _MseLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, states, observed, weights):
        loss, grad = mse_seed(states, observed, weights)
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, gout):
        (grad,) = ctx.saved_tensors
        return grad * gout, None, None


def mse_loss(states: torch.Tensor, observed: torch.Tensor, weights: Optional[Sequence[float]] = None) -> torch.Tensor:
    """Data term of the training loss, ``sum_t w_t * mean((states[t]-observed[t])**2)``, and its
    gradient seed in ONE pass over the states: the returned scalar back-propagates
    the precomputed seed into :func:`rollout`'s reverse pass."""
    return _MseLoss.apply(states, observed, weights)


class SlabBackward:
    """Reverse pass of a slab-decomposed rollout on one rank (C-ABI pdenet_slab_*bwd*).

    ``step(u_ext2, lam_ext, lam_out)`` is called once per step, last to first, with the step's
    input state extended by 2r ghost rows and the adjoint of its output (seed included) extended by
    r ghost rows; ``finish()`` returns this rank's (grad_kernels, grad_symnet) contribution."""

    def __init__(self, kernels: torch.Tensor, symnet: torch.Tensor, L: int, dt: float, B: int, C: int, Hs: int,
                 W: int):
        self.kernels = _req(kernels.detach(), "kernels")
        self.symnet = _req(symnet.detach(), "symnet", self.kernels.dtype)
        self.L, self.dt = int(L), float(dt)
        self.B, self.C, self.Hs, self.W = B, C, Hs, W
        self.N, self.k = self.kernels.shape[0], self.kernels.shape[-1]
        self.r = (self.k - 1) // 2
        es = self.kernels.element_size()
        self.nbytes = _lib.lib().pdenet_slab_bwd_workspace_bytes(es, B, C, Hs, W, self.k, self.N, self.L)
        self.ws = torch.empty(self.nbytes, dtype=torch.uint8, device=self.kernels.device)
        _lib.check(_lib.lib().pdenet_slab_bwd_begin(_ptr(self.ws), self.nbytes, es, B, C, Hs, W, self.k, self.N, self.L,
                                                    _stream()), "pdenet_slab_bwd_begin")

    def step(self, u_ext2: torch.Tensor, lam_ext: torch.Tensor, lam_out: torch.Tensor, out_row_offset: int = 0,
             peer_up: int = 0, peer_dn: int = 0, mirror_rows: int = 0, feats: Optional[torch.Tensor] = None):
        dt_ = self.kernels.dtype
        u_ext2, lam_ext = _req(u_ext2, "u_ext2", dt_), _req(lam_ext, "lam_ext", dt_)
        if tuple(u_ext2.shape) != (self.B, self.C, self.Hs + 4 * self.r, self.W):
            raise ValueError("u_ext2 must carry 2r ghost rows on each side")
        if tuple(lam_ext.shape) != (self.B, self.C, self.Hs + 2 * self.r, self.W):
            raise ValueError("lam_ext must carry r ghost rows on each side")
        if not lam_out.is_contiguous():
            raise ValueError("lam_out must be contiguous")
        if feats is not None:  # the forward step kept the features of the owned rows
            feats = _req(feats, "feats", dt_)
            if tuple(feats.shape) != (self.B, self.C * self.N, self.Hs, self.W):
                raise ValueError("feats must be [B, C*N, Hs, W]")
            _lib.check(_lib.fn("pdenet_slab_step_bwd_kept", dt_)(
                _ptr(u_ext2), _ptr(lam_ext), _ptr(feats), _ptr(self.kernels), _ptr(self.symnet), _ptr(lam_out),
                C.c_void_p(peer_up or None), C.c_void_p(peer_dn or None), mirror_rows, lam_out.shape[2],
                out_row_offset, _ptr(self.ws), self.nbytes, self.B, self.C, self.Hs, self.W, self.k, self.N, self.L,
                self.dt, _stream()), "pdenet_slab_step_bwd_kept")
            return lam_out
        _lib.check(_lib.fn("pdenet_slab_step_bwd_p2p", dt_)(
            _ptr(u_ext2), _ptr(lam_ext), _ptr(self.kernels), _ptr(self.symnet), _ptr(lam_out),
            C.c_void_p(peer_up or None), C.c_void_p(peer_dn or None), mirror_rows, lam_out.shape[2], out_row_offset,
            _ptr(self.ws), self.nbytes, self.B, self.C, self.Hs, self.W, self.k, self.N, self.L, self.dt, _stream()),
            "pdenet_slab_step_bwd_p2p")
        return lam_out

    def finish(self):
        gk, gs = torch.empty_like(self.kernels), torch.empty_like(self.symnet)
        _lib.check(_lib.fn("pdenet_slab_bwd_end", self.kernels.dtype)(
            _ptr(self.ws), self.nbytes, _ptr(gk), _ptr(gs), self.B, self.C, self.Hs, self.W, self.k, self.N, self.L,
            _stream()), "pdenet_slab_bwd_end")
        return gk, gs


def fma_peak(dtype=torch.float64, iters: int = 4096, variant: int = 0):
    """Measured CUDA-core FMA throughput in FLOP/s (register-only loop).  ``variant`` selects the
    operand pattern (0: loop-invariant multiplicands, 1: three distinct register operands per FMA,
    2: one operand shared by four consecutive FMAs, 3 / 4: the filter-gradient kernel's pattern -- 49
    accumulators, one multiplicand shared by 7 FMAs, window rows from shared memory -- with two /
    one columns per lane, 5: the same accumulators with all operands resident in registers, 6: variant 5
    plus the per-row window loads from shared memory)."""
    flops, ms = C.c_double(0), C.c_float(0)
    _lib.check(_lib.lib().pdenet_fma_probe(8 if dtype == torch.float64 else 4, iters, variant, C.byref(flops),
                                           C.byref(ms), _stream()), "pdenet_fma_probe")
    return flops.value / (ms.value * 1e-3)


PROFILE_KINDS = ("fwd_step", "bwd_kernel_A", "bwd_kernel_B1", "bwd_kernel_B2", "small")


def profile_enable(on: bool = True) -> None:
    """Bracket every library launch with CUDA events (measurement aid for bench.py)."""
    _lib.check(_lib.lib().pdenet_profile_enable(1 if on else 0), "pdenet_profile_enable")


def profile_read():
    """-> ({kind: summed ms}, {kind: launches}) since the last read; resets the counters."""
    ms = (C.c_double * len(PROFILE_KINDS))()
    n = (C.c_longlong * len(PROFILE_KINDS))()
    _lib.check(_lib.lib().pdenet_profile_read(ms, n), "pdenet_profile_read")
    return dict(zip(PROFILE_KINDS, list(ms))), dict(zip(PROFILE_KINDS, list(n)))
