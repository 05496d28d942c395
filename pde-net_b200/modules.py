"""Host-side mirror of the reference's Python model API for the time-stepping path.

The mounted reference has no source (README.md:8 points at absent branches), so the names below
follow SURVEY.md section 8(b): upstream ``aTEAM.nn.modules.MK.{M2K,K2M}``, ``aTEAM.nn.modules.FD.FD2d``,
``expr.poly`` and ``polypde.POLYPDE2D`` / ``VariantCoeLinear2d`` (recollection, not verifiable here).
Same constructor arguments, parameter names and method names; the inner loop of ``multistep`` is
one call into the fused CUDA rollout (``ops.rollout``) instead of a Python loop over eager ops.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import torch
import torch.nn as nn

from . import ops


def order_list(max_order: int) -> List[Tuple[int, int]]:
    """Filter-bank order: total order ascending, then order along H ascending."""
    return [(j, o - j) for o in range(max_order + 1) for j in range(o + 1)]


def _pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


def _bank_kernels(model, fds) -> torch.Tensor:
    """Scaled kernels [len(fds), k, k] of a list of FD2d filters: one fused m2k call.  With a
    ``FlatParams`` attached (``optim.FlatParams(model)``) the moments are read as ONE zero-copy view of
    the flat parameter buffer and their gradient is accumulated straight into the flat gradient
    buffer; otherwise they are stacked.  The pinned-entry mask is static and cached per device."""
    fp = getattr(model, "_flat", None)
    k = fds[0].kernel_size
    if fp is not None:
        moments = fp.group([fd.moment for fd in fds], (len(fds), k, k))
    else:
        moments = torch.stack([fd.moment for fd in fds])
    cache = model.__dict__.setdefault("_bank_cache", {})
    key = (tuple(id(fd) for fd in fds), moments.device)
    if key not in cache:
        cache[key] = (torch.stack([fd._pinned for fd in fds]).to(moments.device, torch.uint8).contiguous(),
                      [fd.scale for fd in fds])
    mask, scales = cache[key]
    return ops.m2k(moments, scales, mask)


class M2K(nn.Module):
    """Moment matrix -> convolution kernel (row a2).  ``shape`` is the kernel shape (k, k)."""

    def __init__(self, shape):
        super().__init__()
        self.shape = _pair(shape)

    def forward(self, m: torch.Tensor) -> torch.Tensor:
        return ops.m2k(m)


class K2M(nn.Module):
    """Convolution kernel -> moment matrix (row a3)."""

    def __init__(self, shape):
        super().__init__()
        self.shape = _pair(shape)

    def forward(self, k: torch.Tensor) -> torch.Tensor:
        return ops.k2m(k)


class FD2d(nn.Module):
    """One learnable periodic finite-difference filter parameterised by its moment matrix
    (rows a4-a6).  ``order=(o0, o1)``: derivative order along H and W."""

    def __init__(self, kernel_size, order, dx=1.0, constraint="moment", boundary="periodic"):
        super().__init__()
        if str(boundary).lower() != "periodic":
            raise NotImplementedError("only periodic boundaries are on the B200 hot path")
        k = kernel_size if isinstance(kernel_size, int) else kernel_size[0]
        if k % 2 == 0 or k > 7:
            raise ValueError("kernel_size must be odd and <= 7")
        self.kernel_size = k
        self.order = tuple(order)
        self.dx = _pair(dx)
        self.constraint = constraint
        m = torch.zeros(k, k, dtype=torch.float64)
        m[self.order[0], self.order[1]] = 1
        self.moment = nn.Parameter(m)
        self.register_buffer("_pinned", self._mask(), persistent=False)
        self.register_buffer("_init", m.clone(), persistent=False)

    def _mask(self) -> torch.Tensor:
        k = self.kernel_size
        mask = torch.zeros(k, k, dtype=torch.bool)
        c = self.constraint
        if c == "free" or c == 0:
            return mask
        if c == "frozen":
            return ~mask
        acc = 1 if c == "moment" else int(c)
        s = self.order[0] + self.order[1] + acc
        i = torch.arange(k)
        return (i[:, None] + i[None, :]) < s

    @property
    def scale(self) -> float:
        return 1.0 / (self.dx[0] ** self.order[0] * self.dx[1] ** self.order[1])

    @property
    def kernel(self) -> torch.Tensor:
        return ops.m2k(self.moment[None], [self.scale], self._pinned[None])[0]

    def x_proj(self, *args, **kw):
        """Re-impose the pinned moment entries (sum-rule constraint) after an optimiser update."""
        with torch.no_grad():  # in place: captured CUDA graphs keep reading the same storage
            self.moment.copy_(torch.where(self._pinned, self._init.to(self.moment.dtype), self.moment))

    def grad_proj(self, *args, **kw):
        """Zero the gradient of pinned moment entries."""
        if self.moment.grad is not None:
            self.moment.grad.data.masked_fill_(self._pinned, 0)

    def forward(self, u: torch.Tensor) -> torch.Tensor:
        """u [B,C,H,W] -> this derivative of every channel, [B,C,H,W]."""
        shape = u.shape
        u4 = u.reshape(-1, 1, shape[-2], shape[-1])
        return ops.fd2d(u4, self.kernel[None]).reshape(shape)


class poly(nn.Module):
    """SymNet (row a7): ``hidden_layers`` product units, then a linear read-out."""

    def __init__(self, hidden_layers, channel_num, channel_names=None):
        super().__init__()
        self.hidden_layers = hidden_layers
        self.channel_num = channel_num
        self.channel_names = list(channel_names) if channel_names is not None else [
            "u%d" % i for i in range(channel_num)]
        for i in range(hidden_layers):
            lin = nn.Linear(channel_num + i, 2).to(torch.float64)
            lin.weight.data.zero_()
            lin.bias.data.zero_()
            self.add_module("layer%d" % i, lin)
        fin = nn.Linear(channel_num + hidden_layers, 1).to(torch.float64)
        fin.weight.data.zero_()
        fin.bias.data.zero_()
        self.add_module("layer_final", fin)

    def flat_params(self) -> torch.Tensor:
        """The flat parameter vector in the C-ABI layout (W_i, b_i ..., w_f, b_f)."""
        parts = []
        for i in range(self.hidden_layers):
            lin = getattr(self, "layer%d" % i)
            parts += [lin.weight.reshape(-1), lin.bias.reshape(-1)]
        parts += [self.layer_final.weight.reshape(-1), self.layer_final.bias.reshape(-1)]
        return torch.cat(parts)

    def forward(self, inputs: torch.Tensor) -> torch.Tensor:
        """inputs [..., channel_num] -> [...]; standalone evaluation (the rollout uses the fused kernel)."""
        return ops.symnet_eval(inputs, self.flat_params(), self.hidden_layers)


class POLYPDE2D(nn.Module):
    """PDE-Net 2.0 model (rows a8-a10): ``forward(init, T)`` = state after ``round(T/dt)`` Euler steps.
    ``scheme``: ``'central'`` or ``'upwind'`` (pseudo-upwind first-order inputs, row f-1; supported for
    ``max_order`` 1-2 and up to two channels)."""

    #: the moment gradients this model produces are already zero on pinned entries (the mask is applied inside
    #: ``pdenet_m2k_bwd_*``, see ``_bank_kernels``): ``losses.GraphedTrainingLoss`` then skips ``diff_grad_proj``
    pinned_gradients_masked = True

    def __init__(self, dt, dx, kernel_size, max_order, constraint, channel_names, hidden_layers=2,
                 scheme="central", dtype=torch.float64, batch_chunk=0, keep_features=None):
        super().__init__()
        if scheme not in ("central", "upwind"):
            raise ValueError("scheme must be 'central' or 'upwind', got %r" % (scheme,))
        self.scheme = scheme
        self.dt, self.dx = float(dt), _pair(dx)
        self.kernel_size, self.max_order, self.constraint = kernel_size, max_order, constraint
        self.channel_names = list(channel_names.split(",")) if isinstance(channel_names, str) else list(channel_names)
        self.channel_num = len(self.channel_names)
        self.hidden_layers = hidden_layers
        self.batch_chunk = batch_chunk
        self.keep_features = keep_features  # see ops.rollout: None = keep when they fit the HBM budget
        self.orders = order_list(max_order)
        self._fd_names = []
        for (o0, o1) in self.orders:
            name = "fd%d%d" % (o0, o1)
            self.add_module(name, FD2d(kernel_size, (o0, o1), dx=dx, constraint=constraint))
            self._fd_names.append(name)
        allch = [c + "_%d%d" % o for c in self.channel_names for o in self.orders]
        for c in range(self.channel_num):
            self.add_module("poly%d" % c, poly(hidden_layers, len(allch), allch))
        self.to(dtype)

    # ---- parameter groups / projection hooks (upstream names)
    @property
    def fds(self):
        return [getattr(self, n) for n in self._fd_names]

    @property
    def polys(self):
        return [getattr(self, "poly%d" % c) for c in range(self.channel_num)]

    def diff_params(self):
        return (fd.moment for fd in self.fds)

    def expr_params(self):
        return (p for pl in self.polys for p in pl.parameters())

    def diff_x_proj(self, *a, **kw):
        for fd in self.fds:
            fd.x_proj()

    def diff_grad_proj(self, *a, **kw):
        for fd in self.fds:
            fd.grad_proj()

    # ---- pieces of the step
    def _attach_flat(self, fp) -> None:
        """Called by ``optim.FlatParams(model)``: parameters are now views of one flat buffer."""
        self.__dict__["_flat"] = fp

    def scaled_kernels(self) -> torch.Tensor:
        return _bank_kernels(self, self.fds)

    def symnet_flat(self) -> torch.Tensor:
        """[C, Psym] in the C-ABI layout.  With a ``FlatParams`` attached this is a zero-copy view of
        the flat parameter buffer (the nn.Linear parameters of ``poly0..`` are laid out in exactly
        that order) whose gradient goes straight into the flat gradient buffer."""
        fp = self.__dict__.get("_flat")
        if fp is not None:
            ps = [q for pl in self.polys for q in pl.parameters()]
            return fp.group(ps, (self.channel_num, sum(q.numel() for q in ps) // self.channel_num))
        return torch.stack([pl.flat_params() for pl in self.polys])

    def UInputs(self, u: torch.Tensor) -> torch.Tensor:
        """[B,C,H,W] -> channels-last derivative features [B,H,W,C*N] (feature index c*N + n)."""
        return ops.fd2d(u, self.scaled_kernels()).permute(0, 2, 3, 1)

    def RightHandItems(self, u: torch.Tensor) -> torch.Tensor:
        """F(u) [B,C,H,W]: the fused step kernel writing the SymNet output itself (no Euler update,
        no differencing of states)."""
        return ops.rhs(u, self.scaled_kernels(), self.symnet_flat(), self.hidden_layers, self.batch_chunk,
                       scheme=self.scheme)

    def predict(self, u: torch.Tensor) -> torch.Tensor:
        """One forward-Euler step ``u + dt*F(u)``."""
        return self.multistep(u, 1)

    def multistep(self, inputs: torch.Tensor, stepnum: int, all_states: bool = False) -> torch.Tensor:
        """``stepnum`` forward-Euler steps (a step COUNT, as upstream's ``multistep(inputs, stepnum)``;
        ``forward(init, T)`` is the entry that takes a time).  One fused CUDA rollout."""
        if int(stepnum) != stepnum or stepnum < 1:
            raise ValueError("stepnum must be a positive integer, got %r" % (stepnum,))
        return ops.rollout(inputs, self.scaled_kernels(), self.symnet_flat(), self.hidden_layers, self.dt,
                           int(stepnum), last_only=not all_states, batch_chunk=self.batch_chunk,
                           keep_features=self.keep_features, scheme=self.scheme)

    def forward(self, init: torch.Tensor, T: float, all_states: bool = False) -> torch.Tensor:
        """State after time ``T``; ``T`` must be a whole number of steps of ``dt`` (as upstream, which
        refuses anything else instead of rounding silently)."""
        stepnum = round(T / self.dt)
        if stepnum < 1 or abs(stepnum - T / self.dt) > 1e-5:
            raise NotImplementedError("T=%r is not a positive whole multiple of dt=%r" % (T, self.dt))
        return self.multistep(init, stepnum, all_states=all_states)


class LagrangeInterpFixInputs2d(nn.Module):
    """Coefficient fields on the solver grid as piecewise tensor-product Lagrange polynomials of a
    coarse mesh (upstream ``aTEAM.nn.modules.Interpolation.LagrangeInterpFixInputs``, SURVEY 8f-2;
    restated from its description: the upstream source is not in the mount).

    The domain ``mesh_bound = ((y0, x0), (y1, x1))`` is cut into ``mesh_size = (m0, m1)`` cells; each
    cell carries a degree-``degree`` Lagrange basis on equispaced nodes, so the trainable table
    ``interp_coe`` is ``[channels, m0*degree+1, m1*degree+1]``.  Because the evaluation points are
    the fixed grid nodes, the map table -> field is two small dense matrices (built once):
    ``field[c] = By @ interp_coe[c] @ Bx^T``.  It runs once per rollout, outside the step loop, as
    plain tensor ops (autograd carries ``grad_coe`` from the fused reverse rollout back to the
    table)."""

    def __init__(self, channels, grid, mesh_size=(5, 5), degree=2, mesh_bound=((0.0, 0.0), (2 * math.pi, 2 * math.pi)),
                 dtype=torch.float64):
        super().__init__()
        self.channels, self.grid = channels, _pair(grid)
        self.mesh_size, self.degree, self.mesh_bound = _pair(mesh_size), int(degree), mesh_bound
        H, W = self.grid
        (y0, x0), (y1, x1) = mesh_bound
        ys = y0 + (y1 - y0) * torch.arange(H, dtype=torch.float64) / H     # periodic grid: no end point
        xs = x0 + (x1 - x0) * torch.arange(W, dtype=torch.float64) / W
        self.register_buffer("By", self._basis(ys, y0, y1, self.mesh_size[0]).to(dtype), persistent=False)
        self.register_buffer("Bx", self._basis(xs, x0, x1, self.mesh_size[1]).to(dtype), persistent=False)
        self.interp_coe = nn.Parameter(torch.zeros(channels, self.mesh_size[0] * degree + 1,
                                                   self.mesh_size[1] * degree + 1, dtype=dtype))

    def _basis(self, pts: torch.Tensor, lo: float, hi: float, cells: int) -> torch.Tensor:
        d = self.degree
        n = cells * d + 1
        B = torch.zeros(pts.numel(), n, dtype=torch.float64)
        h = (hi - lo) / cells
        for r, p in enumerate(pts.tolist()):
            c = min(int((p - lo) / h), cells - 1)
            t = (p - lo) / h - c                      # local coordinate in [0, 1]
            nodes = [j / d for j in range(d + 1)]
            for j in range(d + 1):
                w = 1.0
                for q in range(d + 1):
                    if q != j:
                        w *= (t - nodes[q]) / (nodes[j] - nodes[q])
                B[r, c * d + j] = w
        return B

    def forward(self) -> torch.Tensor:
        """-> coefficient fields ``[channels, H, W]`` on the solver grid."""
        return self.By @ self.interp_coe @ self.Bx.t()


class VariantCoeLinear2d(nn.Module):
    """PDE-Net 1 linear variable-coefficient model (row a12):
    ``u+ = D0*u + dt * sum_n c_n(x,y) (D_n*u)`` with learnable D0 and coefficient fields c_n."""

    def __init__(self, kernel_size, max_order, dx, dt, mesh_size, constraint="moment", dtype=torch.float64,
                 interp_degree=None, interp_mesh_size=(5, 5), mesh_bound=((0.0, 0.0), (2 * math.pi, 2 * math.pi))):
        super().__init__()
        self.dt, self.dx = float(dt), _pair(dx)
        self.kernel_size, self.max_order = kernel_size, max_order
        self.mesh_size = _pair(mesh_size)
        self.orders = [o for o in order_list(max_order) if o != (0, 0)]
        self.id = FD2d(kernel_size, (0, 0), dx=dx, constraint=constraint)
        self._fd_names = []
        for (o0, o1) in self.orders:
            name = "fd%d%d" % (o0, o1)
            self.add_module(name, FD2d(kernel_size, (o0, o1), dx=dx, constraint=constraint))
            self._fd_names.append(name)
        # coefficient fields on the grid: direct parameters, or (interp_degree given, as upstream)
        # piecewise Lagrange polynomials of a coarse table evaluated once per call
        if interp_degree is None:
            self.interp = None
            self.coe = nn.Parameter(torch.zeros(len(self.orders), *self.mesh_size, dtype=torch.float64))
        else:
            self.interp = LagrangeInterpFixInputs2d(len(self.orders), self.mesh_size, interp_mesh_size,
                                                    interp_degree, mesh_bound)
        self.to(dtype)

    @property
    def fds(self):
        return [getattr(self, n) for n in self._fd_names]

    def diff_params(self):
        return (fd.moment for fd in [self.id] + self.fds)

    def coe_params(self):
        return iter([self.coe] if self.interp is None else [self.interp.interp_coe])

    def coefficient_fields(self) -> torch.Tensor:
        return self.coe if self.interp is None else self.interp()

    def diff_x_proj(self, *a, **kw):
        for fd in [self.id] + self.fds:
            fd.x_proj()

    def diff_grad_proj(self, *a, **kw):
        for fd in [self.id] + self.fds:
            fd.grad_proj()

    def _attach_flat(self, fp) -> None:
        self.__dict__["_flat"] = fp

    def scaled_kernels(self):
        return _bank_kernels(self, [self.id] + self.fds)

    def forward(self, init: torch.Tensor, stepnum: int, all_states: bool = False) -> torch.Tensor:
        ks = self.scaled_kernels()
        return ops.linear_rollout(init, ks[0], ks[1:], self.coefficient_fields(), self.dt, int(stepnum),
                                  last_only=not all_states)


class LagrangeInterp1d(nn.Module):
    """Learnable piecewise Lagrange polynomial of one variable on a uniform mesh (upstream
    ``aTEAM.nn.modules.Interpolation.LagrangeInterp`` with m = 1 [RECALL]): ``interp_coe``
    [mesh_size*degree + 1] are the node values, zero-initialised; ``forward(x)`` evaluates it at
    every entry of ``x`` with the CUDA op ``ops.lagrange1d``."""

    def __init__(self, degree: int, mesh_bound, mesh_size: int, dtype=torch.float64):
        super().__init__()
        self.degree, self.mesh_size = int(degree), int(mesh_size)
        self.mesh_bound = (float(mesh_bound[0]), float(mesh_bound[1]))
        self.interp_coe = nn.Parameter(torch.zeros(self.mesh_size * self.degree + 1, dtype=dtype))

    def nodes(self) -> torch.Tensor:
        """Positions of the interpolation nodes (for initialising ``interp_coe`` from a function)."""
        lo, hi = self.mesh_bound
        return torch.linspace(lo, hi, self.mesh_size * self.degree + 1, dtype=self.interp_coe.dtype,
                              device=self.interp_coe.device)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return ops.lagrange1d(x, self.interp_coe, self.mesh_bound[0], self.mesh_bound[1], self.mesh_size, self.degree)


class SingleNonLinear2d(nn.Module):
    """PDE-Net 1 model with a nonlinear source term (upstream branch ``PDE-Net``,
    ``SingleNonLinear2d`` [RECALL]):
    ``u+ = D0*u + dt * (sum_n c_n (D_n*u) + f(u))`` with learnable filters, one learnable constant
    coefficient c_n per derivative term and ``f`` a learnable piecewise Lagrange polynomial.

    On the grids the model is used on (50 x 50: whatever fits the cluster-resident small-grid path) the whole
    rollout is ONE fused op per direction (``ops.single_nonlinear_rollout``); larger grids step through the fused
    linear step (``ops.linear_rollout`` with the constants broadcast to fields) plus the interpolation op."""

    def __init__(self, kernel_size, max_order, dx, dt, constraint="moment", nonlinear_interp_degree=2,
                 nonlinear_interp_mesh_bound=(-15.0, 15.0), nonlinear_interp_mesh_size=20, dtype=torch.float64):
        super().__init__()
        self.dt, self.dx = float(dt), _pair(dx)
        self.kernel_size, self.max_order = kernel_size, max_order
        self.orders = [o for o in order_list(max_order) if o != (0, 0)]
        self.id = FD2d(kernel_size, (0, 0), dx=dx, constraint=constraint)
        self._fd_names = []
        for (o0, o1) in self.orders:
            name = "fd%d%d" % (o0, o1)
            self.add_module(name, FD2d(kernel_size, (o0, o1), dx=dx, constraint=constraint))
            self._fd_names.append(name)
        self.coe = nn.Parameter(torch.zeros(len(self.orders), dtype=torch.float64))
        self.nonlinearfitter = LagrangeInterp1d(nonlinear_interp_degree, nonlinear_interp_mesh_bound,
                                                nonlinear_interp_mesh_size)
        self.to(dtype)

    fds = VariantCoeLinear2d.fds
    diff_x_proj = VariantCoeLinear2d.diff_x_proj
    diff_grad_proj = VariantCoeLinear2d.diff_grad_proj
    scaled_kernels = VariantCoeLinear2d.scaled_kernels
    _attach_flat = VariantCoeLinear2d._attach_flat

    def diff_params(self):
        return (fd.moment for fd in [self.id] + self.fds)

    def coe_params(self):
        return iter([self.coe, self.nonlinearfitter.interp_coe])

    def forward(self, init: torch.Tensor, stepnum: int, fused=None) -> torch.Tensor:
        """``fused=None``: the whole rollout as ONE op (``ops.single_nonlinear_rollout``: one launch per direction,
        trajectory resident in shared memory) when the grid fits the small-grid path, else a per-step loop of the
        fused linear step + the interpolation op (larger grids)."""
        ks = self.scaled_kernels()
        B, H, W = init.shape
        nf = self.nonlinearfitter
        if fused is None:
            fused = ops.single_nonlinear_fits(init, ks[1:])
        if fused:
            return ops.single_nonlinear_rollout(init, ks[0], ks[1:], self.coe, nf.interp_coe, self.dt, int(stepnum),
                                                nf.mesh_bound[0], nf.mesh_bound[1], nf.mesh_size, nf.degree)
        fields = self.coe[:, None, None].expand(-1, H, W).contiguous()
        u = init
        for _ in range(int(stepnum)):
            u = ops.linear_rollout(u, ks[0], ks[1:], fields, self.dt, 1, last_only=True) + self.dt * nf(u)
        return u
