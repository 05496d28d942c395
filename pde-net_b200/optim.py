"""Flat parameter / gradient storage for SciPy-style optimisers and data-parallel training.

Stands in for the part of upstream ``aTEAM.optim.NumpyFunctionInterface`` that the training script
uses around the model [RECALL -- the mounted reference holds no source]: a single flat vector of
all trainable parameters, ``x_proj`` applied after every write (pinned moment entries re-imposed)
and ``grad_proj`` applied before every gradient read (their gradients zeroed).

B200-first layout: :class:`FlatParams` OWNS one persistent flat parameter buffer and one persistent
flat gradient buffer on the model's device.  Every ``nn.Parameter`` of the model becomes a VIEW
into the parameter buffer and its ``.grad`` a view into the gradient buffer, so

* ``set_flat`` / ``get_flat`` / ``grad_flat`` are one copy (or none), not one per tensor;
* the fused rollout's parameter gradients land in the gradient buffer with ONE accumulate per
  group (moments, SymNet) through :class:`_FlatGroup` -- no ``torch.cat`` / split / per-parameter
  ``AccumulateGrad``;
* the data-parallel exchange is ONE in-place all-reduce of the gradient buffer (the loss rides in
  its last slot) with no pack / unpack (:meth:`FlatParams.allreduce`);
* the buffers never move, so a CUDA graph captured around an evaluation stays valid across
  optimiser updates.
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch


class _Count(int):
    """An int that is also callable: ``fp.numel`` and upstream-style ``nfi.numel()`` both work."""

    def __call__(self) -> int:
        return int(self)


class _FlatGroup(torch.autograd.Function):
    """Zero-copy bridge between a run of consecutive parameters and the op that consumes them as one
    tensor.  forward: the run's slice of the flat parameter buffer, reshaped (the parameters ARE
    views of it: no ``cat``).  backward: ONE in-place accumulate of the incoming gradient into the
    same slice of the flat gradient buffer; the parameters' ``.grad`` views see it, autograd's
    per-parameter accumulation is skipped."""

    @staticmethod
    def forward(ctx, owner, start, shape, *params):
        n = 1
        for s in shape:
            n *= s
        ctx.owner, ctx.start, ctx.n, ctx.np = owner, start, n, len(params)
        return owner.flat[start:start + n].view(shape)

    @staticmethod
    def backward(ctx, g):
        ctx.owner.gflat[ctx.start:ctx.start + ctx.n].add_(g.reshape(-1))
        return (None, None, None) + (None,) * ctx.np


class FlatParams:
    """One flat parameter buffer + one flat gradient buffer behind a model's parameters.

    ``flat`` [numel] and ``gflat`` [numel + 1] (the extra slot carries the loss through the
    data-parallel all-reduce) are allocated once; ``p.data`` / ``p.grad`` of every trainable
    parameter are re-pointed to views of them, in ``model.parameters()`` order.  Do not move the
    model (``.to`` / ``.cuda`` / ``.double``) afterwards: that would detach the views
    (:meth:`check` raises if it happened)."""

    def __init__(self, model: torch.nn.Module, x_proj: Optional[Callable] = None,
                 grad_proj: Optional[Callable] = None):
        self.model = model
        self.params = [p for p in model.parameters() if p.requires_grad]
        if not self.params:
            raise ValueError("FlatParams: the model has no trainable parameters")
        ref = self.params[0]
        if any(p.dtype != ref.dtype or p.device != ref.device for p in self.params):
            raise ValueError("FlatParams: all parameters must share one dtype and device")
        self.x_proj = x_proj if x_proj is not None else getattr(model, "diff_x_proj", None)
        self.grad_proj = grad_proj if grad_proj is not None else getattr(model, "diff_grad_proj", None)
        self.numel = _Count(sum(p.numel() for p in self.params))
        self.flat = torch.empty(self.numel, dtype=ref.dtype, device=ref.device)
        self._gbuf = torch.zeros(self.numel + 1, dtype=ref.dtype, device=ref.device)
        self.gflat = self._gbuf[:self.numel]
        self.offsets: Dict[int, int] = {}
        o = 0
        with torch.no_grad():
            for p in self.params:
                n = p.numel()
                self.flat[o:o + n].copy_(p.detach().reshape(-1))
                p.data = self.flat[o:o + n].view(p.shape)
                p.grad = self.gflat[o:o + n].view(p.shape)
                self.offsets[id(p)] = o
                o += n
        if hasattr(model, "_attach_flat"):
            model._attach_flat(self)

    # ---- integrity
    def check(self) -> None:
        """Raise if a parameter or gradient no longer aliases the flat buffers."""
        es = self.flat.element_size()
        for p in self.params:
            o = self.offsets[id(p)]
            if p.data_ptr() != self.flat.data_ptr() + o * es or p.grad is None or \
                    p.grad.data_ptr() != self.gflat.data_ptr() + o * es:
                raise RuntimeError("FlatParams: a parameter (or its .grad) was re-allocated after the flat buffers "
                                   "were attached (model.to()/zero_grad(set_to_none=True)?); build a new FlatParams")

    def group(self, params: Sequence[torch.nn.Parameter], shape: Sequence[int]) -> torch.Tensor:
        """The consecutive run ``params`` as ONE tensor of ``shape`` (zero-copy; gradients accumulate
        straight into the flat gradient buffer)."""
        start = self.offsets[id(params[0])]
        o = start
        for p in params:
            if self.offsets.get(id(p)) != o:
                raise ValueError("FlatParams.group: the parameters are not consecutive in the flat buffer")
            o += p.numel()
        if not torch.is_grad_enabled():
            n = o - start
            return self.flat[start:start + n].view(tuple(shape))
        return _FlatGroup.apply(self, start, tuple(shape), *params)

    # ---- parameters
    def get_flat(self) -> torch.Tensor:
        return self.flat.detach().clone()

    def set_flat(self, x) -> None:
        x = torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x)
        if x.numel() != self.numel:
            raise ValueError("flat vector has %d entries, expected %d" % (x.numel(), self.numel))
        with torch.no_grad():
            self.flat.copy_(x.reshape(-1))          # one (H2D) copy; dtype conversion included
        if self.x_proj is not None:
            self.x_proj()

    # ---- gradients
    def grad_flat(self) -> torch.Tensor:
        """The flat gradient (``grad_proj`` applied).  A VIEW of the persistent buffer: clone it to
        keep it across the next ``zero_grad``."""
        if self.grad_proj is not None:
            self.grad_proj()
        return self.gflat

    @property
    def grad_and_loss(self) -> torch.Tensor:
        """The whole gradient buffer ``[numel + 1]``: flat gradient, then the loss slot that
        :meth:`allreduce` fills -- one D2H copy brings back everything an optimiser step needs."""
        return self._gbuf

    def zero_grad(self) -> None:
        self._gbuf.zero_()                          # one memset; the .grad views stay attached

    def allreduce(self, loss: Optional[torch.Tensor] = None, group=None) -> Optional[torch.Tensor]:
        """Data-parallel gradient exchange: ONE in-place all-reduce (sum) of the gradient buffer, the
        loss riding in its last slot.  No packing, no unpacking.  Returns the reduced loss."""
        import torch.distributed as dist
        if loss is not None:
            self._gbuf[self.numel:].copy_(loss.detach().reshape(1))
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self._gbuf if loss is not None else self.gflat, group=group)
        return self._gbuf[self.numel] if loss is not None else None

    # ---- scipy.optimize.fmin_l_bfgs_b-style callables
    def make_f_fprime(self, loss_closure: Callable[[], torch.Tensor]):
        """``loss_closure()`` returns the scalar loss (graph attached).  Returns ``(f, fprime)``
        taking / returning float64 numpy vectors; the pair shares one evaluation per ``x``."""
        cache = {"x": None, "f": None, "g": None}

        def evaluate(x):
            if cache["x"] is None or not np.array_equal(cache["x"], x):
                self.set_flat(x)
                self.zero_grad()
                loss = loss_closure()
                loss.backward()
                cache["x"] = np.array(x, copy=True)
                cache["f"] = float(loss.detach())
                cache["g"] = self.grad_flat().double().cpu().numpy()
            return cache

        return (lambda x: evaluate(x)["f"]), (lambda x: evaluate(x)["g"])


class NumpyFunctionInterface(FlatParams):
    """Name-compatible stand-in for upstream ``aTEAM.optim.NumpyFunctionInterface`` [RECALL]:

        nfi = NumpyFunctionInterface(model_or_param_groups, forward=closure)
        x, f, d = scipy.optimize.fmin_l_bfgs_b(nfi.f, nfi.flat_param, nfi.fprime, ...)

    ``params`` is a module, an iterable of parameters, or a list of group dicts
    ``{'params': [...], 'x_proj': fn, 'grad_proj': fn, 'isfreeze': bool}`` as upstream's
    ``ParamGroupsManager`` takes; frozen groups are left out of the flat vector.  ``forward`` is the
    loss closure.  ``flat_param`` (numpy, get/set), ``f(x)``, ``fprime(x)``, ``set_forward`` and
    ``numel()`` (also usable as the int ``numel``) follow the upstream names; storage is the flat device buffers of :class:`FlatParams`."""

    def __init__(self, params, forward: Optional[Callable[[], torch.Tensor]] = None, isfreeze: bool = False,
                 x_proj: Optional[Callable] = None, grad_proj: Optional[Callable] = None, always_refresh: bool = True,
                 **kw):
        self._group_x_proj: List[Callable] = []
        self._group_grad_proj: List[Callable] = []
        if isinstance(params, torch.nn.Module):
            model = params
        else:
            plist: List[torch.nn.Parameter] = []
            items = list(params)
            if items and isinstance(items[0], dict):
                for g in items:
                    if g.get("isfreeze", isfreeze):
                        continue
                    gp = list(g["params"])
                    plist += gp
                    if g.get("x_proj") is not None:
                        self._group_x_proj.append((g["x_proj"], gp))
                    if g.get("grad_proj") is not None:
                        self._group_grad_proj.append((g["grad_proj"], gp))
            else:
                plist = [] if isfreeze else items
            model = torch.nn.Module()
            model._flat_list = torch.nn.ParameterList(plist)
        self._forward = forward
        self.always_refresh = always_refresh
        self._user_x_proj, self._user_grad_proj = x_proj, grad_proj
        super().__init__(model, x_proj=self._all_x_proj, grad_proj=self._all_grad_proj)
        self._model_x_proj = getattr(params, "diff_x_proj", None) if isinstance(params, torch.nn.Module) else None
        self._model_grad_proj = getattr(params, "diff_grad_proj", None) if isinstance(params, torch.nn.Module) else None
        self._cache = {"x": None, "f": None, "g": None}

    def _all_x_proj(self):
        for fn, gp in self._group_x_proj:
            fn(gp)
        if self._model_x_proj is not None:
            self._model_x_proj()
        if self._user_x_proj is not None:
            self._user_x_proj(self.params)

    def _all_grad_proj(self):
        for fn, gp in self._group_grad_proj:
            fn(gp)
        if self._model_grad_proj is not None:
            self._model_grad_proj()
        if self._user_grad_proj is not None:
            self._user_grad_proj(self.params)

    def set_forward(self, forward: Callable[[], torch.Tensor]) -> None:
        self._forward = forward
        self._cache["x"] = None

    @property
    def flat_param(self) -> np.ndarray:
        return self.flat.detach().double().cpu().numpy()

    @flat_param.setter
    def flat_param(self, x) -> None:
        self.set_flat(x)
        self._cache["x"] = None

    def _evaluate(self, x):
        c = self._cache
        x = np.asarray(x, dtype=np.float64)
        if self.always_refresh or c["x"] is None or not np.array_equal(c["x"], x):
            if self._forward is None:
                raise RuntimeError("NumpyFunctionInterface: no forward closure set")
            self.set_flat(x)
            self.zero_grad()
            loss = self._forward()
            loss.backward()
            c["x"], c["f"] = np.array(x, copy=True), float(loss.detach())
            c["g"] = self.grad_flat().double().cpu().numpy().copy()
        return c

    def f(self, x, *args, **kw) -> float:
        refresh, self.always_refresh = self.always_refresh, False   # f and fprime at the same x share one pass
        try:
            return self._evaluate(x)["f"]
        finally:
            self.always_refresh = refresh

    def fprime(self, x, always_double: bool = True, *args, **kw) -> np.ndarray:
        refresh, self.always_refresh = self.always_refresh, False
        try:
            return self._evaluate(x)["g"]
        finally:
            self.always_refresh = refresh

