"""Flat parameter / gradient view for SciPy-style optimisers.

Stands in for the part of upstream ``aTEAM.optim.NumpyFunctionInterface`` that the training script
uses around the model: a single flat vector of all trainable parameters, ``x_proj`` applied after
every write (pinned moment entries re-imposed) and ``grad_proj`` applied before every gradient
read (their gradients zeroed).  Everything stays on the model's device until the final
``.cpu().numpy()``; one D2H copy per evaluation instead of one per parameter tensor."""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np
import torch


class FlatParams:
    def __init__(self, model: torch.nn.Module, x_proj: Optional[Callable] = None,
                 grad_proj: Optional[Callable] = None):
        self.model = model
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.x_proj = x_proj if x_proj is not None else getattr(model, "diff_x_proj", None)
        self.grad_proj = grad_proj if grad_proj is not None else getattr(model, "diff_grad_proj", None)
        self.numel = sum(p.numel() for p in self.params)

    # ---- parameters
    def get_flat(self) -> torch.Tensor:
        return torch.cat([p.detach().reshape(-1) for p in self.params])

    def set_flat(self, x) -> None:
        ref = self.params[0]
        x = torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x).to(ref.device, ref.dtype)
        if x.numel() != self.numel:
            raise ValueError("flat vector has %d entries, expected %d" % (x.numel(), self.numel))
        o = 0
        with torch.no_grad():
            for p in self.params:
                p.copy_(x[o:o + p.numel()].reshape(p.shape))
                o += p.numel()
        if self.x_proj is not None:
            self.x_proj()

    # ---- gradients
    def grad_flat(self) -> torch.Tensor:
        if self.grad_proj is not None:
            self.grad_proj()
        return torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in self.params])

    def zero_grad(self) -> None:
        for p in self.params:
            p.grad = None

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
