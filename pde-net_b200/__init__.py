"""pdenet_b200 -- B200-native PDE-Net / PDE-Net 2.0 time-stepping hot path.

Importable as ``pdenet_b200`` (see the shim ``pdenet_b200.py`` at the repository root; the
directory itself is called ``pde-net_b200``).  Hand-written CUDA for sm_100a behind a C-ABI
(``include/pdenet_b200.h``), bound with ctypes and wired into torch.autograd.
"""
from . import _lib, losses, ops, optim, parallel
from .optim import FlatParams, NumpyFunctionInterface
from .modules import (FD2d, K2M, LagrangeInterp1d, LagrangeInterpFixInputs2d, M2K, POLYPDE2D, SingleNonLinear2d,
                      VariantCoeLinear2d, order_list, poly)
from .ops import (fd2d, fd2d_features, fma_peak, k2m, lagrange1d, linear_rollout, m2k, mse_loss, rhs, rollout, single_nonlinear_rollout, slab_step, symnet_eval,
                  symnet_param_count)

__all__ = ["losses", "LagrangeInterp1d", "SingleNonLinear2d", "lagrange1d", "LagrangeInterpFixInputs2d", "FlatParams", "NumpyFunctionInterface", "optim", "FD2d", "K2M", "M2K", "POLYPDE2D", "VariantCoeLinear2d", "order_list", "poly", "fd2d",
           "fd2d_features", "fma_peak", "k2m", "linear_rollout", "m2k", "mse_loss", "rhs", "rollout", "single_nonlinear_rollout", "slab_step", "symnet_eval",
           "symnet_param_count", "ops", "parallel"]
