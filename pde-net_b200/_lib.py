"""ctypes binding of libpdenet_b200.so (the C-ABI declared in include/pdenet_b200.h).

The product path has no CPU fallback: if the shared library is missing (or a call returns a
negative code) this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpdenet_b200.so")
TORCH_LIB_PATH = os.path.join(_HERE, "libpdenet_b200_torch.so")

OK, E_BADARG, E_UNSUPPORTED, E_WORKSPACE, E_CUDA = 0, -1, -2, -3, -4
GRAD_LAST, GRAD_ALL = 0, 1
MAX_K, MAX_FILTERS, MAX_LAYERS = 7, 16, 8

_vp, _i, _d, _sz = C.c_void_p, C.c_int, C.c_double, C.c_size_t

# name -> (restype, argtypes); "{t}" expands to f64 / f32
_PROTOS = {
    "pdenet_version": (_i, []),
    "pdenet_build_id": (C.c_char_p, []),
    "pdenet_error_string": (C.c_char_p, [_i]),
    "pdenet_supported": (_i, [_i] * 5),
    "pdenet_symnet_param_count": (_i, [_i, _i]),
    "pdenet_m2k_{t}": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "pdenet_k2m_{t}": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "pdenet_m2k_bwd_{t}": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp]),
    "pdenet_fd2d_fwd_{t}": (_i, [_vp, _vp, _vp] + [_i] * 6 + [_vp]),
    "pdenet_fd2d_bwd_workspace_bytes": (_sz, [_i, _i]),
    "pdenet_fd2d_bwd_{t}": (_i, [_vp] * 6 + [_sz] + [_i] * 6 + [_vp]),
    "pdenet_rollout_fwd_{t}": (_i, [_vp] * 4 + [_i] * 8 + [_d, _vp]),
    "pdenet_rollout_bwd_needs_features": (_i, [_i] * 5),
    "pdenet_rhs_fwd_{t}": (_i, [_vp] * 4 + [_i] * 7 + [_vp]),
    "pdenet_rhs_bwd_{t}": (_i, [_vp] * 8 + [_sz] + [_i] * 8 + [_vp]),
    "pdenet_upwind_supported": (_i, [_i] * 5),
    "pdenet_rollout_fwd_upwind_{t}": (_i, [_vp] * 4 + [_i] * 8 + [_d, _vp]),
    "pdenet_rollout_bwd_upwind_workspace_bytes": (_sz, [_i] * 9),
    "pdenet_rollout_bwd_upwind_{t}": (_i, [_vp, _vp, _vp, _i] + [_vp] * 6 + [_sz] + [_i] * 9 + [_d, _vp]),
    "pdenet_rhs_upwind_fwd_{t}": (_i, [_vp] * 4 + [_i] * 7 + [_vp]),
    "pdenet_rhs_upwind_bwd_{t}": (_i, [_vp] * 8 + [_sz] + [_i] * 8 + [_vp]),
    "pdenet_rollout_fwd_keep_{t}": (_i, [_vp] * 5 + [_i] * 8 + [_d, _vp]),
    "pdenet_rollout_bwd_kept_{t}": (_i, [_vp, _vp, _vp, _vp, _i] + [_vp] * 6 + [_sz] + [_i] * 9 + [_d, _vp]),
    "pdenet_rollout_bwd_workspace_bytes": (_sz, [_i] * 9),
    "pdenet_rollout_bwd_{t}": (_i, [_vp, _vp, _vp, _i] + [_vp] * 6 + [_sz] + [_i] * 9 + [_d, _vp]),
    "pdenet_linear_rollout_fwd_{t}": (_i, [_vp] * 5 + [_i] * 6 + [_d, _vp]),
    "pdenet_linear_small_grid_path": (_i, [_i]),
    "pdenet_single_nonlinear_fits": (_i, [_i] * 6),
    "pdenet_single_nonlinear_rollout_fwd_{t}": (_i, [_vp] * 6 + [_i] * 6 + [_d, _d, _d, _i, _i, _vp]),
    "pdenet_single_nonlinear_rollout_bwd_workspace_bytes": (_sz, [_i] * 9),
    "pdenet_single_nonlinear_rollout_bwd_{t}": (_i, [_vp, _vp, _vp, _i] + [_vp] * 10 + [_sz] + [_i] * 6
                                                + [_d, _d, _d, _i, _i, _vp]),
    "pdenet_small_tile_factor": (_i, [_i]),
    "pdenet_filter_grad_tile_level": (_i, [_i]),
    "pdenet_linear_rollout_bwd_workspace_bytes": (_sz, [_i] * 6),
    "pdenet_linear_rollout_bwd_{t}": (_i, [_vp, _vp, _vp, _i] + [_vp] * 8 + [_sz] + [_i] * 6 + [_d, _vp]),
    "pdenet_slab_step_fwd_{t}": (_i, [_vp] * 4 + [_i] * 10 + [_d, _vp]),
    "pdenet_slab_step_fwd_p2p_{t}": (_i, [_vp] * 6 + [_i] * 11 + [_d, _vp]),
    "pdenet_slab_step_bwd_p2p_{t}": (_i, [_vp] * 7 + [_i, _i, _i, _vp, _sz] + [_i] * 7 + [_d, _vp]),
    "pdenet_slab_step_fwd_keep_{t}": (_i, [_vp] * 7 + [_i] * 11 + [_d, _vp]),
    "pdenet_slab_step_bwd_kept_{t}": (_i, [_vp] * 8 + [_i, _i, _i, _vp, _sz] + [_i] * 7 + [_d, _vp]),
    "pdenet_slab_bwd_workspace_bytes": (_sz, [_i] * 8),
    "pdenet_lagrange1d_fwd_{t}": (_i, [_vp, _vp, _vp, _sz, _d, _d, _i, _i, _vp]),
    "pdenet_lagrange1d_bwd_workspace_bytes": (_sz, [_sz, _i, _i]),
    "pdenet_lagrange1d_bwd_{t}": (_i, [_vp] * 5 + [_sz, _d, _d, _i, _i, _vp, _sz, _vp]),
    "pdenet_slab_bwd_begin": (_i, [_vp, _sz] + [_i] * 8 + [_vp]),
    "pdenet_slab_step_bwd_{t}": (_i, [_vp] * 5 + [_i, _i, _vp, _sz] + [_i] * 7 + [_d, _vp]),
    "pdenet_slab_bwd_end_{t}": (_i, [_vp, _sz, _vp, _vp] + [_i] * 7 + [_vp]),
    "pdenet_mse_seed_{t}": (_i, [_vp] * 5 + [_sz, _i, _vp, _sz, _vp]),
    "pdenet_profile_enable": (_i, [_i]),
    "pdenet_profile_read": (_i, [C.POINTER(_d), C.POINTER(C.c_longlong)]),
    "pdenet_fma_peak": (_i, [_i, _i, C.POINTER(_d), C.POINTER(C.c_float), _vp]),
    "pdenet_fma_probe": (_i, [_i, _i, _i, C.POINTER(_d), C.POINTER(C.c_float), _vp]),
}


def exported_symbols():
    """Every symbol include/pdenet_b200.h declares."""
    out = []
    for name in _PROTOS:
        if "{t}" in name:
            out += [name.format(t="f64"), name.format(t="f32")]
        else:
            out.append(name)
    return out


_lib = None


#: C-ABI revision this binding was written against (``pdenet_version()``); bumped with every
#: signature change so that a stale binary is refused instead of being called with shifted arguments
ABI_VERSION = 200


def _build_module():
    import importlib.util
    spec = importlib.util.spec_from_file_location("pdenet_b200_build", os.path.join(_HERE, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def lib() -> C.CDLL:
    """Load (building or re-building first if the binary is missing or stale) and bind the library."""
    global _lib
    if _lib is None:
        build = _build_module()
        if not build.up_to_date():
            # not a fallback: compile the CUDA library in place (nvcc, sm_100a) or fail loudly
            try:
                build.build()
            except Exception as e:  # noqa: BLE001
                raise RuntimeError(
                    "pdenet_b200: %s is missing or stale and could not be built (%s) -- build it with "
                    "`python pde-net_b200/build.py` (there is no CPU fallback)" % (LIB_PATH, e)) from e
        l = C.CDLL(LIB_PATH)
        l.pdenet_version.restype = _i
        l.pdenet_build_id.restype = C.c_char_p
        if l.pdenet_version() != ABI_VERSION or l.pdenet_build_id().decode() != build.source_hash():
            raise RuntimeError("pdenet_b200: %s was built from other sources (ABI %d, build id %s; expected %d, %s) "
                               "-- rebuild with `python pde-net_b200/build.py --force`"
                               % (LIB_PATH, l.pdenet_version(), l.pdenet_build_id().decode(), ABI_VERSION,
                                  build.source_hash()))
        for name, (res, args) in _PROTOS.items():
            for n in ([name.format(t="f64"), name.format(t="f32")] if "{t}" in name else [name]):
                fn = getattr(l, n)
                fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


_torch_ops = False   # False = not tried yet; None = ctypes binding selected / shim unavailable


def torch_ops():
    """``torch.ops.pdenet_b200`` -- the hot-path entries registered as dispatcher ops by the TORCH_LIBRARY shim
    (csrc/torch_shim.cpp -> libpdenet_b200_torch.so, a torch C++ extension over the same C-ABI) -- or None when
    ``PDENET_BINDING=ctypes`` selects the plain ctypes binding (option B).  Both bindings call
    the same ``extern "C"`` entries of libpdenet_b200.so."""
    global _torch_ops
    if _torch_ops is False:
        _torch_ops = None
        if os.environ.get("PDENET_BINDING", "torch").lower() != "ctypes":
            lib()   # builds / verifies both shared objects
            import torch
            try:
                torch.ops.load_library(TORCH_LIB_PATH)
                _torch_ops = torch.ops.pdenet_b200
            except OSError as e:   # e.g. a torch build with another C++ ABI: same C-ABI through ctypes instead
                import warnings
                warnings.warn("pdenet_b200: the TORCH_LIBRARY shim could not be loaded (%s); using the ctypes binding "
                              "of the same C-ABI" % e)
    return _torch_ops


def call_op(what: str, fn, *args):
    """Call a dispatcher op of the shim; a C-ABI error code comes back as the same :class:`PdenetError` the
    ctypes binding raises."""
    try:
        return fn(*args)
    except RuntimeError as e:
        import re
        m = re.search(r"\(code (-?\d+)\)", str(e))
        if m:
            raise PdenetError(int(m.group(1)), what) from None
        raise


class PdenetError(RuntimeError):
    def __init__(self, code, what):
        self.code = code
        msg = lib().pdenet_error_string(code).decode()
        super().__init__("%s failed: %s (code %d)" % (what, msg, code))


def check(code: int, what: str) -> None:
    if code != 0:
        raise PdenetError(code, what)


def fn(name: str, dtype) -> C._CFuncPtr:
    """Typed entry point for a torch dtype: fn('pdenet_rollout_fwd', torch.float64)."""
    import torch
    if dtype == torch.float64:
        return getattr(lib(), name + "_f64")
    if dtype == torch.float32:
        return getattr(lib(), name + "_f32")
    raise TypeError("pdenet_b200 supports float64 and float32, got %s" % dtype)
