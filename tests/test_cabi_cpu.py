"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol the header
declares; host-side module logic (constraint masks, parameter layout, projection hooks)."""
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as G
    G.build()
    import pdenet_b200
    return pdenet_b200


def test_header_symbols_exported(built):
    from pdenet_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "pdenet_b200.h")).read()
    declared = set(re.findall(r"PDENET_API\s+[\w\s\*]*?\b(pdenet_\w+)\s*\(", hdr))
    assert declared, "no declarations found"
    assert declared == set(_lib.exported_symbols())
    nm = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (pdenet_\w+)", nm))
    assert declared <= exported, declared - exported
    lib = _lib.lib()  # loads with ctypes; no compute call without a GPU
    assert lib.pdenet_version() >= 100
    assert lib.pdenet_symnet_param_count(12, 2) == 69
    assert lib.pdenet_error_string(-2).decode().startswith("unsupported")


def test_sass_is_sm100a_cuda_core_code(built):
    from pdenet_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback(built):
    with pytest.raises(RuntimeError):
        built.fd2d_features(torch.zeros(1, 2, 8, 8, dtype=torch.float64), torch.zeros(6, 5, 5, dtype=torch.float64))
    with pytest.raises(RuntimeError):
        built.rollout(torch.zeros(1, 2, 8, 8, dtype=torch.float64), torch.zeros(6, 5, 5, dtype=torch.float64),
                      torch.zeros(2, 69, dtype=torch.float64), 2, 1e-2, 1)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "pde-net_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no CPU fallback", ""), f


def test_module_layout_and_masks(built):
    from oracle import pdenet_oracle as O
    m = built.POLYPDE2D(1e-2, 0.1, 7, 2, "moment", "u,v", hidden_layers=2)
    assert [fd.order for fd in m.fds] == O.order_list(2)
    for fd in m.fds:
        assert torch.equal(fd._pinned, O.pinned_mask(7, fd.order, "moment"))
    assert m.symnet_flat().shape == (2, O.symnet_param_count(12, 2))
    # flat layout == oracle's symnet_unpack
    with torch.no_grad():
        for i, p in enumerate(m.poly0.parameters()):
            p.copy_(torch.arange(p.numel(), dtype=torch.float64).reshape(p.shape) + 100 * i)
    hidden, (wf, bf) = O.symnet_unpack(m.poly0.flat_params(), 12, 2)
    assert torch.equal(hidden[0][0], m.poly0.layer0.weight) and torch.equal(hidden[1][1], m.poly0.layer1.bias)
    assert torch.equal(wf, m.poly0.layer_final.weight[0]) and torch.equal(bf, m.poly0.layer_final.bias)
    # standalone SymNet evaluation (cold path) == oracle
    x = torch.randn(4, 5, 12, dtype=torch.float64)
    assert torch.allclose(m.poly0(x), O.symnet_apply(x, m.poly0.flat_params(), 2), atol=1e-12)
    # x_proj restores pinned entries, grad_proj zeroes their gradient
    fd = m.fd11
    with torch.no_grad():
        fd.moment.add_(1.0)
    fd.x_proj()
    assert torch.equal(fd.moment.detach()[fd._pinned], fd._init[fd._pinned])
    fd.moment.grad = torch.ones_like(fd.moment)
    fd.grad_proj()
    assert (fd.moment.grad[fd._pinned] == 0).all() and (fd.moment.grad[~fd._pinned] == 1).all()
    assert len(list(m.diff_params())) == 6 and len(list(m.expr_params())) == 12


def test_flat_params_view(built):
    """FlatParams (8f-4): flat get/set round trip, x_proj after a write, grad_proj before a read."""
    m = built.POLYPDE2D(1e-2, 0.1, 5, 2, "moment", "u,v", hidden_layers=2)
    fp = built.FlatParams(m)
    assert fp.numel == sum(p.numel() for p in m.parameters())
    x = torch.arange(fp.numel, dtype=torch.float64) * 1e-3 + 5.0
    fp.set_flat(x.numpy())
    got = fp.get_flat()
    # pinned moment entries were re-imposed (one-hot / zero), everything else took the new value
    o = 0
    for fd in m.fds:
        blk = got[o:o + 25].reshape(5, 5)
        assert torch.equal(blk[fd._pinned], fd._init[fd._pinned])
        assert torch.equal(blk[~fd._pinned], x[o:o + 25].reshape(5, 5)[~fd._pinned])
        o += 25
    assert torch.equal(got[o:], x[o:])
    # parameters and gradients are views of the two persistent flat buffers
    fp.check()
    es = fp.flat.element_size()
    for p in m.parameters():
        assert p.data_ptr() == fp.flat.data_ptr() + fp.offsets[id(p)] * es
        p.grad.fill_(1)                  # in place: the .grad views stay attached
    g = fp.grad_flat()
    assert g.data_ptr() == fp.gflat.data_ptr()
    o = 0
    for fd in m.fds:
        assert (g[o:o + 25].reshape(5, 5)[fd._pinned] == 0).all()
        o += 25
    assert (g[o:] == 1).all()
    fp.zero_grad()
    assert all((p.grad == 0).all() for p in m.parameters())
    fp.check()
    # the zero-copy groups the fused rollout consumes: moments [N,k,k] and SymNet [C,Psym]
    sym = m.symnet_flat()
    assert sym.data_ptr() == next(m.poly0.parameters()).data_ptr() and sym.shape == (2, 69)
    assert torch.equal(sym.detach(), torch.stack([pl.flat_params() for pl in m.polys]).detach())
    (sym * 2.0).sum().backward()         # one accumulate into the flat gradient buffer
    assert (fp.gflat[6 * 25:] == 2).all() and (fp.gflat[:6 * 25] == 0).all()
    # re-allocating a parameter behind FlatParams' back is detected
    m.fd00.moment.grad = None
    with pytest.raises(RuntimeError):
        fp.check()


def test_numpy_function_interface_lbfgs(built):
    """NumpyFunctionInterface (upstream aTEAM.optim names): flat_param / f / fprime drive SciPy's
    L-BFGS on a small quadratic; param-group x_proj / grad_proj hooks are applied."""
    import numpy as np
    from scipy.optimize import fmin_l_bfgs_b
    a = torch.nn.Parameter(torch.zeros(3, dtype=torch.float64))
    b = torch.nn.Parameter(torch.ones(2, 2, dtype=torch.float64))
    ta, tb = torch.tensor([1.0, -2.0, 3.0], dtype=torch.float64), torch.full((2, 2), 0.5, dtype=torch.float64)

    def pin_first(params):               # keep b[0,0] at 1 (an x_proj), and never move it (grad_proj)
        with torch.no_grad():
            params[0][0, 0] = 1.0

    def zero_first(params):
        params[0].grad[0, 0] = 0.0

    nfi = built.NumpyFunctionInterface(
        [dict(params=[a]), dict(params=[b], x_proj=pin_first, grad_proj=zero_first)],
        forward=lambda: ((a - ta) ** 2).sum() + ((b - tb) ** 2).sum())
    assert nfi.numel() == 7 and nfi.numel == 7 and nfi.flat_param.shape == (7,)
    x, f, d = fmin_l_bfgs_b(nfi.f, nfi.flat_param, nfi.fprime, m=10, factr=1e1, pgtol=1e-12)
    assert d["warnflag"] == 0
    nfi.flat_param = x
    assert torch.allclose(a.detach(), ta, atol=1e-7)
    assert b.detach()[0, 0] == 1.0 and torch.allclose(b.detach().reshape(-1)[1:], tb.reshape(-1)[1:], atol=1e-7)
    assert abs(f - 0.25) < 1e-10         # the pinned entry's misfit (1 - 0.5)^2 is all that is left
    g = nfi.fprime(x)
    assert g.dtype == np.float64 and g[3] == 0.0


def test_lagrange_interp_fields(built):
    """LagrangeInterpFixInputs2d (8f-2): partition of unity, exact reproduction of a polynomial
    of the basis degree, continuity across cells, gradient reaches the table."""
    import math
    H, W, deg, mesh = 20, 24, 2, (4, 3)
    li = built.LagrangeInterpFixInputs2d(2, (H, W), mesh, deg, ((0.0, 0.0), (2.0, 3.0)))
    assert torch.allclose(li.By.sum(1), torch.ones(H, dtype=torch.float64), atol=1e-13)
    assert torch.allclose(li.Bx.sum(1), torch.ones(W, dtype=torch.float64), atol=1e-13)
    ny, nx = mesh[0] * deg + 1, mesh[1] * deg + 1
    yn = torch.linspace(0, 2, ny, dtype=torch.float64)
    xn = torch.linspace(0, 3, nx, dtype=torch.float64)
    poly = lambda y, x: 1 + 2 * y - x + 0.5 * y * y + y * x - 0.25 * x * x * y * y   # degree 2 per variable
    with torch.no_grad():
        li.interp_coe[0] = poly(yn[:, None], xn[None, :])
        li.interp_coe[1] = torch.sin(yn)[:, None] * torch.cos(xn)[None, :]
    out = li()
    yg = 2.0 * torch.arange(H, dtype=torch.float64) / H
    xg = 3.0 * torch.arange(W, dtype=torch.float64) / W
    assert torch.allclose(out[0], poly(yg[:, None], xg[None, :]), atol=1e-12)
    assert (out[1] - torch.sin(yg)[:, None] * torch.cos(xg)[None, :]).abs().max() < 2e-2  # smooth field, O(h^3) with h = 1/2 .. 1
    out.sum().backward()
    assert li.interp_coe.grad is not None and li.interp_coe.grad.abs().sum() > 0
    m = built.VariantCoeLinear2d(5, 2, 0.1, 1e-3, (H, W), interp_degree=2, interp_mesh_size=mesh)
    assert m.coefficient_fields().shape == (5, H, W) and len(list(m.coe_params())) == 1


def test_bench_model_equals_oracle_model(built):
    """bench.py builds its synthetic Burgers model without importing the oracle (product arm); it is the
    same model, number for number, as oracle.stable_burgers_model (the parity leg compares the two sides)."""
    import math

    import bench
    from oracle import pdenet_oracle as O
    dx = 2 * math.pi / 1024
    orders, M, scales, sym = bench.model_params(7, 2, 2, 2, dx)
    o2, M2, s2, sym2 = O.stable_burgers_model(7, dx)
    assert orders == o2 and torch.equal(M, M2) and torch.equal(sym, sym2)
    assert torch.equal(torch.tensor(scales, dtype=torch.float64), s2)
    assert bench.step_dt(dx) == dx * dx


def test_torch_library_shim_registers_the_ops(built):
    """The TORCH_LIBRARY shim (csrc/torch_shim.cpp -> libpdenet_b200_torch.so) loads next to the C-ABI library
    and registers the hot-path entries as dispatcher ops (CUDA backend only: a CPU tensor has no kernel)."""
    from pdenet_b200 import _lib
    T = _lib.torch_ops()
    assert T is not None and os.path.exists(_lib.TORCH_LIB_PATH)
    for name in ("rollout_fwd", "rollout_bwd", "rhs_fwd", "rhs_bwd", "m2k", "m2k_bwd", "linear_rollout_fwd",
                 "linear_rollout_bwd"):
        assert hasattr(T, name), name
    schema = str(T.rollout_fwd.default._schema)
    assert "Tensor u0" in schema and "int steps" in schema
    with pytest.raises((NotImplementedError, RuntimeError)):
        T.m2k(torch.zeros(1, 3, 3, dtype=torch.float64), [], None)       # no CPU kernel: there is no CPU fallback
    # the shim links the C-ABI library instead of carrying kernels of its own
    out = subprocess.run(["ldd", _lib.TORCH_LIB_PATH], capture_output=True, text=True).stdout
    assert "libpdenet_b200.so" in out
