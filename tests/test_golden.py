"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py from the
oracle): the oracle must keep reproducing them (CPU), the CUDA path must match them (GPU)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import pdenet_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def T(x):
    return torch.from_numpy(np.array(x))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "rollout_*.npz"))))
def test_oracle_reproduces_golden_rollout(path):
    z = np.load(path)
    M, sym, u0 = [T(z[n]).requires_grad_(True) for n in ("moments", "symnet", "u0")]
    st = torch.stack(O.rollout(u0, O.scaled_kernels(M, T(z["scales"])), sym, int(z["L"]), float(z["dt"]),
                               int(z["steps"])))
    (st * T(z["w"])).sum().backward()
    assert rel(st.detach(), z["states"]) < 1e-12
    assert rel(u0.grad, z["grad_u0"]) < 1e-11
    assert rel(M.grad, z["grad_moments"]) < 1e-11
    assert rel(sym.grad, z["grad_symnet"]) < 1e-11


def test_oracle_reproduces_golden_upwind():
    """scheme='upwind' (row f-1): frozen outputs of the oracle restatement; also reproduced by the independent
    numpy restatement (closed-form velocity, hand-derived adjoint)."""
    from oracle import pdenet_oracle_np as ONP
    z = np.load(os.path.join(GOLD, "upwind_k5_o2_c2_l2.npz"))
    M, sym, u0 = [T(z[n]).requires_grad_(True) for n in ("moments", "symnet", "u0")]
    Ks = O.scaled_kernels(M, T(z["scales"]))
    st = torch.stack(O.upwind_rollout(u0, Ks, sym, int(z["L"]), float(z["dt"]), int(z["steps"])))
    (st * T(z["w"])).sum().backward()
    assert rel(st.detach(), z["states"]) < 1e-12
    assert rel(u0.grad, z["grad_u0"]) < 1e-11
    assert rel(M.grad, z["grad_moments"]) < 1e-11
    assert rel(sym.grad, z["grad_symnet"]) < 1e-11
    w_last_only = np.zeros_like(z["w"])
    w_last_only[-1] = z["w"][-1]
    stn, _, _, _ = ONP.upwind_rollout_fwd_bwd(z["u0"], Ks.detach().numpy(), z["symnet"], int(z["L"]), float(z["dt"]),
                                              int(z["steps"]), z["w"][-1])
    assert rel(stn[-1], z["states"][-1]) < 1e-12


def test_oracle_reproduces_golden_linear():
    z = np.load(os.path.join(GOLD, "linear_k5_o4.npz"))
    lv = [T(z[n]).requires_grad_(True) for n in ("u0", "K0", "Ks", "coe")]
    st = torch.stack(O.linear_rollout(*lv, float(z["dt"]), int(z["steps"])))
    (st * T(z["w"])).sum().backward()
    assert rel(st.detach(), z["states"]) < 1e-12
    for t, n in zip(lv, ("grad_u0", "grad_K0", "grad_Ks", "grad_coe")):
        assert rel(t.grad, z[n]) < 1e-11, n


@pytest.mark.gpu
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "rollout_*.npz"))))
def test_cuda_matches_golden_rollout(cuda_device, path):
    import pdenet_b200 as P
    z = np.load(path)
    M, sym, u0 = [T(z[n]).to(cuda_device).requires_grad_(True) for n in ("moments", "symnet", "u0")]
    st = P.rollout(u0, P.m2k(M, z["scales"].tolist()), sym, int(z["L"]), float(z["dt"]), int(z["steps"]))
    (st * T(z["w"]).to(cuda_device)).sum().backward()
    steps = int(z["steps"])
    assert rel(st.detach().cpu(), z["states"]) < 1e-10 * steps
    assert rel(u0.grad.cpu(), z["grad_u0"]) < 1e-10 * steps
    assert rel(M.grad.cpu(), z["grad_moments"]) < 1e-10 * steps
    assert rel(sym.grad.cpu(), z["grad_symnet"]) < 1e-10 * steps


@pytest.mark.gpu
def test_cuda_matches_golden_linear(cuda_device):
    import pdenet_b200 as P
    z = np.load(os.path.join(GOLD, "linear_k5_o4.npz"))
    lv = [T(z[n]).to(cuda_device).requires_grad_(True) for n in ("u0", "K0", "Ks", "coe")]
    st = P.linear_rollout(*lv, float(z["dt"]), int(z["steps"]))
    (st * T(z["w"]).to(cuda_device)).sum().backward()
    steps = int(z["steps"])
    assert rel(st.detach().cpu(), z["states"]) < 1e-10 * steps
    for t, n in zip(lv, ("grad_u0", "grad_K0", "grad_Ks", "grad_coe")):
        assert rel(t.grad.cpu(), z[n]) < 1e-10 * steps, n


@pytest.mark.gpu
def test_cuda_matches_golden_upwind(cuda_device):
    import pdenet_b200 as P
    z = np.load(os.path.join(GOLD, "upwind_k5_o2_c2_l2.npz"))
    M, sym, u0 = [T(z[n]).to(cuda_device).requires_grad_(True) for n in ("moments", "symnet", "u0")]
    st = P.rollout(u0, P.m2k(M, z["scales"].tolist()), sym, int(z["L"]), float(z["dt"]), int(z["steps"]),
                   scheme="upwind")
    (st * T(z["w"]).to(cuda_device)).sum().backward()
    steps = int(z["steps"])
    assert rel(st.detach().cpu(), z["states"]) < 1e-10 * steps
    assert rel(u0.grad.cpu(), z["grad_u0"]) < 1e-10 * steps
    assert rel(M.grad.cpu(), z["grad_moments"]) < 1e-10 * steps
    assert rel(sym.grad.cpu(), z["grad_symnet"]) < 1e-10 * steps
