"""Time the small BASELINE configs (1-3) fwd+bwd on one GPU: python tools/config_times.py"""
import json
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pdenet_b200 as P
from bench import model_params, synth_fields

dev = torch.device("cuda:0")


KERNEL_MS = {}


def time_graph(fn_graphable, reps=50):
    """Capture one fwd+bwd into a CUDA graph and time the replays (launch overhead removed)."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn_graphable()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn_graphable()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def time_it(fn, reps=20):
    from pdenet_b200 import ops
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    # second pass with per-launch events: how much of the wall time is inside our kernels
    ops.profile_read()
    ops.profile_enable(True)
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    kms, kn = ops.profile_read()
    ops.profile_enable(False)
    KERNEL_MS["last"] = ({k: round(v / reps, 4) for k, v in kms.items()}, {k: v // reps for k, v in kn.items()})
    return ms


def symnet_cfg(name, B, H, W, steps, dtype=torch.float64, k=7, C=2, L=2):
    dx = 2 * math.pi / W
    dt = 0.25 * dx * dx
    orders, M, scales, sym = model_params(k, 2, C, L, dx)
    M = M.to(dev, dtype).requires_grad_(True)
    sym = sym.to(dev, dtype).requires_grad_(True)
    u0 = synth_fields(0, B, C, H, W, dev, dtype)

    def fn():
        M.grad = sym.grad = None
        out = P.rollout(u0, P.m2k(M, scales), sym, L, dt, steps, last_only=True)
        out.square().mean().backward()

    ms = time_it(fn)

    def fng():
        out = P.rollout(u0, P.m2k(M, scales), sym, L, dt, steps, last_only=True)
        return torch.autograd.grad(out.square().mean(), (M, sym))

    msg = time_graph(fng)
    return {"config": name, "ms_fwd_bwd": ms, "ms_fwd_bwd_cuda_graph": msg,
            "grid_point_steps_per_s": B * H * W * steps / (ms * 1e-3),
            "grid_point_steps_per_s_cuda_graph": B * H * W * steps / (msg * 1e-3),
            "kernel_ms": KERNEL_MS["last"][0], "launches": KERNEL_MS["last"][1]}


def linear_cfg(name, B, H, W, steps, k=5, max_order=4):
    from pdenet_b200.modules import order_list
    dx = 2 * math.pi / W
    dt = 1.5e-2 * dx * dx
    g = torch.Generator().manual_seed(0)
    orders = [o for o in order_list(max_order) if o != (0, 0)]
    N = len(orders)
    M = torch.zeros(N + 1, k, k, dtype=torch.float64)
    M[0, 0, 0] = 1
    for n, (a, b) in enumerate(orders):
        M[n + 1, a, b] = 1
    M = (M + 1e-3 * torch.randn(M.shape, generator=g, dtype=torch.float64)).to(dev).requires_grad_(True)
    scales = [1.0] + [1.0 / dx ** (a + b) for a, b in orders]
    coe = torch.randn(N, H, W, generator=g, dtype=torch.float64).to(dev).requires_grad_(True)
    u0 = synth_fields(0, B, 1, H, W, dev)[:, 0].contiguous()

    def fn():
        M.grad = coe.grad = None
        ks = P.m2k(M, scales)
        out = P.linear_rollout(u0, ks[0], ks[1:], coe, dt, steps, last_only=True)
        out.square().mean().backward()

    ms = time_it(fn)

    def fng():
        ks = P.m2k(M, scales)
        out = P.linear_rollout(u0, ks[0], ks[1:], coe, dt, steps, last_only=True)
        return torch.autograd.grad(out.square().mean(), (M, coe))

    msg = time_graph(fng)
    return {"config": name, "ms_fwd_bwd": ms, "ms_fwd_bwd_cuda_graph": msg,
            "grid_point_steps_per_s": B * H * W * steps / (ms * 1e-3),
            "grid_point_steps_per_s_cuda_graph": B * H * W * steps / (msg * 1e-3),
            "kernel_ms": KERNEL_MS["last"][0], "launches": KERNEL_MS["last"][1]}


if __name__ == "__main__":
    res = [symnet_cfg("cfg1: Burgers 100x100 B8 1 block fp64", 8, 100, 100, 1),
           linear_cfg("cfg2: PDE-Net1 linear 50x50 B16 20 blocks fp64", 16, 50, 50, 20),
           symnet_cfg("cfg3: reaction-diffusion shape 128x128 B64 15 blocks fp64", 64, 128, 128, 15),
           symnet_cfg("cfg3 in fp32", 64, 128, 128, 15, torch.float32),
           symnet_cfg("cfg4 slice fp32: 1024x1024 B16 2 blocks", 16, 1024, 1024, 2, torch.float32)]
    for r in res:
        print(json.dumps(r))
