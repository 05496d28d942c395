"""Kernel nodes of the captured fwd+bwd chains of the small BASELINE configs, for an ncu launch list:

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file out.csv python tools/probes/small_cfg_graph_launches.py

Replays, in this order and once each inside cudaProfilerStart/Stop: (1) cfg1 through the library's graphed training
closure, (2) cfg1 as a user-side graph of the public ops + a torch loss, (3) cfg2 as a user-side graph.  A marker
kernel (torch.zeros(7)) separates them.  (Memset / memcpy nodes are not kernels and do not show.)"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import pdenet_b200 as P
from bench import CFG, build_model, model_params, step_dt, synth_fields
from pdenet_b200 import losses
from pdenet_b200.modules import order_list

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)


def capture(fn):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g


B, H, W, steps = 8, 100, 100, 1
model, _ = build_model(P, dev, torch.float64, H, W, 0, True)
u0 = synth_fields(0, B, 2, H, W, dev, torch.float64)
closure = losses.GraphedTrainingLoss(model, u0, torch.stack([u0] * steps))

dx = 2 * math.pi / W
dt = step_dt(dx)
orders, M, scales, sym = model_params(CFG["k"], 2, 2, 2, dx)
M = M.to(dev).requires_grad_(True)
sym = sym.to(dev).requires_grad_(True)


def cfg1():
    o = P.rollout(u0, P.m2k(M, scales), sym, 2, dt, steps, last_only=True)
    return torch.autograd.grad(o.square().mean(), (M, sym))


g1 = capture(cfg1)

B2, H2, W2, steps2, k2, mo2 = 16, 50, 50, 20, 5, 4
dx2 = 2 * math.pi / W2
dt2 = 1.5e-2 * dx2 * dx2
gen = torch.Generator().manual_seed(0)
orders2 = [o for o in order_list(mo2) if o != (0, 0)]
N2 = len(orders2)
M2 = torch.zeros(N2 + 1, k2, k2, dtype=torch.float64)
M2[0, 0, 0] = 1
for n, (a, b) in enumerate(orders2):
    M2[n + 1, a, b] = 1
M2 = (M2 + 1e-3 * torch.randn(M2.shape, generator=gen, dtype=torch.float64)).to(dev).requires_grad_(True)
scales2 = [1.0] + [1.0 / dx2 ** (a + b) for a, b in orders2]
coe = torch.randn(N2, H2, W2, generator=gen, dtype=torch.float64).to(dev).requires_grad_(True)
v0 = synth_fields(0, B2, 1, H2, W2, dev)[:, 0].contiguous()


def cfg2():
    ks = P.m2k(M2, scales2)
    o = P.linear_rollout(v0, ks[0], ks[1:], coe, dt2, steps2, last_only=True)
    return torch.autograd.grad(o.square().mean(), (M2, coe))


g2 = capture(cfg2)
torch.cuda.synchronize()
torch.cuda.profiler.start()
closure()
torch.zeros(7, device=dev)
g1.replay()
torch.zeros(7, device=dev)
g2.replay()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
