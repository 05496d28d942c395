"""fp32 accuracy of the two linear-rollout paths (cluster / tiled) against the fp64 result, config-2 shape."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import pdenet_b200 as P
from pdenet_b200 import _lib
from oracle import pdenet_oracle as O
dev = torch.device("cuda:0")
k, max_order, B, H, W, steps = 5, 4, 16, 50, 50, 20
g = torch.Generator().manual_seed(41)
orders = [o for o in O.order_list(max_order) if o != (0, 0)]
N = len(orders); dx = 2 * math.pi / W; dt = 1.5e-2 * dx * dx
M = O.initial_moments(k, orders) + 1e-3 * torch.randn(N, k, k, generator=g, dtype=torch.float64)
Ks = O.scaled_kernels(M, O.filter_scales(orders, dx))
K0 = O.m2k(O.initial_moments(k, [(0, 0)]) + 1e-3 * torch.randn(1, k, k, generator=g, dtype=torch.float64))[0]
amp = torch.tensor([0.5 * dx ** max(0, o0 + o1 - 2) for (o0, o1) in orders], dtype=torch.float64)
coe = amp[:, None, None] * torch.randn(N, H, W, generator=g, dtype=torch.float64)
y = torch.arange(H, dtype=torch.float64) * dx
u0 = (torch.sin(y)[:, None] * torch.cos(2 * y)[None, :])[None].repeat(B, 1, 1)
def run(dtype, on, st):
    prev = _lib.lib().pdenet_linear_small_grid_path(on)
    lv = [t.to(dev, dtype).clone().requires_grad_(True) for t in (u0, K0, Ks, coe)]
    out = P.linear_rollout(*lv, dt, st, last_only=True)
    out.square().mean().backward()
    _lib.lib().pdenet_linear_small_grid_path(prev)
    return [out.detach().double()] + [t.grad.double() for t in lv]
def rel(a, b): return ((a - b).abs().max() / b.abs().max()).item()
for st in (1, 5, 20):
    ref = run(torch.float64, 0, st)
    for on in (1, 0):
        r = run(torch.float32, on, st)
        print("steps", st, "cluster" if on else "tiled  ", ["%.1e" % rel(a, b) for a, b in zip(r, ref)], "max|u| %.2f" % ref[0].abs().max().item())
    c64 = run(torch.float64, 1, st)
    print("steps", st, "cluster fp64 vs tiled fp64", ["%.1e" % rel(a, b) for a, b in zip(c64, ref)])
