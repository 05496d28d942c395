import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import pdenet_b200 as P
from oracle import pdenet_oracle as O
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
dev = torch.device("cuda:0")
def smooth_field(B, C, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    y = torch.arange(H, dtype=torch.float64) * (2 * math.pi / H); x = torch.arange(W, dtype=torch.float64) * (2 * math.pi / W)
    Y, X = torch.meshgrid(y, x, indexing="ij"); u = torch.zeros(B, C, H, W, dtype=torch.float64)
    for p in range(-2, 3):
        for q in range(-2, 3):
            a = torch.randn(B, C, 1, 1, generator=g, dtype=torch.float64); b = torch.randn(B, C, 1, 1, generator=g, dtype=torch.float64)
            u += a * torch.cos(p * X + q * Y) + b * torch.sin(p * X + q * Y)
    return u / u.abs().amax(dim=(2, 3), keepdim=True)
def rel(a, b): return ((a.double().cpu() - b.double().cpu()).abs().max() / b.double().abs().max()).item()
k, max_order, B, H, W = 5, 2, 4, 50, 50
dx, dt = 2 * math.pi / W, 2e-3
for dtype in (torch.float64, torch.float32):
    model = P.SingleNonLinear2d(k, max_order, dx, dt, nonlinear_interp_degree=2, nonlinear_interp_mesh_bound=(-1.5, 1.5), nonlinear_interp_mesh_size=6, dtype=dtype).to(dev)
    g = torch.Generator().manual_seed(79)
    with torch.no_grad():
        for fd in [model.id] + model.fds:
            fd.moment.add_(1e-2 * torch.randn(fd.moment.shape, generator=g, dtype=torch.float64).to(dev, dtype)); fd.x_proj()
        model.coe.copy_((0.3 * dx ** 2 * torch.randn(model.coe.shape, generator=g, dtype=torch.float64)).to(dtype))
        model.nonlinearfitter.interp_coe.copy_(torch.sin(model.nonlinearfitter.nodes() * 2))
    u0 = smooth_field(B, 1, H, W, 80)[:, 0]
    ks = model.scaled_kernels().detach().double().cpu()
    for steps in (1, 6):
        with torch.no_grad():
            a = model(u0.to(dev, dtype), steps, fused=True); b = model(u0.to(dev, dtype), steps, fused=False)
            u = u0
            for _ in range(steps):
                u = O.single_nonlinear_step(u, ks[0], ks[1:], model.coe.detach().double().cpu(), model.nonlinearfitter.interp_coe.detach().double().cpu(), -1.5, 1.5, 6, 2, dt)
        print(dtype, steps, "fused-oracle %.2e loop-oracle %.2e fused-loop %.2e max|u| %.3f" % (rel(a, u), rel(b, u), rel(a, b), u.abs().max().item()))
