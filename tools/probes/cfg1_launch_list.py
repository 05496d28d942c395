"""Five eager cfg1 fwd+bwd passes (for an ncu launch list: which kernels make up the 13-launch chain)."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import pdenet_b200 as P
from bench import CFG, model_params, step_dt, synth_fields

dev = torch.device("cuda", 0)
B, H, W, steps = 8, 100, 100, 1
dx = 2 * math.pi / W
dt = step_dt(dx)
orders, M, scales, sym = model_params(CFG["k"], 2, 2, 2, dx)
M = M.to(dev).requires_grad_(True)
sym = sym.to(dev).requires_grad_(True)
u0 = synth_fields(0, B, 2, H, W, dev, torch.float64)
for i in range(5):
    torch.cuda.nvtx.range_push("pass%d" % i)
    o = P.rollout(u0, P.m2k(M, scales), sym, 2, dt, steps, last_only=True)
    torch.autograd.grad(o.square().mean(), (M, sym))
    torch.cuda.nvtx.range_pop()
torch.cuda.synchronize()
