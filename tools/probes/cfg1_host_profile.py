"""Where the host time of an eager cfg1 fwd+bwd goes (cProfile over 300 iterations)."""
import cProfile
import math
import os
import pstats
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


def fn():
    o = P.rollout(u0, P.m2k(M, scales), sym, 2, dt, steps, last_only=True)
    return torch.autograd.grad(o.square().mean(), (M, sym))


for _ in range(20):
    fn()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(300):
    fn()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
