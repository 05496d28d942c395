"""Per-kernel times of BASELINE configs 1 and 3 (eager fwd+bwd, library launches bracketed by CUDA events:
ops.profile_enable) next to the flops each kernel does -- where the small configurations lose against cfg4."""
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import pdenet_b200 as P
from bench import CFG, model_params, step_dt, synth_fields
from pdenet_b200 import ops

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
peak = ops.fma_peak(torch.float64, 2048)
FLOPS = {"fwd_step": 1440, "bwd_kernel_A": 788, "bwd_kernel_B1": 1176, "bwd_kernel_B2": 1176}   # kept features
for name, B, H, W, steps in (("cfg1", 8, 100, 100, 1), ("cfg3", 64, 128, 128, 15)):
    dx = 2 * math.pi / W
    dt = step_dt(dx)
    orders, M, scales, sym = model_params(CFG["k"], 2, 2, 2, dx)
    M = M.to(dev).requires_grad_(True)
    sym = sym.to(dev).requires_grad_(True)
    u0 = synth_fields(0, B, 2, H, W, dev, torch.float64)

    def fn():
        o = P.rollout(u0, P.m2k(M, scales), sym, 2, dt, steps, last_only=True)
        return torch.autograd.grad(o.square().mean(), (M, sym))

    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ops.profile_enable(True)
    ops.profile_read()
    reps = 10
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    ms, n = ops.profile_read()
    ops.profile_enable(False)
    rec = {"config": name, "B": B, "H": H, "W": W, "steps": steps, "fma_peak_tflops": peak / 1e12}
    pts = B * H * W
    for kind in ms:
        if n[kind]:
            per = ms[kind] / n[kind]
            rec[kind] = {"launches_per_pass": n[kind] / reps, "us_per_launch": per * 1e3}
            if kind in FLOPS:
                rec[kind]["frac_of_fma_peak"] = FLOPS[kind] * pts / (per * 1e-3) / peak
    rec["sum_ms_per_pass"] = sum(ms.values()) / reps
    print(json.dumps(rec))
