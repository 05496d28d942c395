"""Random-configuration parity sweep of the PDE-Net 1 linear rollout (row a12) and of the fp32
PDE-Net 2 rollout against the CPU oracle:  python tools/random_parity_sweep_linear.py [n] [seed]"""
import json
import math
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch

import pdenet_b200 as P
from oracle import pdenet_oracle as O
from test_parity_gpu import make_model, smooth_field


def relerr(a, b):
    a, b = a.detach().double().cpu(), b.detach().double()
    s = b.abs().max().item()
    return (a - b).abs().max().item() / (s if s > 0 else 1.0)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 150
    rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    dev = torch.device("cuda:0")
    worst, bad = {}, []
    for case in range(n):
        k = rng.choice([3, 5, 7])
        max_order = rng.choice([m for m in (1, 2, 3, 4) if m < k])
        B, H, W, steps = rng.randint(1, 4), rng.randint(1, 60), rng.randint(1, 70), rng.randint(1, 4)
        seed = rng.randint(0, 10 ** 6)
        g = torch.Generator().manual_seed(seed)
        orders = [o for o in O.order_list(max_order) if o != (0, 0)]
        N = len(orders)
        dx = 2 * math.pi / max(H, W, 2)
        dt = 1.5e-2 * dx ** 2
        M = O.initial_moments(k, orders) + 1e-2 * torch.randn(N, k, k, generator=g, dtype=torch.float64)
        Ks = O.scaled_kernels(M, O.filter_scales(orders, dx))
        K0 = O.m2k(O.initial_moments(k, [(0, 0)]) + 1e-3 * torch.randn(1, k, k, generator=g, dtype=torch.float64))[0]
        coe = torch.randn(N, H, W, generator=g, dtype=torch.float64)
        u0 = smooth_field(B, 1, H, W, seed + 1)[:, 0]
        w = torch.randn(steps, B, H, W, generator=g, dtype=torch.float64)
        leaves = [t.clone().requires_grad_(True) for t in (u0, K0, Ks, coe)]
        ref = torch.stack(O.linear_rollout(*leaves, dt, steps))
        (ref * w).sum().backward()
        dl = [t.to(dev).requires_grad_(True) for t in (u0, K0, Ks, coe)]
        out = P.linear_rollout(*dl, dt, steps)
        (out * w.to(dev)).sum().backward()
        cfg = dict(model="linear", k=k, max_order=max_order, B=B, H=H, W=W, steps=steps, seed=seed)
        errs = {"lin_states": relerr(out, ref)}
        for name, a, b in zip(["lin_grad_u0", "lin_grad_K0", "lin_grad_Ks", "lin_grad_coe"], dl, leaves):
            errs[name] = relerr(a.grad, b.grad)
        for name, e in errs.items():
            if e > worst.get(name, (0,))[0]:
                worst[name] = (e, cfg)
            if e > 1e-10 * steps:
                bad.append((name, e, cfg))
    print(json.dumps({"cases": n, "worst": {k: {"rel": v[0], "cfg": v[1]} for k, v in worst.items()},
                      "above_tolerance": [{"what": b[0], "rel": b[1], "cfg": b[2]} for b in bad]}, indent=1))


if __name__ == "__main__":
    main()
