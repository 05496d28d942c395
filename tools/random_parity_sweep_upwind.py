"""Random-configuration parity sweep of scheme='upwind': fused CUDA rollout (forward + reverse) vs the CPU oracle.

    python tools/random_parity_sweep_upwind.py [n_cases] [seed]

Prints the worst relative errors and every case above 1e-10 per step (none expected)."""
import json
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch

import pdenet_b200 as P
from oracle import pdenet_oracle as O
from test_parity_gpu import smooth_field


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    dev = torch.device("cuda:0")
    worst, bad, differs = {}, [], 0
    for case in range(n):
        k = rng.choice([3, 5, 7])
        max_order, C = rng.choice([1, 2]), rng.choice([1, 2])
        if max_order >= k:
            max_order = 1
        L, B = rng.randint(0, 3), rng.randint(1, 3)
        H, W = rng.randint(1, 48), rng.randint(1, 80)
        steps, chunk, grad_all = rng.randint(1, 3), rng.randint(0, 3), rng.random() < 0.5
        seed = rng.randint(0, 10 ** 6)
        orders = O.order_list(max_order)
        N = len(orders)
        if P._lib.lib().pdenet_upwind_supported(8, k, N, C, L) != 1:
            continue
        g = torch.Generator().manual_seed(seed)
        dx = 0.3
        M = O.initial_moments(k, orders) + 5e-2 * torch.randn(N, k, k, generator=g, dtype=torch.float64)
        Ks = O.scaled_kernels(M, O.filter_scales(orders, dx))
        sym = 0.3 * torch.randn(C, O.symnet_param_count(C * N, L), generator=g, dtype=torch.float64)
        u0 = smooth_field(B, C, H, W, seed + 1)
        w = torch.randn(steps, B, C, H, W, generator=g, dtype=torch.float64)
        if not grad_all:
            w[:-1] = 0
        dt = 1e-2
        uc, Kc, sc = [t.clone().requires_grad_(True) for t in (u0, Ks, sym)]
        ref = torch.stack(O.upwind_rollout(uc, Kc, sc, L, dt, steps))
        (ref * w).sum().backward()
        central = torch.stack(O.rollout(u0, Ks, sym, L, dt, steps))
        if ((central - ref.detach()).abs().max() / ref.detach().abs().max()).item() > 1e-7:
            differs += 1
        ud, Kd, sd = [t.to(dev).requires_grad_(True) for t in (u0, Ks, sym)]
        st = P.rollout(ud, Kd, sd, L, dt, steps, batch_chunk=chunk, scheme="upwind")
        (st * w.to(dev)).sum().backward()
        errs = {}
        for name, a, b in [("states", st, ref), ("grad_u0", ud.grad, uc.grad), ("grad_kernels", Kd.grad, Kc.grad),
                           ("grad_symnet", sd.grad, sc.grad)]:
            a, b = a.detach().cpu(), b.detach()
            scale = b.abs().max().item()
            errs[name] = ((a - b).abs().max().item() / scale if scale > 0 else (a - b).abs().max().item(), scale)
        cfg = dict(k=k, max_order=max_order, C=C, L=L, B=B, H=H, W=W, steps=steps, chunk=chunk, grad_all=grad_all,
                   seed=seed)
        for name, (e, scale) in errs.items():
            if e > worst.get(name, (0,))[0]:
                worst[name] = (e, scale, cfg)
            if e > 1e-10 * steps:
                bad.append((name, e, scale, cfg))
    print(json.dumps({"cases": n, "cases_where_upwind_differs_from_central": differs,
                      "worst": {k: {"rel": v[0], "ref_scale": v[1], "cfg": v[2]} for k, v in worst.items()},
                      "above_tolerance": [{"what": b[0], "rel": b[1], "ref_scale": b[2], "cfg": b[3]} for b in bad]},
                     indent=1))


if __name__ == "__main__":
    main()
