"""Random-configuration parity sweep: CUDA rollout (both reverse paths) vs the CPU oracle.

    python tools/random_parity_sweep.py [n_cases] [seed]

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
from test_parity_gpu import make_model, smooth_field

BANKS = [(1, 1), (1, 2), (1, 3), (2, 1), (2, 2), (2, 3), (3, 1), (4, 1)]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    dev = torch.device("cuda:0")
    worst, bad = {}, []
    for case in range(n):
        k = rng.choice([3, 5, 7])
        max_order, C = rng.choice([b for b in BANKS if b[0] < k])   # a k x k moment matrix holds orders < k
        L, B = rng.randint(0, 4), rng.randint(1, 3)
        H, W = rng.randint(1, 40), rng.randint(1, 70)
        steps, keep, chunk, grad_all = rng.randint(1, 3), rng.random() < 0.5, rng.randint(0, 3), rng.random() < 0.5
        seed = rng.randint(0, 10 ** 6)
        if P._lib.lib().pdenet_rollout_bwd_needs_features(8, k, len(O.order_list(max_order)), C, L) == 1:
            keep = True   # the recomputing reverse kernel does not fit this bank (refused; tested separately)
        orders, M, scales, sym = make_model(k, max_order, C, L, 0.3, seed)
        Ks = O.scaled_kernels(M, scales)
        u0 = smooth_field(B, C, H, W, seed + 1)
        w = torch.randn(steps, B, C, H, W, generator=torch.Generator().manual_seed(seed + 2), dtype=torch.float64)
        if not grad_all:
            w[:-1] = 0
        uc, Kc, sc = [t.clone().requires_grad_(True) for t in (u0, Ks, sym)]
        ref = torch.stack(O.rollout(uc, Kc, sc, L, 1e-2, steps))
        (ref * w).sum().backward()
        ud, Kd, sd = [t.to(dev).requires_grad_(True) for t in (u0, Ks, sym)]
        st = P.rollout(ud, Kd, sd, L, 1e-2, steps, batch_chunk=chunk, keep_features=keep)
        (st * w.to(dev)).sum().backward()
        errs = {}
        for name, a, b in [("states", st, ref), ("grad_u0", ud.grad, uc.grad), ("grad_kernels", Kd.grad, Kc.grad),
                           ("grad_symnet", sd.grad, sc.grad)]:
            a, b = a.detach().cpu(), b.detach()
            scale = b.abs().max().item()
            errs[name] = ((a - b).abs().max().item() / scale if scale > 0 else (a - b).abs().max().item(), scale)
        cfg = dict(k=k, max_order=max_order, C=C, L=L, B=B, H=H, W=W, steps=steps, keep=keep, chunk=chunk,
                   grad_all=grad_all, seed=seed)
        for name, (e, scale) in errs.items():
            if e > worst.get(name, (0,))[0]:
                worst[name] = (e, scale, cfg)
            if e > 1e-10 * steps:
                bad.append((name, e, scale, cfg))
    print(json.dumps({"cases": n, "worst": {k: {"rel": v[0], "ref_scale": v[1], "cfg": v[2]} for k, v in worst.items()},
                      "above_tolerance": [{"what": b[0], "rel": b[1], "ref_scale": b[2], "cfg": b[3]} for b in bad]},
                     indent=1))


if __name__ == "__main__":
    main()
