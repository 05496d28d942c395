"""Print fp32 CUDA-vs-fp64-oracle errors next to fp32-oracle-vs-fp64-oracle errors."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from test_parity_gpu import CASES, run_both, rel, make_model, smooth_field
from oracle import pdenet_oracle as O
import pdenet_b200 as P

dev = torch.device("cuda:0")
for case in CASES[:5]:
    got, ref = run_both(P, dev, case, torch.float32, True)
    k, max_order, C, L, B, H, W, steps, dx, dt = case
    orders, M, scales, sym = make_model(k, max_order, C, L, dx, 7)
    Ks = O.scaled_kernels(M, scales)
    u0 = smooth_field(B, C, H, W, 8)
    g = torch.Generator().manual_seed(9)
    w = torch.randn(steps, B, C, H, W, generator=g, dtype=torch.float64)
    lv = [t.float().clone().requires_grad_(True) for t in (u0, Ks, sym)]
    st = torch.stack(O.rollout(lv[0], lv[1], lv[2], L, dt, steps))
    (st * w.float()).sum().backward()
    cpu32 = (st, lv[0].grad, lv[1].grad, lv[2].grad)
    print(case, "cuda32:", ["%.2e" % rel(a, b) for a, b in zip(got, ref)], "cpu32:",
          ["%.2e" % rel(a, b) for a, b in zip(cpu32, ref)])
