"""python tools/fp32_error_report.py > profiles/r2_fp32_parity.json   (on a GPU box)

Where single precision actually sits ("nobody knows"): for the BASELINE grid shapes,
  * the FD2d features per derivative order: CUDA fp32 and the oracle run in fp32 on the CPU, both against the
    fp64 oracle, next to the a-priori bound 2 k u32 ||K_n||_1 ||u||_inf the GPU test uses;
  * a fwd+bwd rollout (training-style loss): state and every gradient, CUDA fp32 and CPU-fp32 oracle against
    the fp64 oracle, next to the tolerance policy of tests/test_parity_gpu.py (fp32_grad_scale)."""
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pdenet_b200 as P
from bench import model_params, step_dt, synth_fields
from oracle import pdenet_oracle as O

dev = torch.device("cuda:0")


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max()).item()


def features(H, W, k=7):
    dx = 2 * math.pi / W
    orders, M, scales, _ = model_params(k, 2, 2, 2, dx)
    Ks = O.scaled_kernels(M, torch.tensor(scales, dtype=torch.float64))
    u = synth_fields(0, 2, 2, H, W, "cpu")
    ref = O.fd2d_apply(u, Ks)
    cpu32 = O.fd2d_apply(u.float(), Ks.float())
    gpu32 = P.fd2d_features(u.to(dev, torch.float32), Ks.to(dev, torch.float32)).cpu()
    N = len(orders)
    out = []
    for n, o in enumerate(orders):
        r = ref.reshape(2, 2, N, H, W)[:, :, n]
        bound = 2 * k * 2.0 ** -24 * Ks[n].abs().sum() * u.abs().max()
        out.append({"order": list(o), "l1_norm_of_taps": Ks[n].abs().sum().item(), "max_abs_feature": r.abs().max().item(),
                    "cuda_fp32_abs_err": (gpu32.double().reshape(2, 2, N, H, W)[:, :, n] - r).abs().max().item(),
                    "cpu_fp32_oracle_abs_err": (cpu32.double().reshape(2, 2, N, H, W)[:, :, n] - r).abs().max().item(),
                    "bound_abs": bound.item()})
        out[-1]["cuda_fp32_rel_err"] = out[-1]["cuda_fp32_abs_err"] / out[-1]["max_abs_feature"]
    return out


def rollout(B, H, W, steps, k=7):
    dx = 2 * math.pi / W
    dt = step_dt(dx)
    orders, M, scales, sym = model_params(k, 2, 2, 2, dx)
    Ks = O.scaled_kernels(M, torch.tensor(scales, dtype=torch.float64))
    u0 = synth_fields(0, B, 2, H, W, "cpu")
    res = {}
    for name, dtype, where in (("ref", torch.float64, "cpu"), ("cpu_fp32_oracle", torch.float32, "cpu"),
                               ("cuda_fp32", torch.float32, dev), ("cuda_fp64", torch.float64, dev)):
        lv = [t.to(where, dtype).clone().requires_grad_(True) for t in (u0, Ks, sym)]
        if where == "cpu":
            out = O.rollout(*lv, 2, dt, steps)[-1]
        else:
            out = P.rollout(*lv, 2, dt, steps, last_only=True)
        out.square().mean().backward()
        res[name] = (out.detach(), lv[0].grad, lv[1].grad, lv[2].grad)
    names = ["state", "grad_u0", "grad_kernels", "grad_symnet"]
    gscale = max(1.0, float(Ks.abs().sum(dim=(1, 2)).max()) / 24.0)
    return {"shape": {"B": B, "H": H, "W": W, "dt_blocks": steps, "dt": dt},
            "tolerance": {"state": 1e-5 * steps, "gradients": 1e-5 * steps * gscale, "fp32_grad_scale": gscale},
            **{who: dict(zip(names, [rel(a, b) for a, b in zip(res[who], res["ref"])]))
               for who in ("cuda_fp32", "cpu_fp32_oracle", "cuda_fp64")}}


if __name__ == "__main__":
    rep = {"what": "fp32 error of the CUDA path and of the oracle itself run in fp32, both against the fp64 oracle "
                   "(synthetic Burgers model of bench.py, dt = dx^2, loss = mean square of the final state)",
           "features_by_order": {"100x100": features(100, 100), "128x128": features(128, 128),
                                 "1024x1024": features(1024, 1024)},
           "rollout": {"cfg1": rollout(8, 100, 100, 1), "cfg3_shape": rollout(8, 128, 128, 15),
                       "cfg4_shape": rollout(2, 1024, 1024, 4)}}
    print(json.dumps(rep, indent=1))
