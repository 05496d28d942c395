"""torchrun --nproc-per-node N tools/slab_multigpu_check.py [--size 2048] [--steps 4] [--dtype f64]

Slab domain decomposition over N GPUs (NCCL halo exchange): checks the decomposed forward + reverse
pass against the single-GPU periodic rollout on rank 0 and prints the slab throughput (config 5)."""
import argparse
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import pdenet_b200 as P
from pdenet_b200 import parallel as PP
from bench import model_params, step_dt, synth_fields


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--p2p", action="store_true", help="fused peer-memory halo writes instead of NCCL send/recv")
    ap.add_argument("--keep", choices=["on", "off"], default="on",
                    help="keep the derivative features of the owned rows for the reverse steps (off = recompute)")
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    dtype = torch.float64 if a.dtype == "f64" else torch.float32
    H = W = a.size
    k, C, L = 7, 2, 2
    dx = 2 * math.pi / W
    dt = step_dt(dx)
    orders, M, scales, sym = model_params(k, 2, C, L, dx)
    Ks = P.m2k(M.to(dev, dtype), scales)
    sym = sym.to(dev, dtype)
    slab = PP.SlabDecomposition(H, 3, rank, world)
    u_full = synth_fields(0, 1, C, H, W, dev, dtype) if (rank == 0 or not a.no_check) else None
    if a.no_check:
        mine = synth_fields(0, 1, C, H, W, dev, dtype)[:, :, slab.row0:slab.row0 + slab.Hs].contiguous()
    else:
        mine = u_full[:, :, slab.row0:slab.row0 + slab.Hs].contiguous()
    n_total = C * H * W

    def seed(f):  # loss = mean square of the final state over the WHOLE grid
        return 2.0 * f / n_total

    if a.p2p:
        pb = PP.SlabPeerBuffers(slab, 1, C, W, a.steps, dtype, dev)

        def run():
            return PP.slab_rollout_fwd_bwd_p2p(slab, pb, mine, Ks, sym, L, dt, a.steps, seed, keep_features=a.keep == "on")
    else:
        def run():
            return slab.rollout_fwd_bwd(mine, Ks, sym, L, dt, a.steps, seed, keep_features=a.keep == "on")

    fin, gu, gk, gs = run()
    torch.cuda.synchronize()
    if a.p2p:
        # the in-place variant (input already in the symmetric buffer, outputs as views, seed written in place) must
        # give the same bits
        def seed_out(f, out=None):
            return torch.mul(f, 2.0 / n_total, out=out)
        pb.input_view().copy_(mine)
        f2, g2, k2, s2 = PP.slab_rollout_fwd_bwd_p2p(slab, pb, pb.input_view(), Ks, sym, L, dt, a.steps, seed_out,
                                                     keep_features=a.keep == "on", copy_outputs=False)
        same = all(torch.equal(x, y) for x, y in ((fin, f2), (gu, g2), (gk, k2), (gs, s2)))
        assert same, "in-place slab pass differs from the copying one"
        del f2, g2
    errs = None
    if not a.no_check:
        ud, Kd, sd = [t.clone().requires_grad_(True) for t in (u_full, Ks, sym)]
        ref = P.rollout(ud, Kd, sd, L, dt, a.steps, last_only=True)
        ref.square().mean().backward()
        sl = slice(slab.row0, slab.row0 + slab.Hs)

        def rel(x, y):
            return ((x - y).abs().max() / y.abs().max()).item()

        errs = torch.tensor([rel(fin, ref.detach()[:, :, sl]), rel(gu, ud.grad[:, :, sl]), rel(gk, Kd.grad),
                             rel(gs, sd.grad)], device=dev)
        dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(a.reps):
        run()
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.reps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        tol = 1e-10 * a.steps if dtype == torch.float64 else 1e-3
        out = {"what": "slab-decomposed fwd+bwd rollout, 1 trajectory" + (", fused peer halo writes" if a.p2p else ", NCCL halo exchange"), "n_gpus": world, "grid": [H, W], "dtype": a.dtype,
               "features": "kept" if a.keep == "on" else "recomputed",
               "steps": a.steps, "ms_per_pass": ms.item(), "grid_point_steps_per_s": H * W * a.steps / (ms.item() * 1e-3),
               "max_rel_err_vs_single_gpu": None if errs is None else
               dict(zip(["state", "grad_u0", "grad_kernels", "grad_symnet"], errs.tolist()))}
        print(json.dumps(out))
        if errs is not None:
            assert errs.max().item() < tol, errs
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
