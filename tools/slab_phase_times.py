"""torchrun --nproc-per-node N tools/slab_phase_times.py [--size 8192] [--steps 4]
Where a slab-decomposed pass spends its time: per-phase CUDA-event times on rank 0 (fused peer-store path)."""
import argparse
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import pdenet_b200 as P
from bench import model_params, step_dt, synth_fields
from pdenet_b200 import ops
from pdenet_b200 import parallel as PP


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=8192)
    ap.add_argument("--steps", type=int, default=4)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    dtype = torch.float64
    H = W = a.size
    k, C, L = 7, 2, 2
    dx = 2 * math.pi / W
    dt = step_dt(dx)
    orders, M, scales, sym = model_params(k, 2, C, L, dx)
    Ks = P.m2k(M.to(dev, dtype), scales)
    sym = sym.to(dev, dtype)
    slab = PP.SlabDecomposition(H, 3, rank, world)
    mine = synth_fields(0, 1, C, H, W, dev, dtype)[:, :, slab.row0:slab.row0 + slab.Hs].contiguous()
    n_total = C * H * W
    pb = PP.SlabPeerBuffers(slab, 1, C, W, a.steps, dtype, dev)

    def run():
        return PP.slab_rollout_fwd_bwd_p2p(slab, pb, mine, Ks, sym, L, dt, a.steps, lambda f: 2.0 * f / n_total,
                                           keep_features=True)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    dist.barrier()
    ops.profile_read()
    ops.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize()
    kms, kn = ops.profile_read()
    ops.profile_enable(False)
    if rank == 0:
        print(json.dumps({"n_gpus": world, "grid": [H, W], "steps": a.steps, "ms_pass": e0.elapsed_time(e1),
                          "kernel_ms": kms, "launches": kn, "kernel_ms_sum": sum(kms.values())}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
