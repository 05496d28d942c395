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


def phase_times(slab, bufs, u_local, kernels, symnet, L, dt, steps, n_total, reps=3):
    """The statements of `parallel.slab_rollout_fwd_bwd_p2p` with a CUDA event between the phases (rank 0's view;
    the events serialise nothing, the stream is in order anyway)."""
    h, h2, Hs = slab.halo, 2 * slab.halo, slab.Hs
    B, C, _, W = u_local.shape
    names = ["copy_in", "exchange_state", "barrier0", "fwd_steps", "final_and_seed", "exchange_seed", "bwd_steps",
             "finish", "allreduce", "outputs"]
    acc = dict.fromkeys(names, 0.0)
    for _ in range(reps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
        feats = PP._slab_feature_buffers(True, steps, B, C, kernels.shape[0], Hs, W, u_local)
        dist.barrier()
        torch.cuda.synchronize()
        ev[0].record()
        bufs.state[0][:, :, h2:h2 + Hs] = u_local
        ev[1].record()
        bufs.barrier()
        bufs.fill_ghosts("state", 0, h2)
        ev[2].record()
        bufs.barrier()
        ev[3].record()
        for t in range(steps):
            up, dn = bufs.state_peers(t + 1)
            ops.slab_step(bufs.state[t], kernels, symnet, L, dt, h2, bufs.state[t + 1], h2, up, dn, h2, feats=feats[t])
            bufs.barrier()
        ev[4].record()
        final = bufs.state[steps][:, :, h2:h2 + Hs].contiguous()
        torch.mul(final, 2.0 / n_total, out=bufs.lam[0][:, :, h:h + Hs])
        ev[5].record()
        bufs.fill_ghosts("lam", 0, h)
        bufs.barrier()
        ev[6].record()
        bwd = ops.SlabBackward(kernels, symnet, L, dt, B, C, Hs, W)
        cur = 0
        for t in range(steps - 1, -1, -1):
            up, dn = bufs.lam_peers(cur ^ 1)
            bwd.step(bufs.state[t], bufs.lam[cur], bufs.lam[cur ^ 1], h, up, dn, h, feats=feats[t])
            bufs.barrier()
            cur ^= 1
        ev[7].record()
        gk, gs = bwd.finish()
        ev[8].record()
        PP.flat_allreduce([gk, gs], group=slab.group)
        ev[9].record()
        out = bufs.lam[cur][:, :, h:h + Hs].contiguous()
        ev[10].record()
        torch.cuda.synchronize()
        for i, n in enumerate(names):
            acc[n] += ev[i].elapsed_time(ev[i + 1]) / reps
        del out, final
    return acc


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
    phases = phase_times(slab, pb, mine, Ks, sym, L, dt, a.steps, n_total)
    if rank == 0:
        print(json.dumps({"n_gpus": world, "grid": [H, W], "steps": a.steps, "ms_pass": e0.elapsed_time(e1),
                          "kernel_ms": kms, "launches": kn, "kernel_ms_sum": sum(kms.values()),
                          "phase_ms": phases}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
