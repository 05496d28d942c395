"""torchrun --nproc-per-node N tools/dp_multigpu_check.py [--size 128] [--batch 8] [--steps 3]

Batch-sharded data parallelism over N GPUs (DIST-2): every rank runs the fused fwd+bwd rollout on its
shard through the model API, the flat gradient buffer is all-reduced in place (NCCL), and the result
is compared on every rank with the whole batch run on one GPU."""
import argparse
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import pdenet_b200 as P
from bench import build_model, synth_fields
from pdenet_b200 import parallel as PP


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--dtype", default="f64")
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    dtype = torch.float64 if a.dtype == "f64" else torch.float32
    H = W = a.size
    model, flat = build_model(P, dev, dtype, H, W, 0, True)
    u_all = synth_fields(0, a.batch, 2, H, W, dev, dtype)
    scale = 1.0 / (u_all[0].numel() * a.batch)
    T = a.steps * model.dt
    # whole batch on this GPU alone
    flat.zero_grad()
    ref_loss = PP.chunked_fwd_bwd(model, u_all, T, scale, a.batch)
    ref = flat.grad_flat().clone()
    # this rank's shard + one in-place all-reduce of the flat gradient buffer (loss in its last slot)
    b0, nb = PP.shard_range(a.batch, rank, world)
    flat.zero_grad()
    ptr = flat.gflat.data_ptr()
    loss = PP.chunked_fwd_bwd(model, u_all[b0:b0 + nb], T, scale, max(1, nb))
    flat.grad_flat()
    total = PP.allreduce_gradients(flat, loss=loss)
    assert flat.gflat.data_ptr() == ptr
    err = torch.stack([(flat.gflat - ref).abs().max() / ref.abs().max(), (total - ref_loss).abs() / ref_loss.abs()])
    dist.all_reduce(err, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"what": "batch-sharded fwd+bwd, in-place flat all-reduce", "n_gpus": world, "grid": [H, W],
                          "batch": a.batch, "steps": a.steps, "dtype": a.dtype,
                          "max_rel_err_vs_single_gpu": {"grad": err[0].item(), "loss": err[1].item()}}))
    tol = 1e-10 * a.steps if dtype == torch.float64 else 1e-4
    assert err.max().item() < tol, err
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
