#!/usr/bin/env python
"""Benchmark of the PDE-Net time-stepping hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Metric: grid-point-steps/sec of one fwd+bwd rollout (B*H*W*blocks per pass / time).  One bench
"step" = one forward + reverse rollout of `blocks` delta-t blocks over the whole batch.
Workload (N=1 and N>1): BASELINE config 4 -- 2-D Burgers shape, 1024x1024 periodic grid, batch 256
trajectories (global; sharded B/N per GPU, strong scaling), 2 channels, 7x7 moment filters,
max_order 2 (6 filters), 2-layer SymNet, fp64, 8 delta-t blocks, loss = mean square of the final
state.  N>1: one process per GPU (torchrun), one NCCL all-reduce of the flat moment+SymNet
gradient buffer per pass.

`--impl reference` times the CPU oracle (the reference mount has no source: SURVEY.md section 0)
on the box's host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

CFG = dict(workload="cfg4: 2D Burgers shape 1024x1024 periodic, batch 256 (global), C=2, 7x7 filters, "
                    "max_order 2 (N=6), SymNet L=2, fp64, 8 dt-blocks fwd+bwd",
           B=256, C=2, H=1024, W=1024, k=7, max_order=2, L=2, blocks=8)
F_TOTAL = 4320.0        # algorithmic flop per grid-point-step fwd+bwd
# per-kernel useful flop per grid point (FMA = 2), recompute included:
KERNEL_FLOPS = {"fwd_step": 1440.0,        # stencils 1176 + SymNet 260 + Euler 4
                "bwd_kernel_A": 1968.0,    # stencil + SymNet recompute 1440, SymNet data + parameter gradient 2*264
                "bwd_kernel_B1": 1176.0,   # data gradient: C*N transposed 7x7 correlations
                "bwd_kernel_B2": 1176.0}   # filter gradient: C*N*49 multiply-adds
KERNEL_NAMES = {"fwd_step": "step_kernel<double,7,6,2,MODE_FWD> (fused stencils+SymNet+Euler)",
                "bwd_kernel_A": "step_kernel<double,7,6,2,MODE_BWD_A> (recompute + SymNet reverse)",
                "bwd_kernel_B1": "adjoint_kernel<double,7> (data gradient)",
                "bwd_kernel_B2": "filter_grad_kernel<double,7> (filter gradient)"}
# algorithmic HBM bytes per grid point of each kernel (fp64)
KERNEL_BYTES = {"fwd_step": 32.0, "bwd_kernel_A": 128.0, "bwd_kernel_B1": 128.0, "bwd_kernel_B2": 112.0}
# with kept features (ops.rollout keep_features): the forward step also writes the 12 feature planes,
# kernel A reads them back instead of recomputing the stencils
KEPT_FLOPS = {"bwd_kernel_A": 788.0}       # SymNet recompute 260 + data and parameter gradient 2*264
KEPT_BYTES = {"fwd_step": 128.0, "bwd_kernel_A": 208.0}
KEPT_NAMES = {"fwd_step": "step_kernel<double,7,6,2,MODE_FWD> (fused stencils+SymNet+Euler, features kept)",
              "bwd_kernel_A": "step_kernel<double,7,6,2,MODE_BWD_A_KEPT> (SymNet reverse on kept features)"}
BYTES_PER_POINT_STEP = 80.0  # 5*C*s
NOMINAL_FP64_TFLOPS = 37.0


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------- synthetic data
def model_params(k, max_order, C, L, dx, dtype=torch.float64):
    """Moments: one-hot + 1e-2*N(0,1) on the non-pinned entries (seed 0); SymNet ~ N(0, 0.1^2) (seed 1)."""
    from pdenet_b200.modules import order_list
    orders = order_list(max_order)
    N = len(orders)
    g = torch.Generator().manual_seed(0)
    M = torch.zeros(N, k, k, dtype=torch.float64)
    idx = torch.arange(k)
    for n, (o0, o1) in enumerate(orders):
        M[n, o0, o1] = 1
        free = (idx[:, None] + idx[None, :]) >= (o0 + o1 + 1)
        M[n] += 1e-2 * torch.randn(k, k, generator=g, dtype=torch.float64) * free
    g = torch.Generator().manual_seed(1)
    P = sum(2 * (C * N + i) + 2 for i in range(L)) + C * N + L + 1
    sym = 0.1 * torch.randn(C, P, generator=g, dtype=torch.float64)
    scales = [1.0 / (dx ** (o0 + o1)) for (o0, o1) in orders]
    return orders, M.to(dtype), scales, sym.to(dtype)


def synth_fields(b0, nb, C, H, W, device, dtype=torch.float64):
    """Random low-frequency periodic Fourier series (|p|,|q| <= 4), max|u0| = 1, seed 1234+131*b+c."""
    y = torch.arange(H, dtype=torch.float64, device=device) * (2 * math.pi / H)
    x = torch.arange(W, dtype=torch.float64, device=device) * (2 * math.pi / W)
    modes = [(p, q) for p in range(-4, 5) for q in range(-4, 5)]
    cy = torch.stack([torch.cos(q * y) for _, q in modes]), torch.stack([torch.sin(q * y) for _, q in modes])
    cx = torch.stack([torch.cos(p * x) for p, _ in modes]), torch.stack([torch.sin(p * x) for p, _ in modes])
    out = torch.empty(nb, C, H, W, dtype=dtype, device=device)
    for i in range(nb):
        for c in range(C):
            g = torch.Generator().manual_seed(1234 + 131 * (b0 + i) + c)
            a = torch.randn(len(modes), generator=g, dtype=torch.float64).to(device)
            b = torch.randn(len(modes), generator=g, dtype=torch.float64).to(device)
            # cos(px+qy) = cx cy - sx sy ; sin(px+qy) = sx cy + cx sy
            u = torch.einsum("m,mh,mw->hw", a, cy[0], cx[0]) - torch.einsum("m,mh,mw->hw", a, cy[1], cx[1]) \
                + torch.einsum("m,mh,mw->hw", b, cy[0], cx[1]) + torch.einsum("m,mh,mw->hw", b, cy[1], cx[0])
            out[i, c] = (u / u.abs().max()).to(dtype)
    return out


# ----------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except Exception:
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU oracle timing
def oracle_pass(sample, threads):
    """One fwd+bwd rollout of the CPU oracle on `sample`; returns grid-point-steps/s."""
    from oracle import pdenet_oracle as O
    torch.set_num_threads(threads)
    B, C, H, W, blocks = sample["B"], CFG["C"], sample["H"], sample["W"], sample["blocks"]
    dx = 2 * math.pi / W
    orders, M, scales, sym = model_params(CFG["k"], CFG["max_order"], C, CFG["L"], dx)
    u0 = synth_fields(0, B, C, H, W, "cpu")
    M = M.requires_grad_(True)
    sym = sym.requires_grad_(True)
    t0 = time.perf_counter()
    Ks = O.scaled_kernels(M, torch.tensor(scales, dtype=torch.float64))
    out = O.rollout(u0, Ks, sym, CFG["L"], 0.25 * dx * dx, blocks)[-1]
    out.square().mean().backward()
    dt = time.perf_counter() - t0
    return B * H * W * blocks / dt, dt


CPU_SAMPLE = dict(B=8, H=1024, W=1024, blocks=4)


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_baseline(threads):
    oracle_pass(dict(B=1, H=128, W=128, blocks=1), threads)  # warm-up (thread pool, allocator)
    v, dt = oracle_pass(CPU_SAMPLE, threads)
    return {"value": v, "unit": "grid-point-steps/s", "cores": threads, "cpu": cpu_model(), "kind": "port",
            "sample": "oracle fwd+bwd, B=%(B)d %(H)dx%(W)d, %(blocks)d dt-blocks, fp64, same model" % CPU_SAMPLE
                      + " (%.1f s)" % dt}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    for _ in range(args.warmup):
        oracle_pass(dict(B=1, H=128, W=128, blocks=1), threads)
    vals, t0 = [], time.perf_counter()
    for _ in range(args.steps):
        v, _ = oracle_pass(CPU_SAMPLE, threads)
        vals.append(v)
    total = time.perf_counter() - t0
    pts = CPU_SAMPLE["B"] * CPU_SAMPLE["H"] * CPU_SAMPLE["W"] * CPU_SAMPLE["blocks"]
    value = pts * args.steps / total
    line = {"impl": "reference", "metric": "grid-point-steps/sec fwd+bwd rollout", "value": value,
            "unit": "grid-point-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": CFG["workload"], "note": "reference mount has no source; CPU oracle (eager "
                       "PyTorch restatement) timed on a bounded sample, throughput is per grid-point-step"},
            "cpu_baseline": {"value": value, "unit": "grid-point-steps/s", "cores": threads, "cpu": cpu_model(),
                             "kind": "port", "sample": "oracle fwd+bwd, B=%(B)d %(H)dx%(W)d, %(blocks)d dt-blocks per step"
                                       % CPU_SAMPLE},
            "e2e": {"value": value, "unit": "grid-point-steps/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch.distributed as dist

    import pdenet_b200 as P
    from pdenet_b200 import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dtype = torch.float64 if args.dtype == "f64" else torch.float32
    B, C, H, W, k, L, blocks = (args.batch or CFG["B"]), CFG["C"], args.size or CFG["H"], args.size or CFG["W"], \
        CFG["k"], CFG["L"], args.blocks or CFG["blocks"]
    assert B % world == 0, "global batch must divide over the ranks"
    Bl = B // world
    dx = 2 * math.pi / W
    dt = 0.25 * dx * dx
    chunk = min(Bl, args.chunk)

    pass_chunk = Bl if args.pass_chunk <= 0 else min(Bl, args.pass_chunk)
    keep = args.keep == "on" and pass_chunk * blocks * C * len(P.order_list(CFG["max_order"])) * H * W \
        * (8 if args.dtype == "f64" else 4) <= ops.FEATURE_CACHE_BYTES
    kflops, kbytes, knames = dict(KERNEL_FLOPS), dict(KERNEL_BYTES), dict(KERNEL_NAMES)
    if keep:
        kflops.update(KEPT_FLOPS)
        kbytes.update(KEPT_BYTES)
        knames.update(KEPT_NAMES)
    model = P.POLYPDE2D(dt, dx, k, CFG["max_order"], "moment", "u,v", hidden_layers=L, batch_chunk=chunk,
                        dtype=dtype, keep_features=keep).to(dev)
    orders, M, scales, sym = model_params(k, CFG["max_order"], C, L, dx)
    with torch.no_grad():
        for fd, m in zip(model.fds, M):
            fd.moment.copy_(m)
        off = 0
        for c, pl in enumerate(model.polys):
            flat, o = sym[c], 0
            for prm in pl.parameters():
                prm.copy_(flat[o:o + prm.numel()].reshape(prm.shape))
                o += prm.numel()
    params = [p for p in model.parameters()]

    u0 = synth_fields(rank * Bl, Bl, C, H, W, dev, dtype)
    host_u0 = torch.empty(u0.shape, dtype=dtype, pin_memory=True)
    host_u0.copy_(u0)
    stage = torch.empty_like(u0)
    nparam = sum(p.numel() for p in params)
    host_out = torch.empty(nparam + 1, dtype=dtype, pin_memory=True)
    from pdenet_b200 import parallel

    def one_pass(inp):
        for p in params:
            p.grad = None
        # the shard goes through the rollout `pass_chunk` trajectories at a time (bounds the live
        # states + kept features); loss = mean square of the final states over the global batch
        loss = parallel.chunked_fwd_bwd(model, inp, blocks * dt, 1.0 / (inp[0].numel() * B), pass_chunk)
        model.diff_grad_proj()
        # one flat all-reduce of every gradient + the loss (no-op at world == 1)
        return parallel.allreduce_gradients(params, loss=loss)

    # chunks of the streamed e2e pass: at least four per shard so the H2D copy overlaps compute
    e2e_chunk = min(Bl, args.e2e_chunk if args.e2e_chunk > 0 else max(8, min(32, Bl // 2)))
    e2e_bufs = [stage[:e2e_chunk], torch.empty_like(stage[:e2e_chunk])]
    copy_stream = torch.cuda.Stream(dev)

    def one_pass_e2e():
        # H2D of this pass's inputs from pinned host memory, streamed in chunks of trajectories so
        # the copy of chunk k+1 overlaps the fwd+bwd rollout of chunk k (parallel.streamed_fwd_bwd)
        for p in params:
            p.grad = None
        loss = parallel.streamed_fwd_bwd(model, host_u0, blocks * dt, 1.0 / (u0[0].numel() * B), e2e_chunk,
                                         bufs=e2e_bufs, copy_stream=copy_stream)
        model.diff_grad_proj()
        total = parallel.allreduce_gradients(params, loss=loss)
        o = 0
        for p in params:                                  # D2H of the gradients + loss
            host_out[o:o + p.numel()].copy_(p.grad.reshape(-1), non_blocking=True)
            o += p.numel()
        host_out[nparam:].copy_(total.reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return host_out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # FP64 CUDA-core peak on this GPU, this run (register-only DFMA loop)
    fma_peak = ops.fma_peak(dtype, 2048)
    # operand-pattern probes: register-bank limit of a 3-vector-operand DFMA
    fma_probe = {"loop_invariant_operands": fma_peak / 1e12,
                 "three_distinct_vector_operands": ops.fma_peak(dtype, 2048, 1) / 1e12,
                 "one_operand_shared_by_4": ops.fma_peak(dtype, 2048, 2) / 1e12,
                 # the filter-gradient kernel's own register pattern (49 accumulators, g shared by 7
                 # FMAs, window rows from shared memory): the ceiling of kernel B2
                 "filter_gradient_pattern": ops.fma_peak(dtype, 512, 3) / 1e12}

    for _ in range(max(args.warmup, 3)):
        one_pass(u0)
    one_pass_e2e()
    barrier()
    ops.profile_read()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- timed region 1: inputs resident in HBM
    ops.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    e0.record()
    marks[0].record()
    for i in range(args.steps):
        one_pass(u0)
        marks[i + 1].record()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    per_step = sorted(marks[i].elapsed_time(marks[i + 1]) for i in range(args.steps))  # this rank's passes
    kind_ms, kind_n = ops.profile_read()
    ops.profile_enable(False)
    # ---- timed region 2: end to end (pinned host inputs -> device, loss+grads -> host)
    barrier()
    e0.record()
    for _ in range(args.steps):
        one_pass_e2e()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    _, kind_n_e2e = ops.profile_read()

    if world > 1:
        t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pts = B * H * W * blocks  # grid-point-steps per pass, whole job
    value = pts * args.steps / (ms * 1e-3)
    e2e = pts * args.steps / (ms_e2e * 1e-3)
    hbm_peak, hbm_src = peaks()
    # dominant kernel = the kind with the largest share of the timed region
    top = max(kflops, key=lambda kk: kind_ms[kk])
    n_top = max(1, kind_n[top])
    ms_top = kind_ms[top] / n_top
    es = 8 if args.dtype == "f64" else 4

    def launch_pts(kk):  # grid points one launch of this kernel covers
        return (pass_chunk if kk == "fwd_step" else min(chunk, pass_chunk)) * H * W

    pts_per_launch = launch_pts(top)
    achieved = kflops[top] * pts_per_launch / (ms_top * 1e-3) / 1e12
    per_kernel = {}
    for kk in kflops:
        if kind_ms[kk] <= 0:
            continue
        sec = kind_ms[kk] / max(1, kind_n[kk]) * 1e-3
        per_kernel[kk] = {"ms_per_launch": sec * 1e3, "launches": kind_n[kk],
                          "tflops": kflops[kk] * launch_pts(kk) / sec / 1e12,
                          "frac_of_fma_peak": kflops[kk] * launch_pts(kk) / sec / fma_peak,
                          "algorithmic_gbs": kbytes[kk] * es / 8 * launch_pts(kk) / sec / 1e9,
                          "frac_of_hbm_peak": kbytes[kk] * es / 8 * launch_pts(kk) / sec / 1e9 / hbm_peak}
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(top)
        except Exception:
            traffic = None
    total_kernel_ms = sum(kind_ms.values())
    line = {
        "metric": "grid-point-steps/sec fwd+bwd rollout", "value": value, "unit": "grid-point-steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
        "ms_per_step_best": per_step[0], "ms_per_step_median": statistics.median(per_step),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": CFG["workload"] if args.dtype == "f64" else CFG["workload"].replace("fp64", "fp32"),
                   "global_batch": B, "per_gpu_batch": Bl, "grid": [H, W],
                   "dt_blocks": blocks, "max_order": CFG["max_order"], "dt": dt, "batch_chunk": chunk,
                   "pass_chunk": pass_chunk, "features": "kept in HBM for the reverse pass" if keep else "recomputed",
                   "parallelism": "dp%d (batch-sharded, one NCCL all-reduce of %d grad scalars)" % (world, nparam + 1),
                   "l2": "inputs (%.1f GB per pass) larger than L2; no flush needed" % (u0.numel() * 8 / 1e9)},
        "e2e": {"value": e2e, "unit": "grid-point-steps/s", "h2d_bytes_per_step": u0.numel() * u0.element_size(),
                "d2h_bytes_per_step": (nparam + 1) * 8, "ms_per_step": ms_e2e / args.steps,
                "how": "pinned host batch streamed in chunks of up to %d trajectories (first chunks smaller: %s); H2D of "
                       "chunk k+1 overlaps fwd+bwd of chunk k; loss+gradients copied back to pinned host memory"
                       % (e2e_chunk, parallel.ramp_chunks(Bl, e2e_chunk)[:3])},
        "gpu_launches": int(sum(kind_n.values())),
        "clocks": clocks,
        "roofline": {
            "bound": "fp64" if args.dtype == "f64" else "fp32", "kernel": knames[top].replace("double", "double" if args.dtype == "f64" else "float"),
            "achieved": achieved, "peak": fma_peak / 1e12, "unit": "TFLOP/s", "frac": achieved / (fma_peak / 1e12),
            "peak_source": "measured this run: pdenet_fma_peak register-only DFMA loop (nominal %.0f)" % NOMINAL_FP64_TFLOPS,
            "flops_per_launch": kflops[top] * pts_per_launch, "ms_per_launch": ms_top, "launches": n_top,
            "traffic": traffic, "per_kernel": per_kernel, "fma_probe_tflops": fma_probe,
            "kernel_ms_share": {kk: (v / total_kernel_ms if total_kernel_ms else None) for kk, v in kind_ms.items()},
            "whole_pass": {"flop_per_point_step": F_TOTAL, "achieved_tflops": F_TOTAL * value / 1e12 / world,
                           "frac_of_fp64_peak": F_TOTAL * value / world / fma_peak,
                           # the peak probe runs for milliseconds at the full boost clock; a long run can
                           # sit at the board's power limit (clocks.reasons): same fraction against the
                           # peak scaled to the median SM clock sampled during the timed region
                           "frac_of_fp64_peak_at_sampled_clock": (
                               F_TOTAL * value / world / (fma_peak * clocks["sm_mhz"] / clocks["sm_max_mhz"])
                               if clocks and clocks.get("sm_mhz") and clocks.get("sm_max_mhz") else None),
                           "hbm_gbs": BYTES_PER_POINT_STEP * value / 1e9 / world,
                           "frac_of_hbm_peak": BYTES_PER_POINT_STEP * value / 1e9 / world / hbm_peak,
                           "hbm_peak_source": hbm_src},
        },
    }
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(os.cpu_count() or 1)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="override the global batch (parity/debug only)")
    ap.add_argument("--size", type=int, default=0, help="override the grid side (debug only)")
    ap.add_argument("--blocks", type=int, default=0, help="override the number of dt-blocks (debug only)")
    ap.add_argument("--chunk", type=int, default=32, help="trajectories per reverse-kernel launch")
    ap.add_argument("--pass-chunk", type=int, default=32,
                    help="trajectories per forward+reverse rollout of the resident pass (0 = the whole shard)")
    ap.add_argument("--keep", choices=["on", "off"], default="on",
                    help="keep the derivative features of each chunk for the reverse pass (off = recompute)")
    ap.add_argument("--e2e-chunk", type=int, default=0,
                    help="largest streamed chunk of the e2e pass (0 = shard/2, clamped to [8, 32])")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"], help="f64 = the reference's precision (default)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
