#!/usr/bin/env python
"""Benchmark of the PDE-Net time-stepping hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Metric: grid-point-steps/sec of one fwd+bwd rollout (B*H*W*blocks per pass / time).  One bench
"step" = one forward + reverse rollout of `blocks` delta-t blocks over the whole batch.
Workload (N=1 and N>1): BASELINE config 4 -- 2-D Burgers shape, 1024x1024 periodic grid, batch 256
trajectories (global; sharded B/N per GPU, strong scaling), 2 channels, 7x7 moment filters,
max_order 2 (6 filters), 2-layer SymNet, fp64, 8 delta-t blocks, loss = mean square of the final
state.  N>1: one process per GPU (torchrun), one in-place NCCL all-reduce of the flat moment+SymNet
gradient buffer per pass.

Sub-records of the one JSON line (rank 0):
  parity         CUDA path vs the CPU oracle on the cpu_baseline sample (8 trajectories of the headline
                 grid, 4 blocks): relative errors of the final state and of every gradient       (N = 1)
  fp32           the same workload in single precision (shorter run), fraction of the FP32 FMA peak (N = 1)
  small_configs  BASELINE configs 1-3 fwd+bwd (eager and as one CUDA graph), fraction of the FMA peak (N = 1)
  cfg5_slab      BASELINE config 5: one 8192x8192 trajectory, slab decomposition over the N GPUs with the
                 halo exchange fused into the step kernels (peer stores over NVLink), fp64 and fp32, checked
                 against the single-GPU rollout of the same grid                                   (N > 1)

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
# nominal CUDA-core peaks: 148 SMs x (64 | 128) FMA/clk x 2 flop x 1.965 GHz
NOMINAL_TFLOPS = {"f64": 37.2, "f32": 74.5}
F_TOTAL_LINEAR_CFG2 = 2496.0   # PDE-Net 1 linear model, 5x5, N=15 (+D0): 3 * 832


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------- synthetic data
def burgers_symnet(nu, N=6, C=2, L=2):
    """SymNet weights of viscous Burgers, channel c -> -u w_x - v w_y + nu (w_xx + w_yy) (flat C-ABI layout;
    features f = c*N + n, filters (0,0),(0,1),(1,0),(0,2),(1,1),(2,0), (0,1) = d/dx along W)."""
    m = C * N
    P = sum(2 * (m + i) + 2 for i in range(L)) + m + L + 1
    sym = torch.zeros(C, P, dtype=torch.float64)
    for c in range(C):
        base = c * N
        W0 = torch.zeros(2, m, dtype=torch.float64)
        W0[0, 0] = 1
        W0[1, base + 1] = 1
        W1 = torch.zeros(2, m + 1, dtype=torch.float64)
        W1[0, N] = 1
        W1[1, base + 2] = 1
        wf = torch.zeros(m + L, dtype=torch.float64)
        wf[m] = wf[m + 1] = -1
        wf[base + 3] = wf[base + 5] = nu
        z2 = torch.zeros(2, dtype=torch.float64)
        sym[c] = torch.cat([W0.flatten(), z2, W1.flatten(), z2, wf, torch.zeros(1, dtype=torch.float64)])
    return sym


def model_params(k, max_order, C, L, dx, dtype=torch.float64, nu=0.05, noise=1e-2, moment_noise=1e-3, seed=0):
    """Synthetic model of every bench / parity run: viscous Burgers + noise*N(0,1) on every SymNet entry and
    moment_noise*N(0,1) on every non-pinned moment entry (all gradients non-trivial; forward Euler stable at
    dt = dx^2).  Same numbers as oracle.stable_burgers_model (tests/test_cabi_cpu.py checks it); a fully
    random SymNet is anti-diffusive half of the time and blows up on 128^2 and finer grids."""
    from pdenet_b200.modules import order_list
    assert (max_order, C, L) == (2, 2, 2)
    orders = order_list(max_order)
    g = torch.Generator().manual_seed(seed)
    M = torch.zeros(len(orders), k, k, dtype=torch.float64)
    idx = torch.arange(k)
    for n, (o0, o1) in enumerate(orders):
        M[n, o0, o1] = 1
        free = (idx[:, None] + idx[None, :]) >= (o0 + o1 + 1)
        M[n] += moment_noise * torch.randn(k, k, generator=g, dtype=torch.float64) * free
    P = sum(2 * (C * len(orders) + i) + 2 for i in range(L)) + C * len(orders) + L + 1
    sym = burgers_symnet(nu) + noise * torch.randn(C, P, generator=g, dtype=torch.float64)
    scales = [1.0 / (dx ** (o0 + o1)) for (o0, o1) in orders]
    return orders, M.to(dtype), scales, sym.to(dtype)


def step_dt(dx):
    """dt of every bench / parity rollout: dx^2 -- diffusion number nu dt ||K_xx||_1 = 0.05 * 6.04 = 0.3 per
    direction, so forward Euler is stable with margin for the noise terms (rounding noise is not amplified step
    over step, in fp32 either: 2.5 dx^2 already grows over 15 steps at 128^2) and dt*F is ~1e-4 of the state."""
    return dx * dx


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
def oracle_pass(sample, threads, keep=False):
    """One fwd+bwd rollout of the CPU oracle on `sample`; returns (grid-point-steps/s, seconds, results).
    `results` (keep=True): final state and the gradients w.r.t. u0 / moments / SymNet, for the parity leg."""
    from oracle import pdenet_oracle as O
    torch.set_num_threads(threads)
    B, C, H, W, blocks = sample["B"], CFG["C"], sample["H"], sample["W"], sample["blocks"]
    dx = 2 * math.pi / W
    orders, M, scales, sym = model_params(CFG["k"], CFG["max_order"], C, CFG["L"], dx)
    u0 = synth_fields(0, B, C, H, W, "cpu").requires_grad_(keep)
    M = M.requires_grad_(True)
    sym = sym.requires_grad_(True)
    t0 = time.perf_counter()
    Ks = O.scaled_kernels(M, torch.tensor(scales, dtype=torch.float64))
    if keep:
        Ks.retain_grad()
    out = O.rollout(u0, Ks, sym, CFG["L"], step_dt(dx), blocks)[-1]
    out.square().mean().backward()
    dt = time.perf_counter() - t0
    res = None
    if keep:
        pinned = torch.stack([O.pinned_mask(CFG["k"], o, "moment") for o in orders])
        # dM = s Q^T dK Q with Q = A^-1 (cond(A) = 294 at k = 7) cancels heavily: the componentwise bound
        # s |Q|^T |dK| |Q| is the scale a backward-stable evaluation of grad_moment is accurate to
        Q = torch.linalg.inv(O.moment_axis_matrix(CFG["k"])).abs()
        bound = torch.tensor(scales, dtype=torch.float64)[:, None, None] * (Q.t() @ Ks.grad.abs() @ Q)
        res = {"state": out.detach(), "grad_u0": u0.grad, "grad_kernels": Ks.grad, "grad_symnet": sym.grad,
               "grad_moment": M.grad.masked_fill(pinned, 0), "grad_moment_scale": bound.masked_fill(pinned, 0).max()}
    return B * H * W * blocks / dt, dt, res


CPU_SAMPLE = dict(B=8, H=1024, W=1024, blocks=4)


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_baseline(threads, keep=False):
    oracle_pass(dict(B=1, H=128, W=128, blocks=1), threads)  # warm-up (thread pool, allocator)
    v, dt, res = oracle_pass(CPU_SAMPLE, threads, keep)
    return {"value": v, "unit": "grid-point-steps/s", "cores": threads, "cpu": cpu_model(), "kind": "port",
            "sample": "oracle fwd+bwd, B=%(B)d %(H)dx%(W)d, %(blocks)d dt-blocks, fp64, same model" % CPU_SAMPLE
                      + " (%.1f s)" % dt}, res


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    for _ in range(args.warmup):
        oracle_pass(dict(B=1, H=128, W=128, blocks=1), threads)
    vals, t0 = [], time.perf_counter()
    for _ in range(args.steps):
        v, _, _ = oracle_pass(CPU_SAMPLE, threads)
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
def build_model(P, dev, dtype, H, W, batch_chunk, keep):
    """POLYPDE2D with the synthetic Burgers parameters, parameters / gradients in flat buffers."""
    dx = 2 * math.pi / W
    model = P.POLYPDE2D(step_dt(dx), dx, CFG["k"], CFG["max_order"], "moment", "u,v", hidden_layers=CFG["L"],
                        batch_chunk=batch_chunk, dtype=dtype, keep_features=keep).to(dev)
    orders, M, scales, sym = model_params(CFG["k"], CFG["max_order"], CFG["C"], CFG["L"], dx)
    with torch.no_grad():
        for fd, m in zip(model.fds, M):
            fd.moment.copy_(m)
        for c, pl in enumerate(model.polys):
            flat, o = sym[c], 0
            for prm in pl.parameters():
                prm.copy_(flat[o:o + prm.numel()].reshape(prm.shape))
                o += prm.numel()
    return model, P.FlatParams(model)


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-300)).item()


def parity_record(P, dev, ref, chunk):
    """The CPU-baseline sample (8 trajectories of the headline grid, 4 blocks) through the CUDA path (public
    ops: m2k -> rollout -> loss -> reverse), fp64 and fp32, against the oracle results `ref` of the same
    inputs / model / dt."""
    B, H, W, blocks = CPU_SAMPLE["B"], CPU_SAMPLE["H"], CPU_SAMPLE["W"], CPU_SAMPLE["blocks"]
    dx = 2 * math.pi / W
    orders, M, scales, sym = model_params(CFG["k"], CFG["max_order"], CFG["C"], CFG["L"], dx)
    idx = torch.arange(CFG["k"])
    pinned = torch.stack([(idx[:, None] + idx[None, :]) < (o0 + o1 + 1) for (o0, o1) in orders]).to(dev)
    rec = {"sample": "B=%d %dx%d, %d dt-blocks, dt = dx^2, loss = mean square of the final state; "
                     "max|cuda - oracle| / max|oracle|" % (B, H, W, blocks),
           "tolerance": {"f64": 1e-10 * blocks, "f32_state": 1e-5 * blocks,
                         "f32_gradients": "1e-5 * blocks * max(1, max_n ||K_n||_1 / 24) = %.2g: conditioning of the "
                                          "1/dx^2-scaled taps, DESIGN.md section 0; the oracle itself run in fp32 "
                                          "is no closer (profiles/r2_fp32_parity.json)"
                                          % (1e-5 * blocks * max(1.0, 6.04 / dx ** 2 / 24)),
                         "grad_moment": "dM = s A^-T dK A^-1 cancels (cond(A) = 294): judged against the componentwise "
                                        "bound s |A^-1|^T |dK| |A^-1| (grad_moment_backward), not against max|dM|"}}
    for name, dtype in (("f64", torch.float64), ("f32", torch.float32)):
        Md = M.to(dev, dtype).requires_grad_(True)
        sd = sym.to(dev, dtype).requires_grad_(True)
        u0 = synth_fields(0, B, CFG["C"], H, W, dev, dtype).requires_grad_(True)
        Ks = P.m2k(Md, scales, pinned)
        Ks.retain_grad()
        out = P.rollout(u0, Ks, sd, CFG["L"], step_dt(dx), blocks, last_only=True, batch_chunk=chunk)
        out.square().mean().backward()
        got = {"state": out, "grad_u0": u0.grad, "grad_kernels": Ks.grad, "grad_symnet": sd.grad, "grad_moment": Md.grad}
        rec[name] = {kk: rel_err(got[kk], ref[kk]) for kk in got}
        rec[name]["grad_moment_backward"] = ((Md.grad.double().cpu() - ref["grad_moment"]).abs().max()
                                             / ref["grad_moment_scale"]).item()
        del Md, sd, u0, Ks, out, got
    tol = rec["tolerance"]["f64"]
    rec["ok"] = all(rec["f64"][kk] < tol for kk in ("state", "grad_u0", "grad_kernels", "grad_symnet",
                                                     "grad_moment_backward")) and \
        rec["f32"]["state"] < rec["tolerance"]["f32_state"]
    return rec


def time_fn(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def time_graph(fn, reps=50):
    """Capture one fwd+bwd into a CUDA graph and time the replays (host launch overhead removed)."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return time_fn(g.replay, reps, warm=2)


def small_configs_record(P, dev, peaks_tflops):
    """BASELINE configs 1-3 fwd+bwd through the public ops (eager, and captured as one CUDA graph)."""
    from pdenet_b200.modules import order_list
    out = []

    def symnet_cfg(name, B, H, W, steps, dtype, tag):
        dx = 2 * math.pi / W
        dt = step_dt(dx)
        orders, M, scales, sym = model_params(CFG["k"], 2, 2, 2, dx)
        M = M.to(dev, dtype).requires_grad_(True)
        sym = sym.to(dev, dtype).requires_grad_(True)
        u0 = synth_fields(0, B, 2, H, W, dev, dtype)

        def fn():
            o = P.rollout(u0, P.m2k(M, scales), sym, 2, dt, steps, last_only=True)
            return torch.autograd.grad(o.square().mean(), (M, sym))

        ms, msg = time_fn(fn, 20), time_graph(fn)
        pts = B * H * W * steps
        # the library's own training closure (losses.GraphedTrainingLoss: POLYPDE2D rollout, fused MSE against
        # observed states of every block, reverse pass, flat gradient -- captured once INSIDE the library, the
        # caller only replays): no user-side graph
        from pdenet_b200 import losses
        model, _ = build_model(P, dev, dtype, H, W, 0, True)
        observed = torch.stack([u0] * steps)
        closure = losses.GraphedTrainingLoss(model, u0, observed)
        msc = time_fn(closure, 50)
        out.append({"config": name, "dtype": tag, "ms_fwd_bwd": ms, "ms_fwd_bwd_cuda_graph": msg,
                    "ms_library_graphed_training_closure": msc,
                    "grid_point_steps_per_s": pts / (ms * 1e-3), "grid_point_steps_per_s_cuda_graph": pts / (msg * 1e-3),
                    "frac_of_fma_peak": F_TOTAL * pts / (ms * 1e-3) / (peaks_tflops[tag] * 1e12),
                    "frac_of_fma_peak_cuda_graph": F_TOTAL * pts / (msg * 1e-3) / (peaks_tflops[tag] * 1e12)})
        del closure, model

    def linear_cfg(name, B, H, W, steps, k=5, max_order=4):
        dx = 2 * math.pi / W
        dt = 1.5e-2 * dx * dx
        g = torch.Generator().manual_seed(0)
        orders = [o for o in order_list(max_order) if o != (0, 0)]
        N = len(orders)
        M = torch.zeros(N + 1, k, k, dtype=torch.float64)
        M[0, 0, 0] = 1
        for n, (a, b) in enumerate(orders):
            M[n + 1, a, b] = 1
        M = (M + 1e-3 * torch.randn(M.shape, generator=g, dtype=torch.float64)).to(dev).requires_grad_(True)
        scales = [1.0] + [1.0 / dx ** (a + b) for a, b in orders]
        coe = torch.randn(N, H, W, generator=g, dtype=torch.float64).to(dev).requires_grad_(True)
        u0 = synth_fields(0, B, 1, H, W, dev)[:, 0].contiguous()

        def fn():
            ks = P.m2k(M, scales)
            o = P.linear_rollout(u0, ks[0], ks[1:], coe, dt, steps, last_only=True)
            return torch.autograd.grad(o.square().mean(), (M, coe))

        ms, msg = time_fn(fn, 20), time_graph(fn)
        pts = B * H * W * steps
        out.append({"config": name, "dtype": "f64", "ms_fwd_bwd": ms, "ms_fwd_bwd_cuda_graph": msg,
                    "grid_point_steps_per_s": pts / (ms * 1e-3), "grid_point_steps_per_s_cuda_graph": pts / (msg * 1e-3),
                    "frac_of_fma_peak": F_TOTAL_LINEAR_CFG2 * pts / (ms * 1e-3) / (peaks_tflops["f64"] * 1e12),
                    "frac_of_fma_peak_cuda_graph": F_TOTAL_LINEAR_CFG2 * pts / (msg * 1e-3) / (peaks_tflops["f64"] * 1e12)})

    symnet_cfg("cfg1: Burgers 100x100, B=8, 1 dt-block", 8, 100, 100, 1, torch.float64, "f64")
    linear_cfg("cfg2: PDE-Net 1 linear 50x50, 5x5, N=15, B=16, 20 dt-blocks", 16, 50, 50, 20)
    symnet_cfg("cfg3: 2-component 128x128, B=64, 15 dt-blocks", 64, 128, 128, 15, torch.float64, "f64")
    symnet_cfg("cfg3 in fp32", 64, 128, 128, 15, torch.float32, "f32")
    return out


def slab_record(P, dev, rank, world, args):
    """BASELINE config 5 under the driver: one 8192^2 trajectory split into row slabs over the ranks; the halo
    exchange is fused into the step kernels (peer stores into the neighbours' ghost rows over NVLink,
    symmetric memory), features of the owned rows kept.  Rank 0 also runs the same grid on its GPU alone: the
    reference for the error check and for the efficiency."""
    import torch.distributed as dist

    from pdenet_b200 import parallel as PP
    size, blocks, reps = args.slab_size, args.slab_blocks, 3
    k, C, L, r = CFG["k"], CFG["C"], CFG["L"], (CFG["k"] - 1) // 2
    rec = {"workload": "cfg5: one %dx%d trajectory, C=2, 7x7, N=6, L=2, %d dt-blocks fwd+bwd, rows split over %d GPUs"
                       % (size, size, blocks, world)}
    for tag, dtype in (("f64", torch.float64), ("f32", torch.float32)):
        H = W = size
        dx = 2 * math.pi / W
        dt = step_dt(dx)
        orders, M, scales, sym = model_params(k, 2, C, L, dx)
        Ks = P.m2k(M.to(dev, dtype), scales)
        sym = sym.to(dev, dtype)
        slab = PP.SlabDecomposition(H, r, rank, world)
        mine = synth_fields(0, 1, C, H, W, dev, dtype)[:, :, slab.row0:slab.row0 + slab.Hs].contiguous()
        n_total = C * H * W

        def seed(f, out=None):  # loss = mean square of the final state over the WHOLE grid
            return torch.mul(f, 2.0 / n_total, out=out)

        how = "fused peer stores (symmetric memory over NVLink), features kept"
        pb = None
        try:
            pb = PP.SlabPeerBuffers(slab, 1, C, W, blocks, dtype, dev)

            pb.input_view().copy_(mine)      # the slab lives in the symmetric buffer: used in place every pass

            def run():
                return PP.slab_rollout_fwd_bwd_p2p(slab, pb, pb.input_view(), Ks, sym, L, dt, blocks, seed,
                                                   keep_features=True, copy_outputs=False)
            fin, gu, gk, gs = run()
        except Exception as e:  # noqa: BLE001 -- symmetric memory unavailable: NCCL send/recv halos
            how = "NCCL send/recv halo exchange, features kept (symmetric memory unavailable: %s)" % type(e).__name__

            def run():
                return slab.rollout_fwd_bwd(mine, Ks, sym, L, dt, blocks, seed, keep_features=True)
            fin, gu, gk, gs = run()
        torch.cuda.synchronize()
        # ---- reference: the same grid on rank 0's GPU alone (periodic rollout, features kept)
        ref_fin = torch.empty(1, C, H, W, dtype=dtype, device=dev)
        ref_gu = torch.empty_like(ref_fin)
        ref_gk, ref_gs = torch.empty_like(gk), torch.empty_like(gs)
        single_ms = torch.zeros(1, device=dev)
        if rank == 0:
            u_full = synth_fields(0, 1, C, H, W, dev, dtype)

            ud, Kd, sd = [t.detach().clone().requires_grad_(True) for t in (u_full, Ks, sym)]

            def single():
                o = P.rollout(ud, Kd, sd, L, dt, blocks, last_only=True, batch_chunk=1)
                return (o.detach(),) + torch.autograd.grad(o.square().mean(), (ud, Kd, sd))
            a, b, c_, d = single()
            ref_fin.copy_(a), ref_gu.copy_(b), ref_gk.copy_(c_), ref_gs.copy_(d)
            del a, b, c_, d
            single_ms[0] = time_fn(single, reps, warm=1)
            del u_full, ud
        for t in (ref_fin, ref_gu, ref_gk, ref_gs, single_ms):
            dist.broadcast(t, 0)
        sl = slice(slab.row0, slab.row0 + slab.Hs)

        def rel(x, y):
            return ((x.double() - y.double()).abs().max() / y.double().abs().max()).item()

        errs = torch.tensor([rel(fin, ref_fin[:, :, sl]), rel(gu, ref_gu[:, :, sl]), rel(gk, ref_gk), rel(gs, ref_gs)],
                            device=dev, dtype=torch.float64)
        # relative to the whole field's maximum, as the single-domain comparison is
        errs[0] = (fin.double() - ref_fin[:, :, sl].double()).abs().max() / ref_fin.double().abs().max()
        errs[1] = (gu.double() - ref_gu[:, :, sl].double()).abs().max() / ref_gu.double().abs().max()
        dist.all_reduce(errs, op=dist.ReduceOp.MAX)
        del ref_fin, ref_gu
        # ---- timing: max over ranks of the device time of `reps` passes
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            run()
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        es = 8 if tag == "f64" else 4
        pts = H * W * blocks
        v, v1 = pts / (ms.item() * 1e-3), pts / (single_ms.item() * 1e-3)
        rec[tag] = {"ms_per_pass": ms.item(), "grid_point_steps_per_s": v, "halo_exchange": how,
                    "single_gpu_same_grid": {"ms_per_pass": single_ms.item(), "grid_point_steps_per_s": v1},
                    "efficiency_vs_single_gpu": v / (world * v1),
                    "halo_bytes_per_step_per_gpu": {"forward (2r state rows to each neighbour)": 2 * 2 * r * W * C * es,
                                                    "reverse (r adjoint rows to each neighbour)": 2 * r * W * C * es},
                    "max_rel_err_vs_single_gpu": dict(zip(["state", "grad_u0", "grad_kernels", "grad_symnet"],
                                                          errs.tolist())),
                    "tolerance": 1e-10 * blocks if tag == "f64" else 1e-5 * blocks}
        del pb, fin, gu
        torch.cuda.empty_cache()
    return rec


def resident_and_e2e(P, dev, args, dtype_tag, B, blocks, steps, warmup, world, rank, local, sample_clocks):
    """The two timed regions of the cfg4 workload in `dtype_tag`.  Returns a dict of raw measurements."""
    import torch.distributed as dist

    from pdenet_b200 import ops, parallel
    dtype = torch.float64 if dtype_tag == "f64" else torch.float32
    es = 8 if dtype_tag == "f64" else 4
    C, H, W = CFG["C"], args.size or CFG["H"], args.size or CFG["W"]
    Bl = B // world
    chunk = min(Bl, args.chunk)
    pass_chunk = Bl if args.pass_chunk <= 0 else min(Bl, args.pass_chunk)
    keep = args.keep == "on" and pass_chunk * blocks * C * len(P.order_list(CFG["max_order"])) * H * W * es \
        <= ops.FEATURE_CACHE_BYTES
    model, flat = build_model(P, dev, dtype, H, W, chunk, keep)
    dt = model.dt
    u0 = synth_fields(rank * Bl, Bl, C, H, W, dev, dtype)
    host_u0 = torch.empty(u0.shape, dtype=dtype, pin_memory=True)
    host_u0.copy_(u0)
    stage = torch.empty_like(u0)
    nparam = int(flat.numel)
    host_out = torch.empty(nparam + 1, dtype=dtype, pin_memory=True)
    scale = 1.0 / (u0[0].numel() * B)

    def one_pass(inp):
        flat.zero_grad()
        # the shard goes through the rollout `pass_chunk` trajectories at a time (bounds the live
        # states + kept features); loss = mean square of the final states over the global batch
        loss = parallel.chunked_fwd_bwd(model, inp, blocks * dt, scale, pass_chunk)
        flat.grad_flat()
        # ONE in-place all-reduce of the flat gradient buffer, the loss in its last slot (no-op at world == 1)
        return flat.allreduce(loss)

    # chunks of the streamed e2e pass: at least four per shard so the H2D copy overlaps compute
    e2e_chunk = min(Bl, args.e2e_chunk if args.e2e_chunk > 0 else max(8, min(32, Bl // 2)))
    e2e_bufs = [stage[:e2e_chunk], torch.empty_like(stage[:e2e_chunk])]
    copy_stream = torch.cuda.Stream(dev)
    # every pass copies every chunk of the pinned host batch once; the copy of chunk k+1 overlaps the fwd+bwd rollout
    # of chunk k, and the first chunk of the NEXT pass is copied behind the last chunk of this one (the next
    # evaluation's inputs are the same batch), so in steady state no copy is exposed (parallel.StreamedBatch)
    streamed = parallel.StreamedBatch(host_u0, e2e_chunk, dev, bufs=e2e_bufs, copy_stream=copy_stream,
                                      prefetch_next=True)

    def one_pass_e2e():
        flat.zero_grad()
        loss = streamed.run(model, blocks * dt, scale)
        flat.grad_flat()
        flat.allreduce(loss)
        host_out.copy_(flat.grad_and_loss, non_blocking=True)   # D2H: every gradient + the loss, one copy
        torch.cuda.current_stream().synchronize()
        return host_out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        one_pass(u0)
    one_pass_e2e()
    barrier()
    ops.profile_read()
    sampler = ClockSampler(local) if (sample_clocks and rank == 0) else None
    if sampler:
        sampler.start()
    # ---- timed region 1: inputs resident in HBM
    ops.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    e0.record()
    marks[0].record()
    for i in range(steps):
        one_pass(u0)
        marks[i + 1].record()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    per_step = sorted(marks[i].elapsed_time(marks[i + 1]) for i in range(steps))  # this rank's passes
    kind_ms, kind_n = ops.profile_read()
    ops.profile_enable(False)
    # ---- timed region 2: end to end (pinned host inputs -> device, loss+grads -> host)
    barrier()
    e0.record()
    for _ in range(steps):
        one_pass_e2e()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    if world > 1:
        t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
    return dict(ms=ms, ms_e2e=ms_e2e, per_step=per_step, kind_ms=kind_ms, kind_n=kind_n, clocks=clocks, keep=keep,
                chunk=chunk, pass_chunk=pass_chunk, e2e_chunk=e2e_chunk, Bl=Bl, H=H, W=W, dt=dt, nparam=nparam,
                h2d=u0.numel() * u0.element_size(), d2h=(nparam + 1) * es,
                ramp=parallel.ramp_chunks(Bl, e2e_chunk)[:3])


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
    tag = args.dtype
    B, blocks = (args.batch or CFG["B"]), args.blocks or CFG["blocks"]
    assert B % world == 0, "global batch must divide over the ranks"
    warmup = max(args.warmup, 3)

    # CUDA-core FMA peaks on this GPU, this run (register-only FMA loops) + operand-pattern probes
    peak = {"f64": ops.fma_peak(torch.float64, 2048), "f32": ops.fma_peak(torch.float32, 2048)}
    dtype = torch.float64 if tag == "f64" else torch.float32
    fma_peak = peak[tag]
    fma_probe = {"loop_invariant_operands": fma_peak / 1e12,
                 "three_distinct_vector_operands": ops.fma_peak(dtype, 2048, 1) / 1e12,
                 "one_operand_shared_by_4": ops.fma_peak(dtype, 2048, 2) / 1e12,
                 # the filter-gradient kernel's own register pattern (49 accumulators, g shared by 7
                 # FMAs, window rows from shared memory): the ceiling of kernel B2
                 "filter_gradient_pattern": ops.fma_peak(dtype, 512, 3) / 1e12}

    m = resident_and_e2e(P, dev, args, tag, B, blocks, args.steps, warmup, world, rank, local, True)
    ms, ms_e2e, kind_ms, kind_n, clocks = m["ms"], m["ms_e2e"], m["kind_ms"], m["kind_n"], m["clocks"]
    keep, chunk, pass_chunk, H, W = m["keep"], m["chunk"], m["pass_chunk"], m["H"], m["W"]
    kflops, kbytes, knames = dict(KERNEL_FLOPS), dict(KERNEL_BYTES), dict(KERNEL_NAMES)
    if keep:
        kflops.update(KEPT_FLOPS)
        kbytes.update(KEPT_BYTES)
        knames.update(KEPT_NAMES)

    slab = None
    if world > 1 and not args.no_slab:
        slab = slab_record(P, dev, rank, world, args)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pts = B * H * W * blocks  # grid-point-steps per pass, whole job
    value = pts * args.steps / (ms * 1e-3)
    e2e = pts * args.steps / (ms_e2e * 1e-3)
    hbm_peak, hbm_src = peaks()
    nominal = NOMINAL_TFLOPS[tag]
    # dominant kernel = the kind with the largest share of the timed region
    top = max(kflops, key=lambda kk: kind_ms[kk])
    n_top = max(1, kind_n[top])
    ms_top = kind_ms[top] / n_top
    es = 8 if tag == "f64" else 4

    def launch_pts(kk):  # grid points one launch of this kernel covers
        return (pass_chunk if kk == "fwd_step" else min(chunk, pass_chunk)) * H * W

    pts_per_launch = launch_pts(top)
    achieved = kflops[top] * pts_per_launch / (ms_top * 1e-3) / 1e12
    per_kernel = {}
    for kk in kflops:
        if kind_ms[kk] <= 0:
            continue
        sec = kind_ms[kk] / max(1, kind_n[kk]) * 1e-3
        tf = kflops[kk] * launch_pts(kk) / sec / 1e12
        per_kernel[kk] = {"ms_per_launch": sec * 1e3, "launches": kind_n[kk], "tflops": tf,
                          "frac_of_fma_peak": tf * 1e12 / fma_peak, "frac_of_nominal_peak": tf / nominal,
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
    whole_tf = F_TOTAL * value / 1e12 / world
    line = {
        "metric": "grid-point-steps/sec fwd+bwd rollout", "value": value, "unit": "grid-point-steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": warmup, "warmup_requested": args.warmup,
        "ms_per_step": ms / args.steps,
        "ms_per_step_best": m["per_step"][0], "ms_per_step_median": statistics.median(m["per_step"]),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": tag, "data": "synthetic",
        "config": {"workload": CFG["workload"] if tag == "f64" else CFG["workload"].replace("fp64", "fp32"),
                   "global_batch": B, "per_gpu_batch": m["Bl"], "grid": [H, W],
                   "dt_blocks": blocks, "max_order": CFG["max_order"], "dt": m["dt"], "batch_chunk": chunk,
                   "model": "viscous Burgers SymNet (nu = 0.05) + 1e-2 N(0,1) on every SymNet entry, 1e-3 N(0,1) on "
                            "every free moment; dt = dx^2 (stable forward Euler)",
                   "pass_chunk": pass_chunk, "features": "kept in HBM for the reverse pass" if keep else "recomputed",
                   "parallelism": "dp%d (batch-sharded, one in-place NCCL all-reduce of the flat %d-scalar gradient "
                                  "buffer incl. the loss)" % (world, m["nparam"] + 1),
                   "l2": "inputs (%.1f GB per pass) larger than L2; no flush needed" % (m["h2d"] / 1e9)},
        "e2e": {"value": e2e, "unit": "grid-point-steps/s", "h2d_bytes_per_step": m["h2d"],
                "d2h_bytes_per_step": m["d2h"], "ms_per_step": ms_e2e / args.steps,
                "how": "pinned host batch streamed in chunks of up to %d trajectories (first chunks smaller: %s); H2D of "
                       "chunk k+1 overlaps fwd+bwd of chunk k, H2D of the next pass's first chunk overlaps this pass's last chunk "
                       "(every pass copies every chunk once); loss+gradients copied back to pinned host memory in one copy"
                       % (m["e2e_chunk"], m["ramp"])},
        "gpu_launches": int(sum(kind_n.values())),
        "clocks": clocks,
        "roofline": {
            "bound": "fp64" if tag == "f64" else "fp32",
            "kernel": knames[top].replace("double", "double" if tag == "f64" else "float"),
            "achieved": achieved, "peak": fma_peak / 1e12, "unit": "TFLOP/s", "frac": achieved / (fma_peak / 1e12),
            "nominal_peak": nominal, "frac_of_nominal_peak": achieved / nominal,
            "peak_source": "measured this run: pdenet_fma_peak register-only FMA loop; nominal = 148 SM x %d FMA/clk x 2 x "
                           "1.965 GHz" % (64 if tag == "f64" else 128),
            "flops_per_launch": kflops[top] * pts_per_launch, "ms_per_launch": ms_top, "launches": n_top,
            "traffic": traffic, "per_kernel": per_kernel, "fma_probe_tflops": fma_probe,
            "kernel_ms_share": {kk: (v / total_kernel_ms if total_kernel_ms else None) for kk, v in kind_ms.items()},
            "whole_pass": {"flop_per_point_step": F_TOTAL, "achieved_tflops": whole_tf,
                           "frac_of_fma_peak": whole_tf * 1e12 / fma_peak, "frac_of_nominal_peak": whole_tf / nominal,
                           # the peak probe runs for milliseconds at the full boost clock; a long run can
                           # sit at the board's power limit (clocks.reasons): same fraction against the
                           # peak scaled to the median SM clock sampled during the timed region
                           "frac_of_fma_peak_at_sampled_clock": (
                               whole_tf * 1e12 / (fma_peak * clocks["sm_mhz"] / clocks["sm_max_mhz"])
                               if clocks and clocks.get("sm_mhz") and clocks.get("sm_max_mhz") else None),
                           "hbm_gbs": BYTES_PER_POINT_STEP * es / 8 * value / 1e9 / world,
                           "frac_of_hbm_peak": BYTES_PER_POINT_STEP * es / 8 * value / 1e9 / world / hbm_peak,
                           "hbm_peak_source": hbm_src},
        },
    }
    if slab is not None:
        line["cfg5_slab"] = slab
    if world == 1 and not args.no_extras:
        # ---- fp32 sub-record: the same workload in single precision, a shorter run (64 trajectories, 4 blocks)
        other = "f32" if tag == "f64" else "f64"
        mo = resident_and_e2e(P, dev, args, other, min(B, 64), min(blocks, 4), 2, 3, 1, 0, local, True)
        pts_o = min(B, 64) * H * W * min(blocks, 4)
        v_o = pts_o * 2 / (mo["ms"] * 1e-3)
        line["fp32" if other == "f32" else "fp64"] = {
            "value": v_o, "unit": "grid-point-steps/s", "e2e": pts_o * 2 / (mo["ms_e2e"] * 1e-3),
            "global_batch": min(B, 64), "dt_blocks": min(blocks, 4), "steps": 2,
            "fma_peak_tflops": peak[other] / 1e12, "nominal_peak_tflops": NOMINAL_TFLOPS[other],
            "whole_pass_tflops": F_TOTAL * v_o / 1e12, "frac_of_fma_peak": F_TOTAL * v_o / peak[other],
            "frac_of_nominal_peak": F_TOTAL * v_o / 1e12 / NOMINAL_TFLOPS[other], "clocks": mo["clocks"],
            "kernel_ms_per_launch": {kk: mo["kind_ms"][kk] / max(1, mo["kind_n"][kk]) for kk in KERNEL_FLOPS}}
        line["small_configs"] = small_configs_record(P, dev, {kk: v / 1e12 for kk, v in peak.items()})
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"], ref = cpu_baseline(os.cpu_count() or 1, keep=not args.no_extras)
        if ref is not None:
            line["parity"] = parity_record(P, dev, ref, chunk)
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
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg (and the parity record)")
    ap.add_argument("--no-extras", action="store_true", help="skip the fp32 / small_configs / parity sub-records")
    ap.add_argument("--no-slab", action="store_true", help="N > 1: skip the cfg5 slab-decomposition sub-record")
    ap.add_argument("--slab-size", type=int, default=8192, help="grid side of the cfg5 sub-record")
    ap.add_argument("--slab-blocks", type=int, default=4, help="dt-blocks of the cfg5 sub-record")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"], help="f64 = the reference's precision (default)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
