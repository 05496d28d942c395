"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: batch sharding + flat gradient
all-reduce (DIST-2) and slab decomposition + ring halo exchange (DIST-1).  The per-rank compute is
the CPU oracle here (the CUDA kernels need a GPU); what is under test is the exchange logic."""
import math
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pdenet_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model(k=5, max_order=2, C=2, L=2, dx=0.3):
    g = torch.Generator().manual_seed(3)
    orders = O.order_list(max_order)
    M = O.initial_moments(k, orders) + 1e-2 * torch.randn(len(orders), k, k, generator=g, dtype=torch.float64)
    Ks = O.scaled_kernels(M, O.filter_scales(orders, dx))
    sym = 0.1 * torch.randn(C, O.symnet_param_count(C * len(orders), L), generator=g, dtype=torch.float64)
    u0 = torch.randn(4, C, 16, 12, generator=g, dtype=torch.float64)
    return Ks, sym, u0


def _weights(u0):
    return torch.randn(u0.shape, generator=torch.Generator().manual_seed(17), dtype=torch.float64)


def _oracle_slab_step(ext, Ks, sym, L, dt, halo, out, out_row_offset):
    """Non-periodic-in-H step on an extended slab, built from the oracle's periodic step: pad H
    periodically is wrong for a slab, so evaluate on the extended block and keep interior rows
    (rows within `halo` of the block edge are contaminated by the wrap and discarded)."""
    Hs = ext.shape[2] - 2 * halo
    full = O.step(ext, Ks, sym, L, dt)
    out[:, :, out_row_offset:out_row_offset + Hs] = full[:, :, halo:halo + Hs]
    return out


class _OracleSlabBackward:
    """CPU stand-in for ops.SlabBackward built from the oracle's autograd (same interface)."""

    def __init__(self, Ks, sym, L, dt, B, C, Hs, W):
        self.Ks, self.sym, self.L, self.dt, self.Hs = Ks, sym, L, dt, Hs
        self.r = (Ks.shape[-1] - 1) // 2
        self.gk, self.gs = torch.zeros_like(Ks), torch.zeros_like(sym)

    def step(self, u_ext2, lam_ext, lam_out, out_row_offset=0):
        r, Hs = self.r, self.Hs
        u = u_ext2.clone().requires_grad_(True)
        K, s = self.Ks.clone().requires_grad_(True), self.sym.clone().requires_grad_(True)
        new = O.step(u, K, s, self.L, self.dt)           # rows within r of the block edge are wrap-contaminated
        val = new[:, :, r:r + Hs + 2 * r]                 # logical rows [-r, Hs + r): clean
        # data gradient: every row that reaches the owned rows contributes
        gu, = torch.autograd.grad((val * lam_ext).sum(), u, retain_graph=True)
        lam_out[:, :, out_row_offset:out_row_offset + Hs] = gu[:, :, 2 * r:2 * r + Hs]
        # parameter gradients: owned rows only
        own = (val[:, :, r:r + Hs] * lam_ext[:, :, r:r + Hs]).sum()
        gk, gs = torch.autograd.grad(own, (K, s))
        self.gk += gk
        self.gs += gs
        return lam_out

    def finish(self):
        return self.gk, self.gs


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import pdenet_b200  # noqa: F401
        from pdenet_b200 import parallel as PP
        torch.set_num_threads(1)
        Ks, sym, u0 = _model()
        L, dt, steps = 2, 1e-2, 3
        # ---- DIST-2: batch-sharded gradients, one flat all-reduce
        b0, nb = PP.shard_range(u0.shape[0], rank, world)
        Kc, sc = Ks.clone().requires_grad_(True), sym.clone().requires_grad_(True)
        out = O.rollout(u0[b0:b0 + nb], Kc, sc, L, dt, steps)[-1]
        loss = out.square().sum() / u0[0].numel() / u0.shape[0]
        loss.backward()
        total = PP.flat_allreduce([Kc.grad, sc.grad], extra=loss.detach())
        # ---- 8f-4: gradients that already live in one flat buffer are reduced in place (no packing)
        lin = torch.nn.Linear(3, 2).double()
        fp = pdenet_b200.FlatParams(lin)
        fp.gflat.copy_(torch.arange(fp.numel, dtype=torch.float64) * (rank + 1))
        ptr = fp.gflat.data_ptr()
        red = PP.allreduce_gradients(fp, loss=torch.tensor(float(rank + 1), dtype=torch.float64))
        tri = world * (world + 1) / 2
        assert fp.gflat.data_ptr() == ptr and float(red) == tri
        assert torch.equal(lin.weight.grad.reshape(-1), torch.arange(6, dtype=torch.float64) * tri)
        assert torch.equal(lin.bias.grad, torch.arange(6, 8, dtype=torch.float64) * tri)
        # ---- DIST-1: slab decomposition with ring halo exchange
        slab = PP.SlabDecomposition(u0.shape[2], 2, rank, world)
        mine = u0[:, :, slab.row0:slab.row0 + slab.Hs].contiguous()
        fin = slab.rollout(mine, Ks, sym, L, dt, steps, step_fn=_oracle_slab_step)
        # ---- slab reverse pass: 2r / r ghost-row exchanges per step + flat all-reduce of the grads
        wloc = _weights(u0)[:, :, slab.row0:slab.row0 + slab.Hs].contiguous()
        fin2, gu0, sgk, sgs = slab.rollout_fwd_bwd(mine, Ks, sym, L, dt, steps, lambda f: wloc,
                                                   step_fn=_oracle_slab_step, bwd_factory=_OracleSlabBackward)
        # by value (numpy), not as shared-memory tensors: the worker may exit before the parent reads
        q.put(tuple(t.detach().numpy() if torch.is_tensor(t) else t
                    for t in (rank, Kc.grad, sc.grad, float(total), fin, fin2, gu0, sgk, sgs)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_dp_and_slab_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([tuple(torch.from_numpy(t) if hasattr(t, "dtype") and hasattr(t, "shape") and getattr(t, "ndim", 0) > 0
                        else t for t in q.get(timeout=240)) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    Ks, sym, u0 = _model()
    L, dt, steps = 2, 1e-2, 3
    Kc, sc = Ks.clone().requires_grad_(True), sym.clone().requires_grad_(True)
    ref = O.rollout(u0, Kc, sc, L, dt, steps)[-1]
    loss = ref.square().sum() / u0[0].numel() / u0.shape[0]
    loss.backward()
    for rank, gk, gs, tot, fin, fin2, gu0, sgk, sgs in res:
        assert torch.allclose(gk, Kc.grad, rtol=1e-11, atol=1e-14)
        assert torch.allclose(gs, sc.grad, rtol=1e-11, atol=1e-14)
        assert abs(tot - float(loss)) < 1e-12 * max(1.0, abs(float(loss)))
    full = torch.cat([r[4] for r in res], dim=2)
    assert torch.allclose(full, ref, rtol=1e-12, atol=1e-13)
    # slab reverse: stitched input gradient and all-reduced parameter gradients == single domain
    u0r, K2, s2 = u0.clone().requires_grad_(True), Ks.clone().requires_grad_(True), sym.clone().requires_grad_(True)
    (O.rollout(u0r, K2, s2, L, dt, steps)[-1] * _weights(u0)).sum().backward()
    assert torch.allclose(torch.cat([r[5] for r in res], dim=2), ref, rtol=1e-12, atol=1e-13)
    assert torch.allclose(torch.cat([r[6] for r in res], dim=2), u0r.grad, rtol=1e-10, atol=1e-12)
    for r_ in res:
        assert torch.allclose(r_[7], K2.grad, rtol=1e-10, atol=1e-12)
        assert torch.allclose(r_[8], s2.grad, rtol=1e-10, atol=1e-12)


def test_shard_range_and_single_rank_slab():
    import pdenet_b200  # noqa: F401
    from pdenet_b200 import parallel as PP
    for total, world in [(256, 8), (10, 4), (3, 4)]:
        got = [PP.shard_range(total, r, world) for r in range(world)]
        assert sum(n for _, n in got) == total
        assert all(got[i][0] + got[i][1] == got[i + 1][0] for i in range(world - 1))
    # world == 1: ghosts wrap onto the rank's own rows -> identical to the periodic rollout
    Ks, sym, u0 = _model()
    slab = PP.SlabDecomposition(u0.shape[2], 2, 0, 1)
    fin = slab.rollout(u0, Ks, sym, 2, 1e-2, 2, step_fn=_oracle_slab_step)
    assert torch.allclose(fin, O.rollout(u0, Ks, sym, 2, 1e-2, 2)[-1], rtol=1e-12, atol=1e-13)


def test_ramp_chunks_cover_the_batch_exactly():
    """Chunk schedule of the streamed e2e pass: short first chunk, the second completes the first
    full chunk, every size within bounds, sizes add up to the batch (host logic, no GPU)."""
    import importlib.util
    import re
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pde-net_b200",
                            "parallel.py")).read()
    ns = {}
    exec(re.search(r"def ramp_chunks.*?return sizes\n", src, re.S).group(0), ns)
    ramp = ns["ramp_chunks"]
    assert ramp(256, 32) == [8, 24] + [32] * 7
    assert ramp(32, 16) == [4, 12, 16]
    assert ramp(7, 3, 3) == [3, 3, 1]
    for B in range(1, 80):
        for chunk in range(1, 40):
            sizes = ramp(B, chunk)
            assert sum(sizes) == B and all(1 <= s <= max(chunk, 1) for s in sizes)
            assert sizes[0] <= max(1, chunk // 4) or B < chunk // 4 or chunk < 4
