"""Multi-GPU plumbing for the hot path: one process per GPU, torch.distributed.

* Batch sharding (config 4): trajectories are independent, so each rank runs the fused rollout on
  its shard and the only exchange is ONE all-reduce of a single flat buffer holding every
  moment / SymNet gradient (plus the loss) -- a few KB, latency-bound over NVLink/NVSwitch.
* Slab domain decomposition (config 5): rows of one huge grid are split across ranks; before each
  step every rank fills its ghost rows from its ring neighbours (periodic in H makes it a true
  ring) and calls the non-periodic slab step of the C-ABI.  The reverse pass exchanges 2r ghost
  rows of the stored state and r ghost rows of the adjoint per step, recomputes on the ghost
  rows, and ends with the same flat all-reduce of the parameter gradients.

The reference is single-process: nothing here replaces a reference interface.
"""
from __future__ import annotations

from typing import Callable, Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard ``(start, count)`` of ``total`` items for ``rank``; sizes differ by <= 1."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def flat_allreduce(tensors: Sequence[torch.Tensor], group=None, extra: Optional[torch.Tensor] = None):
    """Sum ``tensors`` (and the optional scalar ``extra``) across ranks through ONE flat buffer
    and write the result back in place.  Returns the reduced ``extra`` (or None)."""
    tensors = [t for t in tensors if t is not None]
    n = sum(t.numel() for t in tensors)
    ref = tensors[0] if tensors else extra
    flat = torch.empty(n + (1 if extra is not None else 0), dtype=ref.dtype, device=ref.device)
    o = 0
    for t in tensors:
        flat[o:o + t.numel()].copy_(t.reshape(-1))
        o += t.numel()
    if extra is not None:
        flat[n] = extra
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, group=group)
    o = 0
    for t in tensors:
        t.copy_(flat[o:o + t.numel()].reshape(t.shape))
        o += t.numel()
    return flat[n] if extra is not None else None


def allreduce_gradients(params, group=None, loss: Optional[torch.Tensor] = None):
    """Data-parallel gradient exchange: one all-reduce for every ``.grad`` (and the loss).
    ``params``: an ``optim.FlatParams`` (the gradients already live in one flat buffer: reduced in
    place, nothing packed or unpacked) or any iterable of parameters (packed into a flat buffer)."""
    if hasattr(params, "allreduce"):
        return params.allreduce(loss=loss, group=group)
    grads = [p.grad for p in params if p.grad is not None]
    return flat_allreduce(grads, group=group, extra=loss)


class SlabDecomposition:
    """Rows ``[row0, row0+Hs)`` of an H-row periodic grid live on this rank; ``halo`` ghost rows on
    each side are filled from the ring neighbours before every step."""

    def __init__(self, H: int, halo: int, rank: int, world: int, group=None):
        if H % world:
            raise ValueError("H must divide evenly over the ranks")
        self.H, self.halo, self.rank, self.world, self.group = H, halo, rank, world, group
        self.Hs = H // world
        if self.Hs < halo:
            raise ValueError("slab thinner than the halo")
        self.row0 = rank * self.Hs
        self.up = (rank - 1) % world      # owns the rows above (smaller y)
        self.down = (rank + 1) % world    # owns the rows below

    def extend(self, u_local: torch.Tensor, halo: Optional[int] = None) -> torch.Tensor:
        """[B,C,Hs,W] -> [B,C,Hs+2*halo,W] with the interior copied in (ghosts unset)."""
        h = self.halo if halo is None else halo
        B, C, Hs, W = u_local.shape
        ext = torch.empty(B, C, Hs + 2 * h, W, dtype=u_local.dtype, device=u_local.device)
        ext[:, :, h:h + Hs] = u_local
        return ext

    def exchange(self, ext: torch.Tensor, halo: Optional[int] = None) -> None:
        """Fill the ghost rows of ``ext`` in place: top ghosts <- bottom rows of `up`, bottom
        ghosts <- top rows of `down`.  One batched isend/irecv per neighbour."""
        h, Hs = (self.halo if halo is None else halo), self.Hs
        if h > Hs:
            raise ValueError("ghost zone wider than a slab: needs data beyond the ring neighbours")
        top_rows = ext[:, :, h:2 * h].contiguous()            # my first rows -> `up`'s bottom ghosts
        bot_rows = ext[:, :, Hs:Hs + h].contiguous()          # my last rows  -> `down`'s top ghosts
        if self.world == 1:
            ext[:, :, :h] = bot_rows
            ext[:, :, h + Hs:] = top_rows
            return
        recv_top = torch.empty_like(top_rows)
        recv_bot = torch.empty_like(bot_rows)
        # NCCL matches point-to-point operations between a pair of ranks by ORDER, not by tag: with
        # two ranks `up` and `down` are the same peer, so the receives are posted in the order the
        # peer sends (its first send, its top rows, is my bottom ghost zone).
        ops = [dist.P2POp(dist.isend, top_rows, self.up, self.group, tag=0),
               dist.P2POp(dist.isend, bot_rows, self.down, self.group, tag=1),
               dist.P2POp(dist.irecv, recv_bot, self.down, self.group, tag=0),
               dist.P2POp(dist.irecv, recv_top, self.up, self.group, tag=1)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        ext[:, :, :h] = recv_top
        ext[:, :, h + Hs:] = recv_bot

    def rollout(self, u_local: torch.Tensor, kernels: torch.Tensor, symnet: torch.Tensor, L: int, dt: float,
                steps: int, step_fn: Optional[Callable] = None) -> torch.Tensor:
        """``steps`` forward-Euler steps on this rank's slab.  ``step_fn(ext, kernels, symnet, L,
        dt, halo, out, out_row_offset)`` defaults to the CUDA slab step; it must write the new
        interior rows into ``out`` starting at ``out_row_offset``."""
        if step_fn is None:
            from . import ops
            step_fn = ops.slab_step
        h = self.halo
        a = self.extend(u_local)
        b = torch.empty_like(a)
        for _ in range(steps):
            self.exchange(a)
            step_fn(a, kernels, symnet, L, dt, h, b, h)   # new interior lands inside the other buffer
            a, b = b, a
        return a[:, :, h:h + self.Hs].contiguous()

    def rollout_fwd_bwd(self, u_local: torch.Tensor, kernels: torch.Tensor, symnet: torch.Tensor, L: int, dt: float,
                        steps: int, grad_last_local: Callable, step_fn: Optional[Callable] = None,
                        bwd_factory: Optional[Callable] = None, keep_features: Optional[bool] = None):
        """Forward rollout on this rank's slab, then the reverse pass.

        ``grad_last_local(final_local) -> dLoss/d(final state)`` for the owned rows (the loss is
        a sum over ranks of local terms).  Returns ``(final_local, grad_u0_local, grad_kernels,
        grad_symnet)`` with the parameter gradients already summed over the ranks (one flat
        all-reduce).  ``bwd_factory(kernels, symnet, L, dt, B, C, Hs, W)`` defaults to the CUDA
        ``ops.SlabBackward``.

        No full-slab copies: every forward state lives in its own buffer with 2r ghost rows (the
        forward exchanges 2r rows, so the reverse pass needs no state exchange at all), each step
        writes straight into the interior of the next buffer, and the reverse step writes the new
        adjoint into the interior of the next r-ghosted adjoint buffer.

        ``keep_features`` (CUDA steps only): the forward steps keep the derivative features of the
        owned rows and the reverse steps read them instead of recomputing the stencils (ghost rows
        are always recomputed); None = keep when they fit ``ops.FEATURE_CACHE_BYTES``."""
        from . import ops
        cuda_steps = step_fn is None and bwd_factory is None
        if step_fn is None:
            step_fn = ops.slab_step
        if bwd_factory is None:
            bwd_factory = ops.SlabBackward
        h = self.halo
        h2 = 2 * h
        B, C, Hs, W = u_local.shape
        feats = _slab_feature_buffers(keep_features if cuda_steps else False, steps, B, C, kernels.shape[0], Hs, W,
                                      u_local)
        bufs = [self.extend(u_local, h2)]            # u_0 .. u_steps, each with 2r ghost rows
        for t in range(steps):
            self.exchange(bufs[-1], h2)
            nxt = torch.empty_like(bufs[-1])
            if feats is not None:
                step_fn(bufs[-1], kernels, symnet, L, dt, h2, nxt, h2, feats=feats[t])
            else:
                step_fn(bufs[-1], kernels, symnet, L, dt, h2, nxt, h2)
            bufs.append(nxt)
        final = bufs[-1][:, :, h2:h2 + Hs].contiguous()
        lam_ext = self.extend(grad_last_local(final).contiguous(), h)
        lam_nxt = torch.empty_like(lam_ext)
        bwd = bwd_factory(kernels, symnet, L, dt, B, C, Hs, W)
        for t in range(steps - 1, -1, -1):
            self.exchange(lam_ext, h)
            if feats is not None:
                bwd.step(bufs[t], lam_ext, lam_nxt, h, feats=feats[t])
            else:
                bwd.step(bufs[t], lam_ext, lam_nxt, h)   # bufs[t] already carries its 2r ghost rows
            lam_ext, lam_nxt = lam_nxt, lam_ext
        gk, gs = bwd.finish()
        flat_allreduce([gk, gs], group=self.group)
        return final, lam_ext[:, :, h:h + Hs].contiguous(), gk, gs


def _slab_feature_buffers(keep, steps, B, C, N, Hs, W, like: torch.Tensor):
    """[steps, B, C*N, Hs, W] buffer for the kept features of a slab rollout, or None."""
    from . import ops
    nbytes = steps * B * C * N * Hs * W * like.element_size()
    if keep is None:
        keep = nbytes <= ops.FEATURE_CACHE_BYTES
    if not keep:
        return None
    return torch.empty(steps, B, C * N, Hs, W, dtype=like.dtype, device=like.device)


class SlabPeerBuffers:
    """State / adjoint buffers of a slab rollout in SYMMETRIC memory (every rank allocates the same
    layout and maps its ring neighbours' copies over NVLink), so the step kernels can store their
    boundary rows straight into the neighbours' ghost rows: the halo exchange is fused into the
    compute kernel instead of being an NCCL send/recv between steps.  One ``barrier()`` (a
    symmetric-memory signal exchange on the stream) separates consecutive steps."""

    def __init__(self, slab: "SlabDecomposition", B: int, C: int, W: int, steps: int, dtype, device):
        import torch.distributed._symmetric_memory as symm_mem
        self.slab, self.steps = slab, steps
        h, h2, Hs = slab.halo, 2 * slab.halo, slab.Hs
        self.n_state = B * C * (Hs + 2 * h2) * W
        self.n_lam = B * C * (Hs + 2 * h) * W
        total = (steps + 1) * self.n_state + 2 * self.n_lam
        self.flat = symm_mem.empty(total, dtype=dtype, device=device)
        if slab.world > 1:
            group = slab.group if slab.group is not None else dist.group.WORLD
            self.hdl = symm_mem.rendezvous(self.flat, group.group_name)
            ptrs = self.hdl.buffer_ptrs
        else:
            self.hdl, ptrs = None, [self.flat.data_ptr()]
        es = self.flat.element_size()
        self.base_up, self.base_dn = ptrs[slab.up], ptrs[slab.down]
        self.state = [self.flat[i * self.n_state:(i + 1) * self.n_state].view(B, C, Hs + 2 * h2, W)
                      for i in range(steps + 1)]
        o = (steps + 1) * self.n_state
        self.lam = [self.flat[o + i * self.n_lam:o + (i + 1) * self.n_lam].view(B, C, Hs + 2 * h, W) for i in range(2)]
        self._es, self._lam_off = es, o

    def state_peers(self, i):
        off = i * self.n_state * self._es
        return self.base_up + off, self.base_dn + off

    def lam_peers(self, i):
        off = (self._lam_off + i * self.n_lam) * self._es
        return self.base_up + off, self.base_dn + off

    def barrier(self):
        if self.hdl is not None:
            self.hdl.barrier()

    def input_view(self) -> torch.Tensor:
        """The owned rows of the initial-state buffer, [B,C,Hs,W] (strided): an initial condition written here
        is used in place by `slab_rollout_fwd_bwd_p2p` (no slab-sized copy at the start of a pass)."""
        h2 = 2 * self.slab.halo
        return self.state[0][:, :, h2:h2 + self.slab.Hs]

    def _peer_view(self, peer: int, elem_off: int, shape):
        return self.hdl.get_buffer(peer, tuple(shape), self.flat.dtype, elem_off)

    def fill_ghosts(self, which: str, i: int, halo: int) -> None:
        """Ghost fill of ``state[i]`` / ``lam[i]`` without NCCL: store my first / last ``halo`` owned rows straight
        into the ring neighbours' ghost rows of the same buffer (peer copies over NVLink).  The caller brackets it
        with ``barrier()``: one before (the neighbours are done with the buffer), one after (their rows have
        landed in mine)."""
        slab, Hs = self.slab, self.slab.Hs
        ext = (self.state if which == "state" else self.lam)[i]
        base = i * self.n_state if which == "state" else self._lam_off + i * self.n_lam
        top_rows, bot_rows = ext[:, :, halo:2 * halo], ext[:, :, Hs:Hs + halo]
        if self.hdl is None:
            ext[:, :, :halo] = bot_rows
            ext[:, :, halo + Hs:] = top_rows
            return
        up = self._peer_view(slab.up, base, ext.shape)
        dn = self._peer_view(slab.down, base, ext.shape)
        up[:, :, halo + Hs:].copy_(top_rows)      # my first rows -> `up`'s bottom ghosts
        dn[:, :, :halo].copy_(bot_rows)           # my last rows  -> `down`'s top ghosts


def slab_rollout_fwd_bwd_p2p(slab: "SlabDecomposition", bufs: SlabPeerBuffers, u_local: torch.Tensor,
                             kernels: torch.Tensor, symnet: torch.Tensor, L: int, dt: float, steps: int,
                             grad_last_local: Callable, keep_features: Optional[bool] = None,
                             copy_outputs: bool = True):
    """``SlabDecomposition.rollout_fwd_bwd`` with the halo exchanges fused into the step kernels
    (peer stores into the neighbours' ghost rows, `SlabPeerBuffers`).  The initial state and the loss
    seed fill the neighbours' ghost rows with peer copies (`SlabPeerBuffers.fill_ghosts`): no NCCL
    point-to-point on the path.  ``grad_last_local`` may accept ``out=`` (a strided view of the adjoint
    buffer) to write the seed in place.  ``u_local`` may be ``bufs.input_view()`` itself (used in place).
    ``copy_outputs=False`` returns the final state and the input gradient as strided VIEWS of the
    persistent buffers (valid until the next pass) instead of contiguous copies."""
    from . import ops
    h, h2, Hs = slab.halo, 2 * slab.halo, slab.Hs
    B, C, _, W = u_local.shape
    feats = _slab_feature_buffers(keep_features, steps, B, C, kernels.shape[0], Hs, W, u_local)
    own = bufs.input_view()
    if not (u_local.data_ptr() == own.data_ptr() and u_local.stride() == own.stride()):
        own.copy_(u_local)
    bufs.barrier()                                   # every rank is done with the buffers of its previous pass
    bufs.fill_ghosts("state", 0, h2)
    bufs.barrier()                                   # the neighbours' rows have landed
    for t in range(steps):
        up, dn = bufs.state_peers(t + 1)
        ops.slab_step(bufs.state[t], kernels, symnet, L, dt, h2, bufs.state[t + 1], h2, up, dn, h2,
                      feats=None if feats is None else feats[t])
        bufs.barrier()                               # neighbours' boundary rows have landed
    final = bufs.state[steps][:, :, h2:h2 + Hs]
    if copy_outputs:
        final = final.contiguous()
    cur = 0
    seed_view = bufs.lam[cur][:, :, h:h + Hs]
    try:
        seed = grad_last_local(final, out=seed_view)
    except TypeError:                                # a plain callable: one more slab-sized copy
        seed = grad_last_local(final)
    if seed is not seed_view:
        seed_view.copy_(seed)
    bufs.fill_ghosts("lam", cur, h)                  # lam[cur] was last written two steps before the previous
    bufs.barrier()                                   # pass ended: every rank is past reading it
    bwd = ops.SlabBackward(kernels, symnet, L, dt, B, C, Hs, W)
    for t in range(steps - 1, -1, -1):
        up, dn = bufs.lam_peers(cur ^ 1)
        bwd.step(bufs.state[t], bufs.lam[cur], bufs.lam[cur ^ 1], h, up, dn, h,
                 feats=None if feats is None else feats[t])
        bufs.barrier()
        cur ^= 1
    gk, gs = bwd.finish()
    flat_allreduce([gk, gs], group=slab.group)
    grad_u0 = bufs.lam[cur][:, :, h:h + Hs]
    return final, (grad_u0.contiguous() if copy_outputs else grad_u0), gk, gs


def chunked_fwd_bwd(model, u0: torch.Tensor, T: float, loss_scale: float, chunk: int) -> torch.Tensor:
    """Forward + reverse rollout over a DEVICE-resident batch, ``chunk`` trajectories at a time
    (trajectories are independent; parameter gradients accumulate in ``.grad``).  Bounding the
    live set to one chunk is what lets the rollout keep its derivative features for the reverse
    pass (``ops.rollout(keep_features=...)``: 96 B per grid point and step at cfg4) instead of
    recomputing the stencils.  Loss as in :func:`streamed_fwd_bwd`."""
    total = torch.zeros((), dtype=u0.dtype, device=u0.device)
    for b0 in range(0, u0.shape[0], chunk):
        out = model(u0[b0:b0 + chunk], T)
        loss = out.square().sum() * loss_scale
        loss.backward()
        total = total + loss.detach()
    return total


def ramp_chunks(B: int, chunk: int, first: int = 0):
    """Chunk sizes for a streamed pass: a short first chunk (``first``, default chunk/4) so that the
    one host-to-device copy nothing can overlap is short, a second chunk that completes the first
    ``chunk`` trajectories, then full chunks."""
    first = min(B, max(1, first or chunk // 4))
    sizes, left = [first], B - first
    if left > 0 and first < chunk:
        n = min(chunk - first, left)
        sizes.append(n)
        left -= n
    while left > 0:
        n = min(chunk, left)
        sizes.append(n)
        left -= n
    return sizes


class StreamedBatch:
    """A batch that lives in PINNED HOST memory, streamed through the device pass after pass.

    Trajectories are independent, so a pass goes through the device in chunks: the host-to-device copy of
    chunk k+1 (on a side stream) overlaps the fused forward + reverse rollout of chunk k, and the parameter
    gradients accumulate in ``.grad``.  With ``ramp`` the first chunks are smaller (chunk/4, then up to
    ``chunk``): only the very first copy cannot be overlapped.  With ``prefetch_next`` the first chunk of the
    NEXT pass is copied while the last chunk of this one computes (the inputs of an optimiser's next
    evaluation are known in advance: same batch, new parameters), so in steady state no copy is exposed at
    all -- at 8 GPUs, where a pass is 8x shorter, the exposed first copy was ~6 % of the pass.  Every pass
    still copies every chunk exactly once.  ``run`` returns the summed loss
    ``loss_scale * sum(final_state**2)`` (0-dim device tensor)."""

    def __init__(self, host_u0: torch.Tensor, chunk: int, device, bufs=None, copy_stream=None, ramp: bool = True,
                 prefetch_next: bool = False):
        self.host, self.chunk, self.dev, self.prefetch_next = host_u0, chunk, device, prefetch_next
        B = host_u0.shape[0]
        self.copy_stream = copy_stream or torch.cuda.Stream(device)
        self.bufs = bufs or [torch.empty((min(chunk, B),) + tuple(host_u0.shape[1:]), dtype=host_u0.dtype,
                                         device=device) for _ in range(2)]
        self.copied = [torch.cuda.Event() for _ in range(2)]
        self.freed = [torch.cuda.Event() for _ in range(2)]
        self.sizes = ramp_chunks(B, chunk) if ramp else ramp_chunks(B, chunk, chunk)
        self.starts = [sum(self.sizes[:i]) for i in range(len(self.sizes))]
        self.g = 0                 # chunks consumed so far, over all passes: chunk g lives in bufs[g % 2]
        self.issued = 0            # chunks whose copy has been issued so far

    def _issue_copy(self, k: int) -> None:
        """Copy chunk k of the batch into the buffer of global chunk number ``self.issued``."""
        j = self.issued
        b0, nb = self.starts[k], self.sizes[k]
        with torch.cuda.stream(self.copy_stream):
            if j >= 2:
                self.copy_stream.wait_event(self.freed[j % 2])     # global chunk j-2 has finished with this buffer
            self.bufs[j % 2][:nb].copy_(self.host[b0:b0 + nb], non_blocking=True)
            self.copied[j % 2].record(self.copy_stream)
        self.issued += 1

    def run(self, model, T: float, loss_scale: float) -> torch.Tensor:
        main = torch.cuda.current_stream(self.dev)
        n = len(self.starts)
        total = torch.zeros((), dtype=self.host.dtype, device=self.dev)
        if self.issued == self.g:                     # nothing prefetched by the previous pass
            self.copy_stream.wait_stream(main)
            self._issue_copy(0)
        for k in range(n):
            if k + 1 < n:
                self._issue_copy(k + 1)
            elif self.prefetch_next:
                self._issue_copy(0)                   # first chunk of the next pass, behind this pass's last chunk
            b = self.g % 2
            main.wait_event(self.copied[b])
            out = model(self.bufs[b][:self.sizes[k]], T)
            loss = out.square().sum() * loss_scale
            loss.backward()
            total = total + loss.detach()
            self.freed[b].record(main)
            self.g += 1
        return total


def streamed_fwd_bwd(model, host_u0: torch.Tensor, T: float, loss_scale: float, chunk: int,
                     bufs=None, copy_stream=None, ramp: bool = True):
    """One end-to-end pass over a batch in pinned host memory (see :class:`StreamedBatch`; no prefetch across
    passes).  Returns the summed loss (0-dim device tensor)."""
    dev = next(model.parameters()).device
    return StreamedBatch(host_u0, chunk, dev, bufs=bufs, copy_stream=copy_stream, ramp=ramp).run(model, T, loss_scale)
