"""Batch-parallel multi-GPU driver: one process per GPU, each rank owns whole prompts (its Q/K/V and its
per-layer skip state, never communicated); the only collective is the gather of O over NVLink
(the reference has no distributed code on this path at all --
`grep torch.distributed hopper/` is empty, SeqParallelLiteAttention is bookkeeping only).

Two ways to get O to the destination rank:

* fused ("peer store", the default on GPUs when torch symmetric memory is available): the destination's output
  buffer [world, B, S, H, D] lives in symmetric memory; every rank maps it over NVLink/NVSwitch and passes ITS slot
  as the forward kernel's `out` pointer, so the kernel's epilogue stores O straight into the destination GPU while
  the other tiles are still computing -- compute and "gather" are one kernel, there is no collective kernel and no
  exposed tail, only a device-side barrier at the end of the step.  (The forward kernel needs no change: its O
  stores are plain 16-byte st.global from registers, a peer pointer is just a pointer.)
* NCCL gather, pipelined by head group (the plumbing baseline and the CPU/gloo-testable path): the attention kernel
  of head group g+1 runs on the compute stream while group g's O slab is gathered on a side stream; only the last
  group's gather is exposed.

The attention callable is injectable so the sharding / pipelining logic can be exercised on CPU with the gloo backend
(tests/test_dist_cpu.py).
"""
from typing import Callable, List, Optional

import torch
import torch.distributed as dist


def shard_batch(n_items: int, world_size: int, rank: int):
    """Contiguous batch shard [lo, hi) of rank `rank` (prompts are independent: no data-path collective)."""
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def head_groups(num_heads: int, num_groups: int) -> List[slice]:
    num_groups = max(1, min(num_groups, num_heads))
    base, rem = divmod(num_heads, num_groups)
    out, lo = [], 0
    for g in range(num_groups):
        hi = lo + base + (1 if g < rem else 0)
        out.append(slice(lo, hi))
        lo = hi
    return out


class BatchParallelLiteAttention:
    """Per-rank LiteAttention state + pipelined gather of O to `dst`.

    attn_factory() must return a callable (q, k, v) -> O for one head group (a LiteAttention object on GPU; each
    head group keeps its own skip state because lists are indexed [batch, head, qtile]).
    """

    def __init__(self, attn_factory: Callable[[], Callable], num_heads: int, num_groups: int = 5, dst: int = 0,
                 group: Optional[dist.ProcessGroup] = None, gather: bool = True, peer_store: bool = False):
        self.peer_store = bool(peer_store and gather and dist.is_initialized() and dist.get_world_size(group) > 1)
        if self.peer_store:
            num_groups = 1                      # nothing to overlap: the epilogue IS the transfer
        self._symm = self._hdl = self._peer_slot = None
        self.fallback_reason = None
        self.groups = head_groups(num_heads, num_groups)
        self.attn = [attn_factory() for _ in self.groups]
        self.dst = dst
        self.pg = group
        self.gather = gather and dist.is_initialized() and dist.get_world_size(group) > 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._comm_stream = None
        self._recv = None          # on dst: per group, list over ranks of [B, S, Hg, D] buffers (reused)

    def _streams(self, device):
        if device.type == "cuda" and self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=device)
        return self._comm_stream

    def _setup_peer(self, q: torch.Tensor):
        import torch.distributed._symmetric_memory as symm_mem
        shape = (self.world,) + tuple(q.shape)
        pg = self.pg if self.pg is not None else dist.group.WORLD
        self._symm = symm_mem.empty(shape, dtype=q.dtype, device=q.device)      # same size on every rank (symmetric)
        self._hdl = symm_mem.rendezvous(self._symm, group=pg)
        # rank r's view of slot r of the DESTINATION's buffer (peer memory mapped over NVLink)
        self._peer_slot = self._hdl.get_buffer(self.dst, shape, q.dtype)[self.rank]

    def __call__(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor):
        """q, k, v: this rank's (B_local, S, H, D).  Returns (local O slabs per head group,
        gathered) where `gathered` is, on `dst`, a list over head groups of lists over ranks of
        (B_local, S, Hg, D) tensors (None elsewhere / when not gathering).  With peer_store the first element is
        [this rank's slot of the destination buffer] and `gathered` is [[slot 0, slot 1, ...]] on `dst`."""
        if self.peer_store and self._hdl is None:
            try:
                self._setup_peer(q)
            except Exception as e:  # noqa: BLE001 -- no symmetric memory / no P2P on this box: NCCL gather instead
                import warnings
                warnings.warn(f"BatchParallelLiteAttention: peer-store gather unavailable ({e}); using the NCCL gather")
                self.peer_store = False
                self.fallback_reason = str(e)[:200]
        if self.peer_store:
            # `gathered` from the previous call are views of the destination's symmetric buffer: nobody may store into
            # it until the destination's stream has passed its reads of step N-1 (write-after-read across ranks).
            # The result of a call is therefore valid until the next call on ANY rank reaches this barrier.
            self._hdl.barrier()
            o = self.attn[0](q, k, v, out=self._peer_slot)
            self._hdl.barrier()        # device-side, stream-ordered: every rank's stores have landed at dst
            gathered = [[self._symm[r] for r in range(self.world)]] if self.rank == self.dst else None
            return [o], gathered
        cuda = q.device.type == "cuda"
        comm = self._streams(q.device)
        outs, works = [], []
        if self.gather and self.rank == self.dst and self._recv is None:
            self._recv = [[torch.empty((q.shape[0], q.shape[1], g.stop - g.start, q.shape[3]), dtype=q.dtype,
                                       device=q.device) for _ in range(self.world)] for g in self.groups]
        for gi, g in enumerate(self.groups):
            o = self.attn[gi](q[:, :, g], k[:, :, g], v[:, :, g])
            o = o if o.is_contiguous() else o.contiguous()
            outs.append(o)
            if not self.gather:
                continue
            recv = self._recv[gi] if self.rank == self.dst else None
            if cuda:
                ev = torch.cuda.Event()
                ev.record()
                comm.wait_event(ev)
                with torch.cuda.stream(comm):
                    works.append(dist.gather(o, recv, dst=self.dst, group=self.pg, async_op=True))
                o.record_stream(comm)
            else:
                works.append(dist.gather(o, recv, dst=self.dst, group=self.pg, async_op=True))
        for w in works:
            w.wait()
        if cuda and self.gather:
            torch.cuda.current_stream(q.device).wait_stream(comm)
        return outs, (self._recv if (self.gather and self.rank == self.dst) else None)


def ulysses_scatter_heads(x: torch.Tensor, world: int, group=None) -> torch.Tensor:
    """(B, S_local, H, D) sequence-sharded -> (B, S_local * world, H / world, D) head-sharded (tokens in rank order).
    One all_to_all_single; works on any backend (gloo in the CPU tests, NCCL over NVLink on GPUs)."""
    b, sl, h, dd = x.shape
    hl = h // world
    send = x.view(b, sl, world, hl, dd).permute(2, 0, 1, 3, 4).contiguous()        # [dst rank, B, S_local, Hl, D]
    recv = torch.empty_like(send)                                                  # [src rank, B, S_local, Hl, D]
    dist.all_to_all_single(recv, send, group=group)
    return recv.permute(1, 0, 2, 3, 4).reshape(b, world * sl, hl, dd)


class UlyssesLiteAttention:
    """Head-parallel attention of ONE prompt across the ranks of `group` (the reference's
    SeqParallelLiteAttention only keeps one skip state per split, lite_attention.py:322-345, and leaves the data
    movement to the caller).  Every rank holds S / world tokens of all H heads; the call

        o_local = attn(q_local, k_local, v_local)            # (B, S/world, H, D) -> (B, S/world, H, D)

    (1) re-shards q, k, v to all tokens x H / world heads with one all_to_all each (NCCL over NVLink),
    (2) runs the QK-Skip forward on this rank's heads (its skip lists cover exactly those heads, never communicated),
    (3) returns O to sequence sharding INSIDE the forward kernel: the epilogue stores every query row straight into
        the symmetric-memory output buffer of the rank that owns the token (peer stores over NVLink), followed by one
        device-side barrier.  There is no return all-to-all kernel.
    attn_factory() -> a LiteAttention-like callable accepting out= (a PeerScatter)."""

    def __init__(self, attn_factory: Callable[[], Callable], group: Optional[dist.ProcessGroup] = None):
        self.pg = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.attn = attn_factory()
        self._symm = self._hdl = self._scatter = None

    def _setup(self, q_local: torch.Tensor):
        import torch.distributed._symmetric_memory as symm_mem
        from . import _native
        b, sl, h, dd = q_local.shape
        pg = self.pg if self.pg is not None else dist.group.WORLD
        self._symm = symm_mem.empty((b, sl, h, dd), dtype=q_local.dtype, device=q_local.device)
        self._hdl = symm_mem.rendezvous(self._symm, group=pg)
        hl = h // self.world
        # peer r's buffer, restricted to the head columns this rank computes
        peers = [self._hdl.get_buffer(r, (b, sl, h, dd), q_local.dtype)[:, :, self.rank * hl:(self.rank + 1) * hl]
                 for r in range(self.world)]
        self._scatter = _native.PeerScatter(peers, rows_per_peer=sl)

    def __call__(self, q_local: torch.Tensor, k_local: torch.Tensor, v_local: torch.Tensor) -> torch.Tensor:
        b, sl, h, dd = q_local.shape
        assert h % self.world == 0, "heads must divide evenly over the ranks"
        if self._hdl is None or self._symm.shape != q_local.shape:
            self._setup(q_local)
        q, k, v = (ulysses_scatter_heads(t, self.world, self.pg) for t in (q_local, k_local, v_local))
        self._hdl.barrier()                 # everyone has consumed the previous step's O before it is overwritten
        self.attn(q, k, v, out=self._scatter)
        self._hdl.barrier()                 # every rank's rows have landed here
        return self._symm
