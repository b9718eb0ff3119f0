"""CPU, world_size 2, gloo: the batch-parallel sharding + head-group-pipelined gather of O (no GPU needed; the
attention callable is replaced by a deterministic stand-in so only the host-side plumbing is exercised)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from liteattention_b200.dist import BatchParallelLiteAttention, head_groups, shard_batch


def test_shard_batch_and_head_groups():
    assert [shard_batch(8, 8, r) for r in range(8)] == [(r, r + 1) for r in range(8)]
    assert [shard_batch(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert [shard_batch(2, 4, r) for r in range(4)] == [(0, 1), (1, 2), (2, 2), (2, 2)]
    gs = head_groups(40, 5)
    assert [g.stop - g.start for g in gs] == [8] * 5 and gs[0].start == 0 and gs[-1].stop == 40
    assert [g.stop - g.start for g in head_groups(7, 3)] == [3, 2, 2]
    assert len(head_groups(2, 5)) == 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q_all):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_batch(q_all.shape[0], world, rank)
        q = q_all[lo:hi]
        calls = []

        def factory():
            def attn(qq, kk, vv):
                calls.append(tuple(qq.shape))
                return qq * 2 + kk - vv * 0.5
            return attn

        bp = BatchParallelLiteAttention(factory, num_heads=q.shape[2], num_groups=3, dst=0)
        for step in range(2):                                    # buffers are reused across steps
            outs, gathered = bp(q + step, q, q)
            assert len(outs) == 3 and len(calls) == 3 * (step + 1)
            local = torch.cat(outs, dim=2)
            assert torch.equal(local, (q + step) * 2 + q - q * 0.5)
            if rank == 0:
                full = torch.cat([torch.cat([gathered[g][r] for g in range(3)], dim=2) for r in range(world)], dim=0)
                exp = (q_all + step) * 2 + q_all - q_all * 0.5
                assert torch.equal(full, exp)
            else:
                assert gathered is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_batch_parallel_gather_world2():
    torch.manual_seed(0)
    q_all = torch.randn(2, 16, 7, 8)
    mp.spawn(_worker, args=(2, _free_port(), q_all), nprocs=2, join=True)


def _ulysses_worker(rank, world, port, x_all):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from liteattention_b200.dist import ulysses_scatter_heads
        b, s, h, d = x_all.shape
        sl, hl = s // world, h // world
        mine = x_all[:, rank * sl:(rank + 1) * sl].contiguous()           # sequence shard, all heads
        got = ulysses_scatter_heads(mine, world)                           # all tokens, my heads
        assert got.shape == (b, s, hl, d)
        assert torch.equal(got, x_all[:, :, rank * hl:(rank + 1) * hl])
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_ulysses_head_scatter_layout_world2():
    """The all_to_all that turns sequence sharding into head sharding (the input side of UlyssesLiteAttention; the
    output side is the forward kernel's peer-store epilogue, GPU-only: tools/check_ulysses.py, tests/test_dist_gpu.py)."""
    torch.manual_seed(1)
    x_all = torch.randn(2, 12, 6, 8)
    mp.spawn(_ulysses_worker, args=(2, _free_port(), x_all), nprocs=2, join=True)
