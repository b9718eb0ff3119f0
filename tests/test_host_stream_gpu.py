"""GPU: LiteAttention called on PINNED host tensors (host-resident / offloaded activations).  The object streams q, k, v
up and O down by head groups around per-group launches; everything observable -- O, LSE, the skip lists over several
chained calls -- must be bit-identical to the same calls on device tensors, and pageable CPU tensors must raise (there
is no CPU compute path)."""
import pytest
import torch

from liteattention_b200 import LiteAttention
from liteattention_b200.lite_attention import host_head_groups

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _qkv(b, s, h, seed):
    g = torch.Generator().manual_seed(seed)
    return [(torch.randn(b, s, h, 128, generator=g) * (3.0 if i == 0 else 1.0)).to(torch.bfloat16) for i in range(3)]


@pytest.mark.parametrize("b,s,h,compact", [(1, 1500, 8, False), (2, 777, 5, False), (1, 2100, 40, True)])
def test_host_call_equals_device_call(native_lib, b, s, h, compact):
    q, k, v = _qkv(b, s, h, seed=b * 100 + h)
    hq, hk, hv = (t.pin_memory() for t in (q, k, v))
    dq, dk, dv = (t.to(DEV) for t in (q, k, v))
    la_d = LiteAttention(True, -4.0, max_batch_size=b, compact_state=compact)
    la_h = LiteAttention(True, -4.0, max_batch_size=b, compact_state=compact)
    for step in range(4):
        o_d, lse_d = la_d(dq, dk, dv, return_softmax_lse=True)
        o_h, lse_h = la_h(hq, hk, hv, return_softmax_lse=True)
        assert o_h.device.type == "cpu" and o_h.is_pinned()
        la_h.wait_host_copies()
        assert torch.equal(o_h, o_d.cpu()), f"step {step}: O differs between the host-streamed and the device call"
        assert torch.equal(lse_h, lse_d)
        rd, rh = la_d.read_list, la_h.read_list
        assert torch.equal(rd[:b, ..., 0], rh[:b, ..., 0])
        ln = int(rd[:b, ..., 0].max())
        assert torch.equal(rd[:b, ..., :ln + 1], rh[:b, ..., :ln + 1]), f"step {step}: skip lists differ"
    assert la_d.last_sparsity(b) == la_h.last_sparsity(b)


def test_back_to_back_host_calls_without_waiting(native_lib):
    """Calls queued while the previous one is still in flight take the coarse head-group schedule and the other staging
    slot; results and lists must not change."""
    b, s, h = 1, 6000, 40
    q, k, v = _qkv(b, s, h, seed=21)
    hq, hk, hv = (t.pin_memory() for t in (q, k, v))
    dq, dk, dv = (t.to(DEV) for t in (q, k, v))
    la_d = LiteAttention(True, -4.0, max_batch_size=b)
    la_h = LiteAttention(True, -4.0, max_batch_size=b)
    outs = [torch.empty_like(q).pin_memory() for _ in range(5)]
    refs = []
    for i in range(5):
        refs.append(la_d(dq, dk, dv))
        la_h(hq, hk, hv, out=outs[i])                                  # no wait in between
    la_h.wait_host_copies()
    torch.cuda.synchronize()
    for i in range(5):
        assert torch.equal(outs[i], refs[i].cpu()), f"call {i}"
    ln = int(la_d.read_list[:b, ..., 0].max())
    assert torch.equal(la_d.read_list[:b, ..., :ln + 1], la_h.read_list[:b, ..., :ln + 1])


def test_staging_sets_are_recycled_safely(native_lib):
    """Only two staging sets (one per tensor shape) are kept; a third shape drains and drops them while earlier calls may
    still be in flight.  Results must not be disturbed."""
    las, outs, refs = [], [], []
    for i, s in enumerate((1100, 1900, 2700, 1100)):
        q, k, v = _qkv(1, s, 4, seed=40 + i)
        la = LiteAttention(enable_skipping=False)
        outs.append(la(q.pin_memory(), k.pin_memory(), v.pin_memory(), out=torch.empty_like(q).pin_memory()))   # no wait
        refs.append(LiteAttention(enable_skipping=False)(q.to(DEV), k.to(DEV), v.to(DEV)))
        las.append(la)
    for la in las:
        la.wait_host_copies()
    torch.cuda.synchronize()
    for o, r in zip(outs, refs):
        assert torch.equal(o, r.cpu())


def test_host_call_with_must_do_and_user_out(native_lib):
    b, s, h = 1, 1300, 6
    q, k, v = _qkv(b, s, h, seed=7)
    hq, hk, hv = (t.pin_memory() for t in (q, k, v))
    out = torch.empty_like(q).pin_memory()
    la_d = LiteAttention(True, -2.0, max_batch_size=1)
    la_h = LiteAttention(True, -2.0, max_batch_size=1)
    md = [700, 350]
    for _ in range(3):
        o_d = la_d(q.to(DEV), k.to(DEV), v.to(DEV), must_do_list=md)
        o_h = la_h(hq, hk, hv, must_do_list=md, out=out)
        assert o_h is out
        la_h.wait_host_copies()
        assert torch.equal(out, o_d.cpu())
    assert torch.equal(la_d.read_list[:, ..., 0], la_h.read_list[:, ..., 0])


def test_dense_mode_and_stream_ordered_join(native_lib):
    b, s, h = 1, 900, 4
    q, k, v = _qkv(b, s, h, seed=11)
    hq, hk, hv = (t.pin_memory() for t in (q, k, v))
    la = LiteAttention(enable_skipping=False)
    o_h = la(hq, hk, hv)
    la.join_host_copies()                                               # the current stream now waits for the download
    torch.cuda.current_stream().synchronize()
    ref = LiteAttention(enable_skipping=False)(q.to(DEV), k.to(DEV), v.to(DEV))
    assert torch.equal(o_h, ref.cpu())


def test_pageable_cpu_tensors_raise(native_lib):
    q, k, v = _qkv(1, 300, 2, seed=3)
    with pytest.raises((RuntimeError, NotImplementedError)):            # the op has no CPU kernel
        LiteAttention()(q, k, v)
    hq = q.pin_memory()
    with pytest.raises(RuntimeError, match="pinned"):
        LiteAttention()(hq, k, v)


def test_head_group_schedule():
    for heads in (1, 2, 3, 4, 5, 8, 12, 16, 24, 40, 64, 96):
        g = host_head_groups(heads)
        assert sum(g) == heads and min(g) >= 1
        gb = host_head_groups(heads, busy=True)
        assert sum(gb) == heads and min(gb) >= 1 and len(gb) <= 2
        if heads >= 20:
            assert g[0] <= heads // 10 and g[-1] <= heads // 8          # small first upload, small last download
