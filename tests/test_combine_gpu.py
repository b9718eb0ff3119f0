"""GPU: LSE-combine kernel (partial-attention composition, README.md:222-250 of the reference; the op
lite_attention::fwd_combine, flash_api.cpp:1620-1720) against an fp32 torch combine of the same partials."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _torch_combine(o_parts, lse_parts):
    """fp32: lse = logsumexp_i lse_i;  out = sum_i exp(lse_i - lse) * o_i.   o_i (b,s,h,d), lse_i (b,h,s)."""
    L = torch.stack([l.float() for l in lse_parts])                       # (n,b,h,s)
    lse = torch.logsumexp(L, dim=0)
    w = torch.exp(L - lse[None]).nan_to_num(0.0)                          # all -inf -> 0
    O = torch.stack([o.float() for o in o_parts])                         # (n,b,s,h,d)
    out = (w.permute(0, 1, 3, 2)[..., None] * O).sum(0)
    return out, lse


@pytest.mark.parametrize("n,b,s,h", [(2, 1, 300, 2), (3, 2, 1056, 3), (8, 1, 129, 1)])
def test_combine_matches_fp32_torch(native_lib, n, b, s, h):
    from liteattention_b200 import flash_attn_combine
    g = torch.Generator().manual_seed(n * 100 + s)
    o_parts = [torch.randn(b, s, h, 128, generator=g).to(torch.bfloat16).to(DEV) for _ in range(n)]
    lse_parts = [(torch.randn(b, h, s, generator=g) * 4).to(DEV) for _ in range(n)]
    lse_parts[-1][:, :, ::7] = float("-inf")                              # a shard that saw nothing for some rows
    out, lse = flash_attn_combine(o_parts, lse_parts)
    o_ref, lse_ref = _torch_combine(o_parts, lse_parts)
    assert (lse - lse_ref).abs().max() < 1e-5
    assert ((out.float() - o_ref).abs() <= 2.0 ** -8 * o_ref.abs() + 1e-5).all()     # one bf16 rounding (RN) of the fp32 result
    # fp32 result buffer
    buf = torch.empty(b, s, h, 128, device=DEV)
    out32, _ = flash_attn_combine(o_parts, lse_parts, out=buf)
    assert out32.data_ptr() == buf.data_ptr() and (out32 - o_ref).abs().max() < 1e-5
    with pytest.raises(RuntimeError):
        flash_attn_combine(o_parts, lse_parts, out=buf[:, ::2])           # non-contiguous out is rejected


def test_reference_op_fwd_combine(native_lib):
    """The reference's op surface: out_partial (n,b,s,h,d) fp32, lse_partial (n,b,s,h) contiguous in seqlen."""
    import liteattention_b200  # noqa: F401  (registers the op)
    g = torch.Generator().manual_seed(7)
    n, b, s, h = 4, 2, 500, 3
    op = torch.randn(n, b, s, h, 128, generator=g).to(DEV)
    lp = (torch.randn(n, b, h, s, generator=g) * 3).to(DEV).transpose(2, 3)   # (n,b,s,h), stride(-2) == 1
    out, lse = torch.ops.lite_attention.fwd_combine(op, lp, None, torch.bfloat16)
    o_ref, lse_ref = _torch_combine(list(op), [lp[i].transpose(1, 2) for i in range(n)])
    assert out.dtype == torch.bfloat16 and out.shape == (b, s, h, 128) and lse.shape == (b, s, h)
    assert (lse.transpose(1, 2) - lse_ref).abs().max() < 1e-5
    assert ((out.float() - o_ref).abs() <= 2.0 ** -8 * o_ref.abs() + 1e-5).all()     # one bf16 rounding (RN) of the fp32 result
    out32, _ = torch.ops.lite_attention.fwd_combine(op, lp)
    assert out32.dtype == torch.float32 and (out32 - o_ref).abs().max() < 1e-5
    with pytest.raises(RuntimeError):
        torch.ops.lite_attention.fwd_combine(op.half(), lp)


def test_split_kv_attention_recombines(native_lib):
    from liteattention_b200 import flash_attn_func, flash_attn_combine
    g = torch.Generator().manual_seed(0)
    b, s, h = 2, 1056, 3
    q, k, v = (torch.randn(b, s, h, 128, generator=g).to(torch.bfloat16).to(DEV) for _ in range(3))
    parts = [flash_attn_func(q, k[:, a:e], v[:, a:e], return_softmax_lse=True) for a, e in ((0, 352), (352, 880), (880, s))]
    out, lse = flash_attn_combine([p[0] for p in parts], [p[1] for p in parts])
    # against plain fp32 attention over the whole K/V (independent of the kernels under test)
    qf, kf, vf = (t.float().permute(0, 2, 1, 3) for t in (q, k, v))
    sc = (qf @ kf.transpose(-1, -2)) * 128 ** -0.5
    lse_ref = torch.logsumexp(sc, -1)
    o_ref = (torch.softmax(sc, -1) @ vf).permute(0, 2, 1, 3)
    assert (lse - lse_ref).abs().max() < 1e-3
    assert (out.float() - o_ref).abs().max() < 1e-2
    # -inf partial (an all-skipped shard) contributes nothing
    full, lse_full = flash_attn_func(q, k, v, return_softmax_lse=True)
    dead_o = torch.zeros_like(full)
    dead_l = torch.full_like(lse_full, float("-inf"))
    out2, lse2 = flash_attn_combine([full, dead_o], [lse_full, dead_l])
    assert torch.equal(out2, full) and torch.allclose(lse2, lse_full, atol=1e-6)
