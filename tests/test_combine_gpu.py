"""GPU: LSE-combine kernel (partial-attention composition, README.md:222-250 of the reference)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_split_kv_attention_recombines(native_lib):
    from liteattention_b200 import flash_attn_func, flash_attn_combine
    g = torch.Generator().manual_seed(0)
    b, s, h = 2, 1056, 3
    q, k, v = (torch.randn(b, s, h, 128, generator=g).to(torch.bfloat16).to(DEV) for _ in range(3))
    full, lse_full = flash_attn_func(q, k, v, return_softmax_lse=True)
    parts = [flash_attn_func(q, k[:, a:e], v[:, a:e], return_softmax_lse=True) for a, e in ((0, 352), (352, 880), (880, s))]
    out, lse = flash_attn_combine([p[0] for p in parts], [p[1] for p in parts])
    assert (lse - lse_full).abs().max() < 1e-3
    assert (out.float() - full.float()).abs().max() < 1e-2
    # -inf partial (an all-skipped shard) contributes nothing
    dead_o = torch.zeros_like(full)
    dead_l = torch.full_like(lse_full, float("-inf"))
    out2, lse2 = flash_attn_combine([full, dead_o], [lse_full, dead_l])
    assert torch.equal(out2, full) and torch.allclose(lse2, lse_full, atol=1e-6)
