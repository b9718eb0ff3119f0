"""Fused 3-D RoPE + bf16 cast: oracle properties on CPU, kernel vs oracle on the GPU."""
import pytest
import torch

from oracle import rope as orope

DEV = "cuda"


def test_rope_oracle_properties():
    """The restated rope_apply is a per-pair rotation: identity at grid position (0,0,0), norm preserving, and two
    steps along one axis compose (angle additivity) -- the properties any implementation of the Wan formula has."""
    d, hds = 128, 3
    freqs = orope.wan_freqs(d)
    assert freqs.shape == (1024, 64)
    c = d // 2
    assert [c - 2 * (c // 3), c // 3, c // 3] == [22, 21, 21]
    g = torch.Generator().manual_seed(0)
    grid = torch.tensor([[3, 4, 5]])
    x = torch.randn(1, 70, hds, d, generator=g)                 # 60 grid tokens + 10 pass-through tokens
    y = orope.rope_apply(x, grid, freqs)
    assert torch.allclose(y[0, 0], x[0, 0], atol=1e-6)          # position (0,0,0): angle 0
    assert torch.equal(y[0, 60:], x[0, 60:])                    # beyond f*h*w: unchanged
    n_in = x.view(1, 70, hds, c, 2).norm(dim=-1)
    n_out = y.view(1, 70, hds, c, 2).norm(dim=-1)
    assert torch.allclose(n_in, n_out, atol=1e-5)
    # token 1 is (0,0,1): only the width block (last 21 pairs) rotates, by freqs[1] of that block
    w_rot = freqs[1, 43:]
    xc = torch.view_as_complex(x[0, 1].double().reshape(hds, c, 2))
    exp = torch.view_as_real(torch.cat([xc[:, :43], xc[:, 43:] * w_rot], dim=1)).flatten(1).float()
    assert torch.allclose(y[0, 1], exp, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_rope_cast_kernel_matches_oracle(dtype):
    from liteattention_b200.rope import rope_apply_bf16
    d, hds = 128, 5
    freqs = orope.wan_freqs(d)
    g = torch.Generator().manual_seed(1)
    grid = torch.tensor([[3, 7, 9], [2, 5, 11]])
    s = 200                                                      # 189 / 110 grid tokens, the rest pass through
    x = torch.randn(2, s, hds, d, generator=g).to(dtype)
    ref = orope.rope_apply(x.float(), grid, freqs)               # fp64 inside, fp32 out
    got = rope_apply_bf16(x.to(DEV), grid, freqs).cpu()
    assert got.dtype == torch.bfloat16 and got.shape == x.shape
    ref_bf = ref.to(torch.bfloat16)
    # fp32 rotation vs fp64 rotation can differ by one bf16 rounding step at most
    ulp = (ref.abs().clamp_min(1e-3) * 2.0 ** -7)
    assert ((got.float() - ref).abs() <= ulp).all()
    assert (got != ref_bf).float().mean() < 0.02
    assert torch.equal(got[0, 189:], x[0, 189:].to(torch.bfloat16))
    # strided input (a fused-QKV view) goes through without a copy
    big = torch.randn(2, s, 3, hds, d, generator=g).to(dtype).to(DEV)
    xv = big[:, :, 1]
    got2 = rope_apply_bf16(xv, grid, freqs).cpu()
    ref2 = orope.rope_apply(xv.cpu().float(), grid, freqs)
    assert ((got2.float() - ref2).abs() <= (ref2.abs().clamp_min(1e-3) * 2.0 ** -7)).all()


@pytest.mark.gpu
def test_rope_cast_wan_shape_properties():
    """Full Wan2.1-14B shape (21 x 45 x 80 grid, 40 heads): too big for the fp64 oracle, so size-independent
    properties: norm preservation per complex pair and exact agreement with the small-shape kernel on a slice."""
    from liteattention_b200.rope import rope_apply_bf16
    d, hds, grid = 128, 40, torch.tensor([[21, 45, 80]])
    s = 21 * 45 * 80
    freqs = orope.wan_freqs(d)
    g = torch.Generator(device=DEV).manual_seed(2)
    x = torch.randn(1, s, hds, d, device=DEV, generator=g)
    y = rope_apply_bf16(x, grid, freqs)
    rows = torch.randint(0, s, (64,), generator=torch.Generator().manual_seed(3))
    ref = orope.rope_apply(torch.zeros(1, s, 1, d), grid, freqs)  # cheap: only to fail loudly if the oracle breaks
    assert ref.shape == (1, s, 1, d)
    xs, ys = x[0, rows.to(DEV)].cpu(), y[0, rows.to(DEV)].float().cpu()
    n_in = xs.view(64, hds, d // 2, 2).norm(dim=-1)
    n_out = ys.view(64, hds, d // 2, 2).norm(dim=-1)
    assert torch.allclose(n_in, n_out, rtol=2e-2, atol=2e-2)
    # the same rows through the oracle (token index -> grid position is what the slice must reproduce)
    full_ref = []
    for t in rows.tolist():
        xt = torch.zeros(1, s, hds, d)
        full_ref.append(t)
    f_, y_, x_ = rows // (45 * 80), (rows // 80) % 45, rows % 80
    fr = freqs.split([22, 21, 21], dim=1)
    ang = torch.cat([fr[0][f_], fr[1][y_], fr[2][x_]], dim=1)                     # [64, 64] complex
    xc = torch.view_as_complex(xs.double().reshape(64, hds, d // 2, 2))
    exp = torch.view_as_real(xc * ang[:, None, :]).flatten(2).float()
    assert ((ys - exp).abs() <= exp.abs().clamp_min(1e-3) * 2.0 ** -7).all()
