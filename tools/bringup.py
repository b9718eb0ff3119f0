#!/usr/bin/env python3
"""GPU bring-up checks for the forward kernel against a plain torch fp32 reference (run under gpurun).
Not a test: prints diagnostics for debugging descriptors / pipelines."""
import os
import sys
import math
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from liteattention_b200 import _native as N  # noqa: E402

torch.manual_seed(0)
dev = "cuda"


def ref_attn(q, k, v, scale, tile_mask=None):
    # q,k,v (B,S,H,D) bf16 -> fp32 reference; tile_mask (B,H,qtiles,ktiles) bool or None
    qf, kf, vf = (t.float().permute(0, 2, 1, 3) for t in (q, k, v))
    s = torch.matmul(qf, kf.transpose(-1, -2)) * scale
    if tile_mask is not None:
        Sq, Sk = q.shape[1], k.shape[1]
        m = tile_mask.repeat_interleave(128, 2)[:, :, :Sq].repeat_interleave(176, 3)[:, :, :, :Sk]
        s = s.masked_fill(~m, float("-inf"))
    lse = torch.logsumexp(s, -1)
    p = torch.softmax(s, -1)
    o = torch.matmul(p, vf).permute(0, 2, 1, 3)
    return o, lse, s


def run(B, S, H, read_list=None, tile_mask=None, dbg=False, name=""):
    D = 128
    q = torch.randn(B, S, H, D, device=dev, dtype=torch.bfloat16)
    k = torch.randn(B, S, H, D, device=dev, dtype=torch.bfloat16)
    v = torch.randn(B, S, H, D, device=dev, dtype=torch.bfloat16)
    out = torch.full((B, S, H, D), float("nan"), device=dev, dtype=torch.bfloat16)
    lse = torch.full((B, H, S), float("nan"), device=dev, dtype=torch.float32)
    qt, kt = (S + 127) // 128, (S + 175) // 176
    stat = torch.full((B, H, qt, kt), float("nan"), device=dev, dtype=torch.float32)
    scale = D ** -0.5
    dbg_t = None
    dbg = False   # the raw-S dump hook was removed from the kernel after bring-up
    try:
        N.fwd(q, k, v, out, lse, scale, read_list, stat)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print(f"[{name}] FAILED: {e}")
        try:
            print("watchdog:", N.watchdog_read())
        except Exception as e2:  # noqa: BLE001
            print("watchdog read failed:", e2)
        raise
    o_ref, lse_ref, s_ref = ref_attn(q, k, v, scale, tile_mask)
    err_o = (out.float() - o_ref).abs().max().item()
    err_l = (lse - lse_ref).abs().max().item()
    print(f"[{name}] B={B} S={S} H={H}: max|O-ref|={err_o:.4e}  max|LSE-ref|={err_l:.4e}  "
          f"nan_out={int(out.float().isnan().sum())} nan_lse={int(lse.isnan().sum())}")
    if dbg:
        # raw S of first visited tile (= last K tile) of q-tile 0, head 0, batch 0
        n = kt - 1
        cols = min(176, S - n * 176)
        rows = min(128, S)
        s_exp = (s_ref[0, 0, :rows, n * 176:n * 176 + cols] / scale)
        d = (dbg_t[:rows, :cols] - s_exp).abs()
        print(f"   dbg S: max err {d.max().item():.4e}; kernel S[0,:4]={dbg_t[0,:4].tolist()} ref={s_exp[0,:4].tolist()}")
        if d.max().item() > 0.1:
            bad = (d > 0.1).nonzero()
            print("   first bad idx:", bad[:8].tolist(), "count", bad.shape[0])
            print("   row0 kernel:", dbg_t[0, :16].tolist())
            print("   row0 ref   :", s_exp[0, :16].tolist())
            print("   col0 kernel:", dbg_t[:16, 0].tolist())
            print("   col0 ref   :", s_exp[:16, 0].tolist())
    return err_o, err_l, out, o_ref, stat


if __name__ == "__main__":
    print("device:", torch.cuda.get_device_name(0))
    run(1, 176, 1, dbg=True, name="1tile")
    run(1, 128, 1, dbg=True, name="ragged-k")
    run(1, 352, 1, dbg=True, name="2tiles")
    run(1, 1000, 2, dbg=True, name="6tiles")
    run(2, 4096, 4, name="4k")
    # list-gated: random 50% mask
    B, S, H = 1, 2048, 2
    qt, kt = (S + 127) // 128, (S + 175) // 176
    g = torch.Generator().manual_seed(1234)
    keep = torch.rand(B, H, qt, kt, generator=g) < 0.5
    keep[..., kt - 1] = True
    rl = torch.zeros(B, H, qt, kt + 1, dtype=torch.int32)
    for b in range(B):
        for h in range(H):
            for m in range(qt):
                ent = []
                n = kt - 1
                while n >= 0:
                    if keep[b, h, m, n]:
                        s = n
                        while n - 1 >= 0 and keep[b, h, m, n - 1]:
                            n -= 1
                        ent += [s, n]
                    n -= 1
                rl[b, h, m, 0] = len(ent)
                rl[b, h, m, 1:1 + len(ent)] = torch.tensor(ent, dtype=torch.int32)
    run(B, S, H, read_list=rl.to(dev), tile_mask=keep.to(dev), name="list50")
    # timing at a mid size
    B, S, H, D = 1, 16384, 8, 128
    q = torch.randn(B, S, H, D, device=dev, dtype=torch.bfloat16)
    k = torch.randn_like(q); v = torch.randn_like(q)
    out = torch.empty_like(q); lse = torch.empty(B, H, S, device=dev)
    for _ in range(3):
        N.fwd(q, k, v, out, lse, D ** -0.5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        N.fwd(q, k, v, out, lse, D ** -0.5)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 4 * B * H * S * S * D
    print(f"dense S={S} H={H}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s")
