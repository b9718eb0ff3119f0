#!/usr/bin/env python3
"""Small end-to-end pass over every kernel of the library for compute-sanitizer (memcheck / racecheck / initcheck /
synccheck): list-gated forward + update over three chained calls (plain, with a must-do list, compact state), the list
pack / unpack kernels, the LSE combine, the RoPE cast and a host-resident call.  No torch reference kernels run here, so
every finding belongs to this library."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from liteattention_b200 import LiteAttention, flash_attn_combine
from liteattention_b200.rope import rope_apply_bf16
S, H = int(os.environ.get("S", 1300)), int(os.environ.get("H", 3))
g = torch.Generator().manual_seed(0)
q, k, v = ((torch.randn(1, S, H, 128, generator=g) * (3.0 if i == 0 else 1.0)).to(torch.bfloat16) for i in range(3))
dq, dk, dv = q.cuda(), k.cuda(), v.cuda()
for kw in ({}, {"compact_state": True}):
    la = LiteAttention(True, -3.0, max_batch_size=1, **kw)
    for step in range(3):
        o = la(dq, dk, dv, must_do_list=[700, 350] if step == 2 else None)
    print("chained calls", kw, "sparsity", round(la.last_sparsity(1), 3), "checksum", float(o.float().abs().mean()))
la = LiteAttention(True, -3.0, max_batch_size=1)
oh = la(q.pin_memory(), k.pin_memory(), v.pin_memory()); la.wait_host_copies()
print("host-resident call checksum", float(oh.float().abs().mean()))
half = S // 2
parts = [LiteAttention(enable_skipping=False)(dq, dk[:, sl], dv[:, sl], return_softmax_lse=True) for sl in (slice(0, half), slice(half, S))]
oc, lc = flash_attn_combine([p[0] for p in parts], [p[1] for p in parts])
print("combine checksum", float(oc.float().abs().mean()))
f, hh, ww = 2, 5, 13
x = torch.randn(1, f * hh * ww, H, 128, device="cuda")
freqs = torch.polar(torch.ones(1024, 64), torch.randn(1024, 64))
xr = rope_apply_bf16(x, torch.tensor([[f, hh, ww]]), freqs)
torch.cuda.synchronize()
print("rope checksum", float(xr.float().abs().mean()))
