#!/usr/bin/env python3
"""Bandwidth of the pitched head-group copies (la_copy2d_async) the host-resident call uses, against whole-tensor
copies, at the Wan shape: one (1, 75600, 40, 128) bf16 tensor, pinned host <-> device."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from liteattention_b200 import _native as N
S, H, D = 75600, 40, 128
h = torch.empty(1, S, H, D, dtype=torch.bfloat16).pin_memory()
d = torch.empty(1, S, H, D, dtype=torch.bfloat16, device="cuda")
st = torch.cuda.current_stream()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def t(fn, nbytes, reps=3):
    fn(); torch.cuda.synchronize()
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
full = h.numel() * 2
print(f"whole tensor H2D {t(lambda: d.copy_(h, non_blocking=True), full):6.1f} GB/s   D2H {t(lambda: h.copy_(d, non_blocking=True), full):6.1f} GB/s")
for g in (1, 2, 4, 8, 13, 36, 40):
    fh = lambda: N.copy2d_async(d.view(S, H * D), h.view(S, H * D), 0, g * D, st)
    fd = lambda: N.copy2d_async(h.view(S, H * D), d.view(S, H * D), 0, g * D, st)
    print(f"{g:2d} heads ({g * D * 2:5d} B rows, pitch {H * D * 2} B): H2D {t(fh, full * g // H):6.1f} GB/s   D2H {t(fd, full * g // H):6.1f} GB/s")
