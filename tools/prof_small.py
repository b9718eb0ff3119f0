#!/usr/bin/env python3
"""Small dense forward (S=16384, H=8) for ncu captures of la_fwd_kernel."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from liteattention_b200 import _native as N
B, S, H, D = 1, int(os.environ.get("S", 16384)), int(os.environ.get("H", 8)), 128
q = torch.randn(B, S, H, D, device="cuda", dtype=torch.bfloat16); k = torch.randn_like(q); v = torch.randn_like(q)
out = torch.empty_like(q); lse = torch.empty(B, H, S, device="cuda")
for _ in range(3):
    N.fwd(q, k, v, out, lse, D ** -0.5)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    N.fwd(q, k, v, out, lse, D ** -0.5)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"dense S={S} H={H}: {ms:.3f} ms  {4*B*H*S*S*D/ms/1e9:.1f} TFLOP/s")
