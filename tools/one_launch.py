#!/usr/bin/env python3
"""Three forward launches at a bench configuration -- the subject of bench.py's live `ncu --metrics dram__bytes_*`
measurement (roofline.traffic) and of tools/capture_profiles.sh."""
import argparse, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from liteattention_b200 import LiteAttention, _native as N, synth
ap = argparse.ArgumentParser()
ap.add_argument("--seq", type=int, default=75600)
ap.add_argument("--heads", type=int, default=40)
ap.add_argument("--sparsity", type=float, default=0.42)
ap.add_argument("--update", action="store_true", help="also run the list update kernel three times (thr = -1)")
a = ap.parse_args()
B, S, H, D = 1, a.seq, a.heads, 128
g = torch.Generator(device="cuda").manual_seed(1000)
q, k, v = (torch.randn(B, S, H, D, device="cuda", generator=g).to(torch.bfloat16) for _ in range(3))
qt, kt = synth.tile_counts(S)
rl = (synth.exact_sparsity_list(B, H, qt, kt, a.sparsity, seed=1234, device="cuda")[0] if a.sparsity > 0
      else LiteAttention.init_skip_list(B, S, H, D, False, torch.bfloat16, "cuda")[0])
out = torch.empty_like(q); lse = torch.empty(B, H, S, device="cuda"); stat = torch.empty(B, H, qt, kt, device="cuda")
for _ in range(3):
    N.fwd(q, k, v, out, lse, D ** -0.5, rl, stat)
if a.update:
    wl = torch.zeros_like(rl)
    for _ in range(3):
        N.skip_update(rl, None, wl, stat, B, H, qt, kt, -1.0)
torch.cuda.synchronize()
