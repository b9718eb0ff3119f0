#!/usr/bin/env python3
"""Replays tests/test_fwd_gpu.py::test_linearity_and_spot_checks_at_config2_size in a fresh process and says what differs."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from liteattention_b200 import _native as N, synth
b, s, h = 1, 32768, 16
g = torch.Generator(device="cuda").manual_seed(0)
q, k, v1, v2 = (torch.randn(b, s, h, 128, device="cuda", generator=g).to(torch.bfloat16) for _ in range(4))
qt, kt = synth.tile_counts(s)
rl, keep = synth.random_skip_list(b, h, qt, kt, 0.5, seed=1234, device="cuda")
outs, lses = [], []
for vv in (v1, v2, (v1.float() + v2.float()).to(torch.bfloat16), v1):
    o = torch.empty_like(q); l = torch.empty(b, h, s, device="cuda")
    N.fwd(q, k, vv, o, l, 128 ** -0.5, rl, None)
    outs.append(o.float()); lses.append(l)
import time
t_sync = time.time()
try:
    torch.cuda.synchronize()
except Exception as e:
    import ctypes
    print(f"SYNC FAILED after {time.time() - t_sync:.2f} s:", str(e).splitlines()[0])
    sys.exit(1)
dl = (lses[0] - lses[1]).abs()
lin = (outs[0] + outs[1] - outs[2]).abs()
rep = (outs[0] - outs[3]).abs()
print(f"LSE(v1) vs LSE(v2): max {dl.max().item():.4g}, rows differing {(dl > 0).sum().item()};  linearity max {lin.max().item():.4g};  "
      f"O(v1) first vs fourth launch: max {rep.max().item():.4g}, LSE first vs fourth {(lses[0]-lses[3]).abs().max().item():.4g}")
for name, d in (("lse01", dl), ("lse03", (lses[0] - lses[3]).abs())):
    if d.max() > 0:
        idx = (d > 0).nonzero()
        print("  ", name, "bad (b,h,row) sample:", idx[:10].tolist(), " q-tiles:", sorted(set((int(x[1]), int(x[2]) // 128) for x in idx.tolist()))[:10])
if lin.max() > 2e-2:
    idx = (lin.amax(dim=(0, 3)) > 2e-2).nonzero()
    print("   lin bad (row, head):", idx[:10].tolist(), "q-tiles:", sorted(set((int(x[1]), int(x[0]) // 128) for x in idx.tolist()))[:10])
