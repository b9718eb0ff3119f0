#!/usr/bin/env python3
"""Cycle accounting of la_fwd_kernel.  Needs a -DLA_PROFILE_CLOCKS build:
     python tools/build_variants.py prof=LA_PROFILE_CLOCKS
     LITEATTN_B200_LIB=$PWD/tools/_build/lib_prof.so python tools/prof_clocks.py"""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from liteattention_b200 import _native as N
L = N.lib()
B, S, H, D = 1, int(os.environ.get("S", 32768)), int(os.environ.get("H", 16)), 128
q = torch.randn(B, S, H, D, device="cuda", dtype=torch.bfloat16); k = torch.randn_like(q); v = torch.randn_like(q)
out = torch.empty_like(q); lse = torch.empty(B, H, S, device="cuda")
buf = (ctypes.c_ulonglong * 32)()
for _ in range(2):
    N.fwd(q, k, v, out, lse, D ** -0.5)
torch.cuda.synchronize()
L.la_prof_read(buf, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); N.fwd(q, k, v, out, lse, D ** -0.5); e1.record(); torch.cuda.synchronize()
L.la_prof_read(buf, 1)
v_ = list(buf)
tiles = v_[6]
ms = e0.elapsed_time(e1)
print(f"kernel {ms:.3f} ms ({4*B*H*S*S*D/ms/1e9:.0f} TFLOP/s), speculative tiles {tiles}, per-SM tiles {tiles/148:.0f}, "
      f"=> {ms*1e-3/(tiles/148)*1e9:.0f} ns per tile per SM")
names = ["wait::ld S(i)", "exp loop + xchg + try_wait", "verdict", "P st + prefetch issue", "publish", "exact tile (total)"]
for base, tag in ((0, "warp 0 (cols 0-87)"), (20, "warp 4 (cols 88-175)")):
    tot = 0
    print(f" softmax {tag}")
    for j, nm in enumerate(names):
        per = v_[base + j] / max(tiles, 1)
        tot += per
        print(f"   {nm:28s} {per:8.0f} clk/tile")
    print(f"   {'total':28s} {tot:8.0f} clk/tile")
names = ["wait K full", "issue QK (8 MMA + 2 commit)", "wait V full", "wait P full", "issue PV (11 MMA + commits)"]
print(" MMA warp")
tot = 0
for j, nm in enumerate(names):
    per = v_[8 + j] / max(tiles, 1); tot += per
    print(f"   {nm:28s} {per:8.0f} clk/tile")
print(f"   {'total':28s} {tot:8.0f} clk/tile")
print(" TMA lane")
for j, nm in enumerate(["wait K empty", "wait V empty"]):
    print(f"   {nm:28s} {v_[16 + j] / max(tiles, 1):8.0f} clk/tile")
if v_[31]:
    n = v_[31]
    print(f" per item ({n} items, {tiles / n:.1f} speculative tiles each): descriptor -> S(0) ready {v_[27] / n:.0f} clk, first (exact) tile {v_[28] / n:.0f}, "
          f"epilogue {v_[29] / n:.0f}, whole item {v_[30] / n:.0f} clk  => fixed part ~{(v_[27] + v_[28] + v_[29]) / n:.0f} clk "
          f"= {(v_[27] + v_[28] + v_[29]) / max(v_[30], 1) * 100:.1f} % of the item")
