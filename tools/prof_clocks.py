#!/usr/bin/env python3
"""Cycle accounting of la_fwd_kernel (needs a library built with -DLA_PROFILE_CLOCKS at tools/_lib_prof.so)."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from liteattention_b200 import _native as N
L = N.lib()
B, S, H, D = 1, int(os.environ.get("S", 32768)), int(os.environ.get("H", 16)), 128
q = torch.randn(B, S, H, D, device="cuda", dtype=torch.bfloat16); k = torch.randn_like(q); v = torch.randn_like(q)
out = torch.empty_like(q); lse = torch.empty(B, H, S, device="cuda")
buf = (ctypes.c_ulonglong * 16)()
for _ in range(2):
    N.fwd(q, k, v, out, lse, D ** -0.5)
torch.cuda.synchronize()
L.la_prof_read(buf, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); N.fwd(q, k, v, out, lse, D ** -0.5); e1.record(); torch.cuda.synchronize()
L.la_prof_read(buf, 1)
v_ = list(buf)
tiles = v_[15]
ctas = H * ((S + 127) // 128)
print(f"kernel {e0.elapsed_time(e1):.3f} ms, tiles {tiles}, per-SM tiles {tiles/148:.0f}")
names = ["wait S", "pass1", "mbox wait+stat", "exp/alpha + pass2", "pvdone wait + rescale", "wait st + arrive"]
tot = 0
for j, nm in enumerate(names):
    per = v_[j] / tiles          # summed over the two WG leaders -> per tile (each tile handled by one WG)
    tot += per
    print(f"  softmax {nm:24s} {per:8.0f} cyc/tile")
print(f"  softmax total per tile (one WG) {tot:8.0f}  => per-WG cycle for 2 tiles = {2*tot:.0f}")
names = ["issue QK(i+1) incl. wait K", "wait V", "wait P", "issue PV"]
for j, nm in enumerate(names):
    print(f"  mma {nm:28s} {v_[8+j]/tiles:8.0f} cyc/tile")
