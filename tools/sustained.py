#!/usr/bin/env python3
"""Sustained (power-capped) throughput of la_fwd_kernel: run back-to-back for ~SECS seconds, report TFLOP/s of the
last half and the SM clock / power seen by nvidia-smi."""
import os, subprocess, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from liteattention_b200 import _native as N
B, S, H, D = 1, int(os.environ.get("S", 32768)), int(os.environ.get("H", 16)), 128
secs = float(os.environ.get("SECS", 3))
q = torch.randn(B, S, H, D, device="cuda", dtype=torch.bfloat16); k = torch.randn_like(q); v = torch.randn_like(q)
out = torch.empty_like(q); lse = torch.empty(B, H, S, device="cuda")
N.fwd(q, k, v, out, lse, D ** -0.5); torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record(); N.fwd(q, k, v, out, lse, D ** -0.5); e[1].record(); torch.cuda.synchronize()
one = e[0].elapsed_time(e[1])
n = max(4, int(secs * 1e3 / one))
smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "200"],
                       stdout=subprocess.PIPE, text=True)
e[0].record()
for _ in range(n // 2): N.fwd(q, k, v, out, lse, D ** -0.5)
e[1].record()
for _ in range(n // 2): N.fwd(q, k, v, out, lse, D ** -0.5)
e[2].record(); torch.cuda.synchronize()
smi.terminate(); lines = [l.split(",") for l in smi.stdout.read().strip().splitlines() if "," in l]
fl = 4.0 * B * H * S * S * D
t1, t2 = e[0].elapsed_time(e[1]) / (n // 2), e[1].elapsed_time(e[2]) / (n // 2)
clk = sorted(float(l[0]) for l in lines[len(lines)//2:]) if lines else [0]
pw = max(float(l[1]) for l in lines) if lines else 0
print(f"first {one:.3f} ms ({fl/one/1e9:.0f} TF) | 1st half {fl/t1/1e9:.0f} TF | 2nd half {fl/t2/1e9:.0f} TF | clk~{clk[len(clk)//2]:.0f} MHz, max power {pw:.0f} W")
