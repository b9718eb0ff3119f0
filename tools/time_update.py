#!/usr/bin/env python3
"""Times la_skip_update_kernel alone at the Wan shape: plain 42 % / 77 % / dense lists and a must-do list (CUDA events
over back-to-back launches; `LITEATTN_B200_LIB=...` selects another build)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from liteattention_b200 import LiteAttention, _native as N, synth
B, S, H, D = 1, 75600, 40, 128
qt, kt = synth.tile_counts(S)
stat = torch.randn(B, H, qt, kt, device="cuda") * 3
q = torch.empty(B, 1, H, D, device="cuda", dtype=torch.bfloat16)
md = LiteAttention._expand_must_do_list([511, 0], (B, H, qt, kt + 1), q, q)
md2 = LiteAttention._expand_must_do_list([75000, 74000, 40000, 39000, 511, 0], (B, H, qt, kt + 1), q, q)
cases = [("dense", LiteAttention.init_skip_list(B, S, H, D, False, torch.bfloat16, "cuda")[0], None),
         ("42", synth.exact_sparsity_list(B, H, qt, kt, 0.42, seed=1234, device="cuda")[0], None),
         ("42 bernoulli", synth.random_skip_list(B, H, qt, kt, 0.42, seed=99, device="cuda")[0], None),
         ("77", synth.exact_sparsity_list(B, H, qt, kt, 0.77, seed=1234, device="cuda")[0], None),
         ("42 + must-do [511,0]", synth.exact_sparsity_list(B, H, qt, kt, 0.42, seed=1234, device="cuda")[0], md),
         ("42 + must-do 3 ranges", synth.exact_sparsity_list(B, H, qt, kt, 0.42, seed=1234, device="cuda")[0], md2),
         ("dense + must-do [511,0]", LiteAttention.init_skip_list(B, S, H, D, False, torch.bfloat16, "cuda")[0], md)]
for name, rl, m in cases:
    wl = torch.zeros_like(rl)
    for thr in (-1.0,):
        for _ in range(5): N.skip_update(rl, m, wl, stat, B, H, qt, kt, thr)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): N.skip_update(rl, m, wl, stat, B, H, qt, kt, thr)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 50 * 1e3
        by = synth.update_bytes(rl, wl)
        print(f"{name:26s} thr {thr:5.1f}: {us:7.1f} us per launch (back to back), {by/1e6:6.1f} MB algorithmic = {by/us/1e3:7.1f} GB/s")
