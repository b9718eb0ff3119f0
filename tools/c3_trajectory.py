#!/usr/bin/env python3
"""BASELINE config C3: Wan2.1-14B self-attention shape (B=1, S=75600, H=40, D=128, bf16), evolving QK-Skip over 50
chained calls on ONE LiteAttention(threshold=-10) object, video-like synthetic Q/K: per-call ms (CUDA
events around LiteAttention.__call__ = forward + list update) and the sparsity of the list each call used."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from liteattention_b200 import LiteAttention, synth
steps = int(os.environ.get("STEPS", 50))
thr = float(os.environ.get("THR", -10.0))
S, H = int(os.environ.get("S", 75600)), int(os.environ.get("H", 40))
gen = synth.VideoLikeQKV(1, H, device="cuda", seq_len=None if S == 75600 else S, amp=float(os.environ.get("AMP", 14.0)))
la = LiteAttention(enable_skipping=True, threshold=thr, max_batch_size=1)
ms, sp = [], []
dense = synth.flops_dense(1, H, gen.S, gen.S, 128)
for t in range(steps):
    q, k, v = gen.next()
    sp.append(la.last_sparsity(1))              # sparsity of the list this call reads (0 on the first call)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); o = la(q, k, v); e1.record(); torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
    del q, k, v, o
pick = [i for i in (0, 9, 24, 49) if i < steps]
out = {"config": f"C3: B=1 S={gen.S} H={H} D=128 bf16, {steps} chained calls, threshold={thr}, video-like generator (amp 14, rho 0.95)",
       "ms_at_step": {str(i + 1): round(ms[i], 2) for i in pick}, "sparsity_read_at_step": {str(i + 1): round(sp[i], 4) for i in pick},
       "final_list_sparsity": round(la.last_sparsity(1), 4), "mean_ms": round(sum(ms) / len(ms), 2),
       "mean_effective_tflops": round(dense / (sum(ms) / len(ms)) / 1e9, 1), "ms": [round(x, 2) for x in ms],
       "sparsity": [round(x, 4) for x in sp]}
print(json.dumps(out))
