#!/usr/bin/env python3
"""BASELINE config C4: Wan2.2-14B shape, per-step error-calibrated thresholds, sparsity sweep 0/21/42/57/77 %.
For each target: thr is calibrated on the video-like generator's first step (liteattention_b200.calibrate: one
forward + bisection over the 60-us update kernel), then a LiteAttention object runs 6 chained steps at that threshold
(re-calibrated every step towards the same target, the "per-timestep" part) and the last call is timed."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from liteattention_b200 import LiteAttention, synth
from liteattention_b200.calibrate import calibrate_threshold
S, H = int(os.environ.get("S", 75600)), int(os.environ.get("H", 40))
res = []
for target in (0.0, 0.21, 0.42, 0.57, 0.77):
    gen = synth.VideoLikeQKV(1, H, device="cuda", seq_len=None if S == 75600 else S, amp=float(os.environ.get("AMP", 14.0)))
    la = LiteAttention(enable_skipping=True, threshold=-60.0, max_batch_size=1)
    thrs, ms = [], 0.0
    for step in range(6):
        q, k, v = gen.next()
        if target > 0:
            thr, sp_c, _ = calibrate_threshold(q, k, v, target, read_list=la.read_list[:1] if la.read_list is not None else None)
            la.set_threshold(thr)
            thrs.append(round(thr, 2))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sp_read = la.last_sparsity(1)
        e0.record(); o = la(q, k, v); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        del q, k, v, o
    res.append({"target_sparsity": target, "sparsity_of_list_read_by_timed_call": round(sp_read, 4), "thresholds": thrs,
                "ms": round(ms, 2), "effective_tflops": round(synth.flops_dense(1, H, gen.S, gen.S, 128) / ms / 1e9, 1)})
    print(res[-1], file=sys.stderr, flush=True)
    del gen, la
    torch.cuda.empty_cache()
print(json.dumps({"config": f"C4: B=1 S={S} H={H} D=128 bf16, video-like generator, thresholds calibrated per step to a target list sparsity", "results": res}))
