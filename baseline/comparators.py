"""Dense Blackwell attention comparators for bench.py's `gpu_comparators` block (SURVEY.md section 8(d):
"GPU comparators available on the box (context, not the reference)").  BENCH-SIDE ONLY: nothing under
liteattention_b200/ imports this, and none of these kernels is on the product path.

Same device, same q/k/v (B, S, H, D) bf16, CUDA events on the current stream, `sleep` between kernels like the
reference's own benchmark does "to avoid residual power throttling" (hopper/_internal/benchmarks/benchmark_attn.py:312).
  * torch SDPA forced to the cuDNN backend, and to the flash backend (torch's bundled FA2)
  * pip flash_attn 2.8.x (FA2; its sm_80-class mma.sync kernel recompiled for the box)
  * CUTLASS examples/77_blackwell_fmha (C++ tcgen05 FMHA, fp16), built from the reference tree's vendored CUTLASS by
    baseline/build_comparators.sh into baseline/_ref/ -- runs as a subprocess on its own random data
"""
import os
import re
import subprocess
import time

HERE = os.path.dirname(os.path.abspath(__file__))
FMHA77 = os.path.join(HERE, "_ref", "cutlass_fmha77_fp16")


def _time(fn, steps, warmup):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def time_dense_comparators(q, k, v, steps=5, warmup=2, pause=1.0):
    """q, k, v: (B, S, H, D) bf16 on the GPU.  Returns {name: {"ms", "tflops"} | {"error"}} for the dense problem."""
    import torch
    import torch.nn.functional as F
    B, S, H, D = q.shape
    flops = 4.0 * B * H * S * float(k.shape[1]) * D
    res = {}

    def record(name, fn, note=None):
        try:
            ms = _time(fn, steps, warmup)
            res[name] = {"ms": ms, "tflops": flops / ms / 1e9}
            if note:
                res[name]["note"] = note
        except Exception as e:  # noqa: BLE001 -- a comparator that does not run on this box is reported, not fatal
            res[name] = {"error": f"{type(e).__name__}: {str(e)[:160]}"}
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
        time.sleep(pause)

    qt, kt_, vt = (t.transpose(1, 2) for t in (q, k, v))                    # (B, H, S, D) views
    try:
        from torch.nn.attention import SDPBackend, sdpa_kernel

        def sdpa(backend):
            def run():
                with sdpa_kernel(backend):
                    return F.scaled_dot_product_attention(qt, kt_, vt)
            return run
        record("torch_sdpa_cudnn", sdpa(SDPBackend.CUDNN_ATTENTION), f"cuDNN {torch.backends.cudnn.version()}")
        record("torch_sdpa_flash", sdpa(SDPBackend.FLASH_ATTENTION), "torch's bundled FA2")
    except Exception as e:  # noqa: BLE001
        res["torch_sdpa"] = {"error": str(e)[:160]}
    try:
        import flash_attn
        record("flash_attn_pip", lambda: flash_attn.flash_attn_func(q, k, v), f"flash_attn {flash_attn.__version__} (FA2)")
    except Exception as e:  # noqa: BLE001
        res["flash_attn_pip"] = {"error": f"{type(e).__name__}: {str(e)[:160]}"}

    if os.path.exists(FMHA77):
        for flag, name in (("--persistent", "cutlass_77_blackwell_fmha_persistent"), ("", "cutlass_77_blackwell_fmha")):
            cmd = [FMHA77, f"--b={B}", f"--h={H}", f"--q={S}", f"--k={k.shape[1]}", f"--d={D}", "--mask=no",
                   f"--iterations={steps}", f"--warmup_iterations={warmup}", "--verbose"] + ([flag] if flag else [])
            try:
                out = subprocess.run(cmd, capture_output=True, text=True, timeout=300).stdout
                m = re.search(r":\s*([0-9.eE+-]+)\s*TFLOPS/s\s*\n\s*t=([0-9.eE+-]+)ms", out)
                if m:
                    res[name] = {"ms": float(m.group(2)), "tflops": float(m.group(1)),
                                 "note": "fp16, CUTLASS 4.0 example built from the reference tree (baseline/build_comparators.sh)"}
                else:
                    res[name] = {"error": "could not parse: " + out[-200:].replace("\n", " | ")}
            except Exception as e:  # noqa: BLE001
                res[name] = {"error": f"{type(e).__name__}: {str(e)[:160]}"}
            time.sleep(pause)
    else:
        res["cutlass_77_blackwell_fmha"] = {"error": "baseline/_ref/cutlass_fmha77_fp16 not built (baseline/build_comparators.sh)"}
    return res
