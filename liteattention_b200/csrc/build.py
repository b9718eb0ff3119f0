#!/usr/bin/env python3
"""Builds liteattention_b200/libliteattn_b200.so (C ABI, sm_100a only) with nvcc.  In-tree output so the
library travels with the repo snapshot to the GPU box.  Usage: python liteattention_b200/csrc/build.py [-v]"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "libliteattn_b200.so")
SOURCES = ["la_all.cu"]
DEPS = ["la_all.cu", "la_fwd_sm100.cu", "la_skip_update.cu", "la_combine.cu", "la_rope_cast.cu", "la_list_codec.cu", "la_api.cu", "la_ptx.cuh",
        "la_tmem_ptx.cuh", "la_kernels.h", os.path.join("..", "..", "include", "liteattn_b200.h")]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(os.path.join(HERE, d)) <= t for d in DEPS)


def build(verbose=False, force=False, out=None, defines=()):
    """out/defines: experimental variants (tools/build_variants.py); the product build uses neither."""
    if out is None and not force and up_to_date():
        return OUT
    out = out or OUT
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
           "-Xptxas", "-v", "--shared", "-Xcompiler", "-fPIC"] + [f"-D{d}" for d in defines] + [
           "-o", out] + [os.path.join(HERE, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libliteattn_b200.so")
    return out


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
