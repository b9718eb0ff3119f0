#!/bin/bash
# Builds the dense Blackwell comparator named in SURVEY.md section 8(d): the vendored CUTLASS example
# 77_blackwell_fmha (C++ tcgen05 FMHA, fp16 -- same tensor rate as bf16), compiled from the sources where they lie
# under /root/reference into baseline/_ref/ (git-ignored, travels to the GPU box with the snapshot).  ~3 min of nvcc.
# Only bench.py's gpu_comparators leg runs the binary; nothing in the product path touches it.
set -e
REF=${1:-/root/reference}/csrc/cutlass
OUT=$(dirname "$(readlink -f "$0")")/_ref
mkdir -p "$OUT"
[ -d "$REF/examples/77_blackwell_fmha" ] || { echo "no CUTLASS tree at $REF"; exit 0; }
[ -x "$OUT/cutlass_fmha77_fp16" ] && { echo "up to date: $OUT/cutlass_fmha77_fp16"; exit 0; }
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --use_fast_math --expt-relaxed-constexpr \
  --expt-extended-lambda -ftemplate-backtrace-limit=0 -DFP16 \
  -I "$REF/include" -I "$REF/tools/util/include" -I "$REF/examples/common" -I "$REF/examples/77_blackwell_fmha" \
  -o "$OUT/cutlass_fmha77_fp16" "$REF/examples/77_blackwell_fmha/77_blackwell_fmha.cu"
echo "built $OUT/cutlass_fmha77_fp16"
