#!/bin/bash
# Instrumented build of the library (-DOSB_PP_TRACE -DOSB_TA_TRACE: per-role clock64 timeline of CTA 0 in the ping-pong attention
# kernel); used by tests/pp_trace.py only.  Output: open-sora_b200/osb200/libosb200_trace.so (git-ignored).
set -e
cd "$(dirname "$0")/.."
mkdir -p open-sora_b200/csrc/build_trace
objs=""
for f in api gemm_sm100 attn_short_sm100 attn_tiles_sm100 elementwise vae_ops; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -DOSB_PP_TRACE -DOSB_TA_TRACE -Xcompiler -fPIC \
    -c open-sora_b200/csrc/$f.cu -o open-sora_b200/csrc/build_trace/$f.o &
  objs="$objs open-sora_b200/csrc/build_trace/$f.o"
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o open-sora_b200/osb200/libosb200_trace.so $objs -cudart static
echo built open-sora_b200/osb200/libosb200_trace.so
