#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tests/attn_prof.py 0 4 2>&1 | tee gpurun_out/attn_prof_pp2.log | tail -n 8
timeout 300 python tests/pp_trace.py 150 > gpurun_out/pp_trace.log 2>&1; echo "trace rc=$?"
grep "===" gpurun_out/pp_trace.log
OSB_ATTN_IMPL=4 timeout 600 python -m pytest tests/test_stdit3_gpu.py -m gpu -q -x > gpurun_out/pp_stdit3.log 2>&1; echo "stdit3 impl4 rc=$?"
grep -E "passed|failed|timed out|Error" gpurun_out/pp_stdit3.log | head -n 12
