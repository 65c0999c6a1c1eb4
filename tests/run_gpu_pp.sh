#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attn or ln_modulate" > gpurun_out/pp8_pytest.log 2>&1; rc=$?; echo "pytest attn rc=$rc"
grep -E "passed|failed|^FAILED|^ERROR|timed out" gpurun_out/pp8_pytest.log | tail -n 6
[ $rc -ne 0 ] && exit 0
timeout 90 python tests/attn_prof.py 0 4 > gpurun_out/attn_prof_pp8.log 2>&1; rc=$?; echo "attn_prof rc=$rc"; grep -E "attn impl|timed out" gpurun_out/attn_prof_pp8.log | head -n 8
[ $rc -ne 0 ] && exit 0
timeout 60 python tests/pp_trace.py 130 > gpurun_out/pp_trace7.log 2>&1; rc=$?; echo "trace rc=$rc"
grep "===" gpurun_out/pp_trace7.log
[ $rc -ne 0 ] && exit 0
OSB_ATTN_IMPL=4 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-vae > gpurun_out/bench_pp8.json 2> gpurun_out/bench_pp8.err; echo "bench pp rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_pp8.json").read().strip().splitlines()[-1])
    print("pp value",round(d["value"],2),"ms",round(d["ms_per_step"],2), {k:(round(v["ms_per_step"],2)) for k,v in d["roofline"]["families"].items()})
except Exception as e: print("no json", e)
PY
