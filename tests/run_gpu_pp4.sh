#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attn" > gpurun_out/pp5_pytest.log 2>&1; echo "pytest attn rc=$?"
grep -E "passed|failed|^FAILED|^ERROR|timed out" gpurun_out/pp5_pytest.log | tail -n 20
timeout 300 python tests/attn_prof.py 0 4 2>&1 | tee gpurun_out/attn_prof_pp5.log | tail -n 8
timeout 300 python tests/pp_trace.py 120 > gpurun_out/pp_trace5.log 2>&1; echo "trace rc=$?"
grep "===" gpurun_out/pp_trace5.log
OSB_ATTN_IMPL=4 timeout 600 python -m pytest tests/test_stdit3_gpu.py -m gpu -q -x > gpurun_out/pp5_stdit3.log 2>&1; echo "stdit3 impl4 rc=$?"
grep -E "passed|failed|timed out" gpurun_out/pp5_stdit3.log | head -n 8
OSB_ATTN_IMPL=4 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-vae > gpurun_out/bench_pp5.json 2> gpurun_out/bench_pp5.err; echo "bench pp rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_pp5.json").read().strip().splitlines()[-1])
    print("pp value",round(d["value"],2),"ms",round(d["ms_per_step"],2),"e2e",round(d["e2e"]["value"],2), d["roofline"]["families"]["attn_short"])
except Exception as e: print("no json", e)
PY
