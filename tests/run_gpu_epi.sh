#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_stdit3_gpu.py tests/test_mmdit_gpu.py -m gpu -q -x > gpurun_out/epi_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/epi_pytest.log | tail -n 4
[ $rc -ne 0 ] && exit 0
timeout 200 python tests/gemm_tune.py > gpurun_out/gemm_tune_v6.log 2>&1; grep -E "^[a-z]|cta2 bn256|cta1 bn192|cta2 bn192" gpurun_out/gemm_tune_v6.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-vae > gpurun_out/bench_epi.json 2> gpurun_out/bench_epi.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_epi.json").read().strip().splitlines()[-1])
    print("value",round(d["value"],2),"ms",round(d["ms_per_step"],2), "clk", d["clocks"]["sm_mhz"], {k:(round(v["ms_per_step"],2)) for k,v in d["roofline"]["families"].items()})
except Exception as e: print("no json", e)
PY
