#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/check2_pytest.log 2>&1; echo "pytest -m gpu rc=$?"
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/check2_pytest.log | tail -n 12
timeout 300 python tests/attn_prof.py 0 2>&1 | tee gpurun_out/attn_prof6.log | tail -n 4
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-vae --graph > gpurun_out/bench_graph.json 2> gpurun_out/bench_graph.err; echo "bench --graph rc=$?"; tail -n 3 gpurun_out/bench_graph.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_graph.json").read().strip().splitlines()[-1])
    print("graph value",round(d["value"],2),"ms",round(d["ms_per_step"],2),"e2e",round(d["e2e"]["value"],2), "launches", d["gpu_launches"])
except Exception as e: print("no json", e)
PY
