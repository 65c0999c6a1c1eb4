#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/check_pytest.log 2>&1; echo "pytest -m gpu rc=$?"
grep -E "passed|failed|Error|error" gpurun_out/check_pytest.log | tail -n 8
timeout 300 python tests/attn_prof.py 0 1 3 2>&1 | tee gpurun_out/attn_prof5.log | tail -n 10
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_check.json 2> gpurun_out/bench_check.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_check.json").read().strip().splitlines()[-1])
print("value",round(d["value"],2),"ms",round(d["ms_per_step"],2),"e2e",round(d["e2e"]["value"],2)); print(json.dumps(d["roofline"]["families"])); print(json.dumps(d.get("vae"))[:1500])
PY
