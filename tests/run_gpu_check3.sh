#!/bin/bash
# full gpu suite + PDL on/off A/B of the eager and graph-replayed step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/check3_pytest.log 2>&1; echo "pytest -m gpu rc=$?"
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/check3_pytest.log | tail -n 12
run() {  # name, env, extra flags
  env $2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-vae $3 > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err; echo "bench $1 rc=$?"
  python - "$1" <<PY
import json, sys
try:
    d=json.loads(open("gpurun_out/bench_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value",round(d["value"],2),"ms",round(d["ms_per_step"],2),"e2e",round(d["e2e"]["value"],2), "launches", d["gpu_launches"], "gemm TF/s", d["roofline"]["achieved"])
except Exception as e: print("no json", e)
PY
}
run pdl_on OSB_PDL=1 ""
run pdl_off OSB_PDL=0 ""
run pdl_on_graph OSB_PDL=1 "--graph"
run pdl_off_graph OSB_PDL=0 "--graph"
