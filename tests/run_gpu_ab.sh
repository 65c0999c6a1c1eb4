#!/bin/bash
# full GPU suite on the default dispatch, then A/B of the step with the ping-pong kernel on two-block key sets
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/ab_pytest.log 2>&1; echo "pytest -m gpu rc=$?"
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/ab_pytest.log | tail -n 8
run() {
  env $2 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-vae > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err; echo "bench $1 rc=$?"
  python - "$1" <<PY
import json, sys
try:
    d=json.loads(open("gpurun_out/bench_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value",round(d["value"],2),"ms",round(d["ms_per_step"],2), {k:(round(v["ms_per_step"],2)) for k,v in d["roofline"]["families"].items()})
except Exception as e: print("no json", e)
PY
}
run ab_default OSB_ATTN_PP=0
run ab_pp OSB_ATTN_PP=1
run ab_default2 OSB_ATTN_PP=0
run ab_pp2 OSB_ATTN_PP=1
