#!/bin/bash
# step-time A/B of the default dispatch against the ping-pong kernel on two-block key sets (30 timed steps, interleaved)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "gemm" > gpurun_out/ab3_pytest.log 2>&1; echo "pytest gemm rc=$?"
grep -E "passed|failed" gpurun_out/ab3_pytest.log | tail -n 2
run() {
  env $2 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-vae $3 > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err; echo "bench $1 rc=$?"
  python - "$1" <<PY
import json, sys
try:
    d=json.loads(open("gpurun_out/bench_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value",round(d["value"],2),"ms",round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],2), "clk", d["clocks"]["sm_mhz"], {k:(round(v["ms_per_step"],2)) for k,v in d["roofline"]["families"].items()})
except Exception as e: print("no json", e)
PY
}
run ab3_default OSB_ATTN_PP=0 ""
run ab3_pp OSB_ATTN_PP=1 ""
run ab3_default_b OSB_ATTN_PP=0 ""
run ab3_pp_b OSB_ATTN_PP=1 ""
run ab3_pp_graph OSB_ATTN_PP=1 "--graph"
run ab3_default_graph OSB_ATTN_PP=0 "--graph"
