#!/bin/bash
# warp-uniform tcgen05 issue in the GEMM / flash / resident kernels: full suite, per-shape GEMM rates, step A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/ab2_pytest.log 2>&1; rc=$?; echo "pytest -m gpu rc=$rc"
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/ab2_pytest.log | tail -n 8
[ $rc -ne 0 ] && exit 0
timeout 200 python tests/gemm_tune.py > gpurun_out/gemm_tune_v5.log 2>&1; grep -E "cuBLAS|cta2 bn256|cta1 bn192|cta2 bn192" gpurun_out/gemm_tune_v5.log
timeout 90 python tests/attn_prof.py 0 > gpurun_out/attn_prof_v5.log 2>&1; grep "attn impl" gpurun_out/attn_prof_v5.log
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
run ab2_default OSB_ATTN_PP=0
run ab2_pp OSB_ATTN_PP=1
