#!/bin/bash
# ping-pong attention kernel: parity, per-shape timing, whole-step bench with it selected
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attn" > gpurun_out/pp_pytest.log 2>&1; echo "pytest attn rc=$?"
grep -E "passed|failed|^FAILED|^ERROR|rel_l2|timed out" gpurun_out/pp_pytest.log | tail -n 30
timeout 300 python tests/attn_prof.py 0 4 2>&1 | tee gpurun_out/attn_prof_pp.log | tail -n 8
OSB_ATTN_IMPL=4 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-vae > gpurun_out/bench_pp.json 2> gpurun_out/bench_pp.err; echo "bench pp rc=$?"
tail -n 3 gpurun_out/bench_pp.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_pp.json").read().strip().splitlines()[-1])
    print("pp value",round(d["value"],2),"ms",round(d["ms_per_step"],2),"e2e",round(d["e2e"]["value"],2), d["roofline"]["families"])
except Exception as e: print("no json", e)
PY
OSB_ATTN_IMPL=4 timeout 600 python -m pytest tests/test_stdit3_gpu.py -m gpu -q 2>&1 | tail -n 5
