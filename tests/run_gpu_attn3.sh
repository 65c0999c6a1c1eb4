#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attn" > gpurun_out/attn3_pytest.log 2>&1; echo "attn tests rc=$?"
grep -E "passed|failed|Error|error|timed out" gpurun_out/attn3_pytest.log | tail -n 12
timeout 300 python tests/attn_prof.py 1 2 3 2>&1 | tee gpurun_out/attn_prof3.log | tail -n 12
timeout 300 python -m pytest tests/test_vae_gpu.py -m gpu -q -x -s -k roundtrip 2>&1 | grep -E "determinism|passed|failed" | tail -n 4
for impl in 2 3; do
OSB_ATTN_IMPL=$impl timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-vae > gpurun_out/bench_attn$impl.json 2> gpurun_out/bench_attn$impl.err; echo "bench impl$impl rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_attn$impl.json").read().strip().splitlines()[-1]); print("impl$impl value",round(d["value"],2),"ms",round(d["ms_per_step"],2)); print(json.dumps(d["roofline"]["families"]))
except Exception as e: print("no json", e)
PY
done
