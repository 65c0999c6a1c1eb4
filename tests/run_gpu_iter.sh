#!/bin/bash
# iteration run: regression tests, GEMM sweep, attention timings (+ optional ncu of attention), bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x > gpurun_out/iter_pytest.log 2>&1; echo "kernel tests rc=$?"; tail -n 3 gpurun_out/iter_pytest.log
timeout 300 python tests/gemm_tune.py 2>&1 | tee gpurun_out/gemm_tune.log
timeout 300 python tests/attn_prof.py 2>&1 | tee gpurun_out/attn_prof.log
if [ "$1" == "ncu" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_short -s 9 -c 1 -o gpurun_out/prof_attn_spatial -f python tests/attn_prof.py > gpurun_out/ncu_attn.log 2>&1; echo "ncu rc=$?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_short -s 35 -c 1 -o gpurun_out/prof_attn_cross -f python tests/attn_prof.py >> gpurun_out/ncu_attn.log 2>&1; echo "ncu rc=$?"
fi
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err; echo "bench rc=$?"; cat gpurun_out/bench_iter.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value']); print(json.dumps(d['roofline']['families']))"
