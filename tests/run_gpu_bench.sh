#!/bin/bash
# Model parity + bench + ncu evidence, one gpurun call.  Logs land in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== model parity" | tee gpurun_out/model.log
timeout 900 python -m pytest tests/test_stdit3_gpu.py -m gpu -q -x -s >> gpurun_out/model.log 2>&1
echo "rc=$?" | tee -a gpurun_out/model.log
grep -E "parity|passed|failed|Error|error" gpurun_out/model.log | tail -n 30
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 5
echo "=== bench"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "rc=$?"; tail -n 5 gpurun_out/bench.err; cat gpurun_out/bench.json
if [ "$1" == "ncu" ]; then
  echo "=== ncu launch list"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 700 --csv \
      --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  echo "rc=$?"; tail -n 3 gpurun_out/launches.csv
  echo "=== ncu full (gemm)"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 40 -c 3 \
      -o gpurun_out/prof_gemm -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
  echo "rc=$?"; tail -n 3 gpurun_out/ncu_full.log
fi
