#!/bin/bash
# Runs the GPU checks in separate processes (a device-side trap poisons the CUDA context, so one
# failing group must not take the others with it).  Logs land in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.used --format=csv > gpurun_out/smi.txt 2>&1
for c in ln gemm1 gemm2 attn; do
  echo "=== debug $c" | tee -a gpurun_out/debug.log
  timeout 180 python tests/gpu_debug.py $c >> gpurun_out/debug.log 2>&1
  echo "rc=$?" | tee -a gpurun_out/debug.log
done
for k in ln_modulate "gemm and cta1" "gemm and cta2" "gemm and not cta1 and not cta2" attn errors; do
  echo "=== pytest -k '$k'" | tee -a gpurun_out/pytest.log
  timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -s -k "$k" >> gpurun_out/pytest.log 2>&1
  echo "rc=$?" | tee -a gpurun_out/pytest.log
done
tail -n 60 gpurun_out/debug.log
grep -E "passed|failed|error|rc=|===" gpurun_out/pytest.log | tail -n 40
