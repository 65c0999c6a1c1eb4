#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for k in group_stats vae_prep causal_conv3d residual goldens roundtrip; do
  echo "=== pytest -k $k" | tee -a gpurun_out/vae_pytest.log
  timeout 600 python -m pytest tests/test_vae_gpu.py -m gpu -q -x -s -k "$k" >> gpurun_out/vae_pytest.log 2>&1
  echo "rc=$?" | tee -a gpurun_out/vae_pytest.log
done
grep -E "parity|passed|failed|Error|error:|rc=|===" gpurun_out/vae_pytest.log | tail -n 60
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x 2>&1 | tail -n 3
