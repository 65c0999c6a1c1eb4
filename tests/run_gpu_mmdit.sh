#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mmdit_gpu.py -m gpu -q -x -s > gpurun_out/mmdit_pytest.log 2>&1; echo "mmdit rc=$?"
grep -E "parity|passed|failed|Error|error" gpurun_out/mmdit_pytest.log | tail -n 20
timeout 600 python -m pytest tests/test_vae_gpu.py -m gpu -q -x -s -k "goldens or roundtrip or group_stats" > gpurun_out/vae_pytest3.log 2>&1; echo "vae rc=$?"
grep -E "parity|passed|failed|Error" gpurun_out/vae_pytest3.log | tail -n 12
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_stdit3_gpu.py -m gpu -q -x 2>&1 | tail -n 3
