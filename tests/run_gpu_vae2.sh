#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vae_gpu.py -m gpu -q -x -s > gpurun_out/vae_pytest2.log 2>&1; echo "vae tests rc=$?"
grep -E "parity|passed|failed|Error" gpurun_out/vae_pytest2.log | tail -n 30
echo "=== vae bench small"; timeout 600 python tests/vae_bench.py 17 256 256 2 2>&1 | tail -n 3
echo "=== vae bench 33x480x848"; timeout 900 python tests/vae_bench.py 33 480 848 1 2>&1 | tail -n 3
echo "=== vae bench full"; timeout 1200 python tests/vae_bench.py 65 720 1280 1 2>&1 | tail -n 5
