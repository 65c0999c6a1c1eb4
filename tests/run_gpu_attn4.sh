#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attn" > gpurun_out/attn4_pytest.log 2>&1; echo "attn tests rc=$?"
grep -E "passed|failed|Error|error|timed out|parity.*impl" gpurun_out/attn4_pytest.log | tail -n 14
timeout 300 python tests/attn_prof.py 1 2 3 2>&1 | tee gpurun_out/attn_prof4.log | tail -n 12
timeout 300 python -m pytest tests/test_vae_gpu.py -m gpu -q -x -k "roundtrip or causality" 2>&1 | tail -n 3
