#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_vae_gpu.py tests/test_sampling_gpu.py -m gpu -q > gpurun_out/last_pytest.log 2>&1; echo "pytest vae+sampling rc=$?"
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/last_pytest.log | tail -n 4
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -n 3
