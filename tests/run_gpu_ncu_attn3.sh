#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_flash -s 9 -c 1 -o gpurun_out/prof_attn3_spatial -f python tests/attn_prof.py 3 > gpurun_out/ncu_attn3.log 2>&1; echo "ncu rc=$?"
timeout 300 python -m pytest tests/test_vae_gpu.py -m gpu -q -x -s -k roundtrip 2>&1 | tail -n 25
