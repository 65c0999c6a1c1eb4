#!/bin/bash
# Evidence run: launch list of the bench command + ncu --set full of the four GEMM shapes and the attention kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_final.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-vae > gpurun_out/launches_final.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -c 12 -o gpurun_out/prof_gemm_final -f \
  python tests/gemm_prof.py > gpurun_out/ncu_gemm_final.log 2>&1; echo "ncu gemm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 6 -c 1 -o gpurun_out/prof_attn_final_spatial -f \
  python tests/attn_prof.py 0 > gpurun_out/ncu_attn_final.log 2>&1; echo "ncu attn spatial rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 19 -c 1 -o gpurun_out/prof_attn_final_temporal -f \
  python tests/attn_prof.py 0 >> gpurun_out/ncu_attn_final.log 2>&1; echo "ncu attn temporal rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 32 -c 1 -o gpurun_out/prof_attn_final_cross -f \
  python tests/attn_prof.py 0 >> gpurun_out/ncu_attn_final.log 2>&1; echo "ncu attn cross rc=$?"
ls -la gpurun_out/*.ncu-rep
