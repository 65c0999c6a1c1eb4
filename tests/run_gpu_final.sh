#!/bin/bash
# round-end evidence: ncu launch list of exactly one step, then the default bench line (VAE leg + CPU baseline included)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_step.csv \
  python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-vae --profile-step > gpurun_out/launches_step.log 2>&1; echo "launch list rc=$?"
wc -l gpurun_out/launches_step.csv
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
tail -n 2 gpurun_out/bench_final.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
    print("final value",round(d["value"],2),"ms",round(d["ms_per_step"],2),"e2e",round(d["e2e"]["value"],2),"launches",d["gpu_launches"],"frac",round(d["roofline"]["frac"],3),"step_frac",round(d["roofline"]["step_frac_of_peak"],3))
    print("cpu", d["cpu_baseline"]); print("vae", {k:v for k,v in (d["vae"] or {}).items() if not isinstance(v,dict)})
except Exception as e: print("no json", e)
PY
