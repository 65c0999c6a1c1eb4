#!/bin/bash
# multi-GPU: sequence-parallel parity check + bench under torchrun.  usage: run_gpu_multi.sh N
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "=== sp parity check ($N GPUs)"
timeout 600 $TR --master-port 29511 tests/sp_gpu_check.py 2>&1 | grep -E "sp$N|SP_CHECK|Error|error|Warning" | tee gpurun_out/r02_sp_check_$N.log | tail -n 24
if [ -n "${VAE_TP:-}" ]; then
  echo "=== frame-sharded VAE decode check ($N GPUs)"
  timeout 600 $TR --master-port 29513 tests/vae_tp_gpu_check.py 2>&1 | grep -E "vae-tp|VAE_TP|Error|error" | tee gpurun_out/vae_tp_check_$N.log | tail -n 20
fi
if [ -n "${MMDIT_SP:-}" ]; then
  echo "=== MMDiT Ulysses sequence-parallel check ($N GPUs)"
  timeout 600 $TR --master-port 29514 tests/mmdit_sp_gpu_check.py 2>&1 | grep -E "mmdit-sp|MMDIT_SP|Error|error" | tee gpurun_out/mmdit_sp_check_$N.log | tail -n 20
fi
for mode in ${MODES:-sp}; do
  echo "=== bench --parallel $mode ($N GPUs)"
  timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --parallel $mode ${BENCH_FLAGS:-} > gpurun_out/bench_${mode}_$N.json 2> gpurun_out/bench_${mode}_$N.err
  echo "rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${mode}_$N.json").read().strip().splitlines()[-1])
    print("${mode}$N value", round(d["value"], 2), "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2), d["scaling"], d.get("sp_check"), d.get("dp_replicas"), d["config"].get("exchange"))
except Exception as e:
    print("no json", e); print(open("gpurun_out/bench_${mode}_$N.err").read()[-2000:])
PY
done
