#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for m in spatial temporal cross; do
  timeout 120 python tests/pp_small.py $m 4 2>&1 | grep -v "^$" | tail -n 6
  echo "$m rc=$?"
done
