cd /root/repo
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k attn 2>&1 | tail -n 3
for n in 9:spatial 22:temporal 35:cross; do s=${n%%:*}; name=${n##*:};
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_short -s $s -c 1 -o gpurun_out/prof_attn2_$name -f python tests/attn_prof.py > gpurun_out/ncu_attn2_$name.log 2>&1; echo "ncu $name rc=$?"; done
