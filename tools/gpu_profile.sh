#!/bin/bash
# On the GPU box: bench line, ncu launch list, ncu --set full of the three tensor-core kernels (fine pass).
mkdir -p gpurun_out
echo "=== bench N=1"
timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench.json
tail -5 gpurun_out/bench.err
echo "=== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no_cpu_baseline --images 4 > gpurun_out/ncu_list.log 2>&1
tail -3 gpurun_out/ncu_list.log
echo "=== ncu full (fwd/dgrad/wgrad of the fine pass)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_ -s 7 -c 3 -f -o gpurun_out/prof \
    python bench.py --steps 1 --warmup 1 --no_cpu_baseline --images 4 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
