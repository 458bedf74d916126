#!/bin/bash
# On the GPU box: ncu --set full (with source-level stall sampling) of one forward (training) and one dgrad launch.
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mlp_(fwd|dgrad)_tc" -s 2 -c 2 -f -o gpurun_out/tc3 \
    python tools/kernel_bench.py > gpurun_out/ncu_tc3.log 2>&1
tail -3 gpurun_out/ncu_tc3.log
ls -la gpurun_out/*.ncu-rep
