#!/bin/bash
# Round-2 profiles (B200_PROFILING.md recipe): per-launch device time + DRAM bytes + tensor-pipe share of ONE training step per
# precision, and an `ncu --set full` capture of the MLP kernels.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread
timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02f_step_metrics.csv python tools/step_once.py bf16,bf16x3 > gpurun_out/r02f_step_metrics.log 2>&1; echo "metrics rc=$?"
timeout 1200 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:mlp_ -o gpurun_out/r02f_prof -f python tools/step_once.py bf16,bf16x3 > gpurun_out/r02f_prof.log 2>&1; echo "full rc=$?"
ls -la gpurun_out | tail -5
