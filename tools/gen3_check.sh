#!/bin/bash
# On the GPU box: bf16 MLP parity tests on the CTA-pair kernels, kernel timings + role accounting, optional ncu capture.
mkdir -p gpurun_out
{
for n in test_mlp_bf16_forward_backward_vs_emulation test_training_reduces_loss_bf16 test_full_size_properties; do
  echo "=== $n"
  timeout 200 python -m pytest "tests/test_gpu_mlp.py::$n" -x -q -m gpu --tb=short 2>&1 | grep -E "^E  |passed|failed" | head -5
done
echo "=== timings"
KB_TAG=pair timeout 200 python tools/kernel_bench.py 2>&1 | tail -3
echo "=== role accounting"
for dbg in 0 1 2 3; do
FLNERF_FWD_DBG=$dbg FLNERF_TC_PROF=1 timeout 200 python tools/kernel_bench.py 2>&1 | grep tcprof | awk 'NR%13==5' | sed "s/^/dbg$dbg /" | cut -c1-700
done
if [ -n "$NCU" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mlp_(fwd|dgrad)_tc" -s 1 -c 2 -f -o gpurun_out/tc4 \
      python tools/kernel_bench.py > gpurun_out/ncu_tc4.log 2>&1
  tail -2 gpurun_out/ncu_tc4.log
fi
} 2>&1 | tee gpurun_out/gen3_check.log
