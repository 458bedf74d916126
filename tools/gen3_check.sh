#!/bin/bash
# On the GPU box: bf16 MLP parity tests on the default (gen-3, CTA pair) kernels, then kernel timings + role accounting.
mkdir -p gpurun_out
{
for n in test_mlp_bf16_forward_backward_vs_emulation test_training_reduces_loss_bf16 test_full_size_properties; do
  echo "=== $n"
  timeout 200 python -m pytest "tests/test_gpu_mlp.py::$n" -x -q -m gpu 2>&1 | tail -15
done
echo "=== timings"
KB_TAG=gen3 timeout 200 python tools/kernel_bench.py 2>&1 | tail -3
FLNERF_TC_GEN=2 KB_TAG=gen2 timeout 200 python tools/kernel_bench.py 2>&1 | tail -3
echo "=== role accounting gen3"
FLNERF_TC_PROF=1 timeout 200 python tools/kernel_bench.py 2>&1 | grep tcprof | awk 'NR%13==5'
} 2>&1 | tee gpurun_out/gen3_check.log
