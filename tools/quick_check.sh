#!/bin/bash
# On the GPU box: fast regression loop for the tensor-core MLP kernels (3 parity tests, kernel timings, short bench).
mkdir -p gpurun_out
{
for n in test_mlp_bf16_forward_backward_vs_emulation test_full_size_properties; do
  timeout 200 python -m pytest "tests/test_gpu_mlp.py::$n" -x -q -m gpu --tb=short 2>&1 | grep -E "^E  |passed|failed" | head -5
done
KB_TAG=pair timeout 200 python tools/kernel_bench.py 2>&1 | tail -2
[ -x tools/ub/tmem_bw ] && [ -n "$UB" ] && timeout 120 tools/ub/tmem_bw
[ -n "$BENCH" ] && timeout 600 python bench.py --steps 20 --warmup 3 --no_cpu_baseline 2>/dev/null | tail -1 | cut -c1-400
} 2>&1 | tee gpurun_out/quick_check.log
