#!/bin/bash
# On the GPU box: role-level cycle accounting of the forward / dgrad kernels (FLNERF_TC_PROF) next to their
# CUDA-event timings, with and without the activation stores / mask generation.
mkdir -p gpurun_out
{
echo "=== timings (no instrumentation)"
KB_TAG=base timeout 300 python tools/kernel_bench.py
FLNERF_FWD_DBG=2 KB_TAG=fwd_nostore timeout 300 python tools/kernel_bench.py
FLNERF_FWD_DBG=3 KB_TAG=fwd_nostore_nomask timeout 300 python tools/kernel_bench.py
echo "=== role accounting"
for dbg in 0 2 3; do
  echo "--- FLNERF_FWD_DBG=$dbg"
  FLNERF_FWD_DBG=$dbg FLNERF_TC_PROF=1 timeout 300 python tools/kernel_bench.py 2>&1 | grep tcprof | awk 'NR%13==5'
done
} 2>&1 | tee gpurun_out/tc_prof.log
