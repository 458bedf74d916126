#!/bin/bash
# Runs on the GPU box (under gpurun): every GPU test node in its own process (a trapped kernel kills only that
# process), then smoke().  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
NODES=$(python -m pytest tests -m gpu --collect-only -q 2>/dev/null | grep "::")
: > gpurun_out/tests.log
for n in $NODES; do
  echo "=== $n" >> gpurun_out/tests.log
  timeout 300 python -m pytest "$n" -x -q -m gpu 2>&1 | tail -40 >> gpurun_out/tests.log
done
grep -E "^===|passed|failed|error|Error|timeout|assert" gpurun_out/tests.log | head -150
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -15 | tee gpurun_out/smoke.log
