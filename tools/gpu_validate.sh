#!/bin/bash
# Round-end validation on one B200 (run through gpurun): the full -m gpu suite, smoke(), the lego and nerf++ bench lines.
# Usage: bash tools/gpu_validate.sh <tag>      -> gpurun_out/<tag>_{gpu.log,smoke.log,bench.json,bench_nerfpp.json}
TAG=${1:-val}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/${TAG}_gpu.log 2>&1; echo "gpu suite rc=$?"; tail -6 gpurun_out/${TAG}_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.load(open('gpurun_out/${TAG}_bench.json'))
print('lego', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'x3', d['parity_mode']['value'], d['parity_mode']['ms_per_step'],
      'ref gpu', d['reference_gpu'] and d['reference_gpu'].get('value'), 'ref cpu', d['cpu_baseline'] and d['cpu_baseline'].get('value'))
print({k: (round(v['ms'], 3), round(v['tensor_frac_burst'], 3)) for k, v in d['roofline']['kernels'].items()}, d['roofline']['step']['frac'])
PY
timeout 600 python bench.py --workload nerfpp > gpurun_out/${TAG}_bench_nerfpp.json 2> gpurun_out/${TAG}_bench_nerfpp.err; echo "nerfpp rc=$?"
python -c "import json; d=json.load(open('gpurun_out/${TAG}_bench_nerfpp.json')); print('nerfpp', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity_mode']['value'])"
