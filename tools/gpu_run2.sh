mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_graph.py -x -q -s > gpurun_out/r02b_graph.log 2>&1; echo "graph rc=$?"
tail -25 gpurun_out/r02b_graph.log
timeout 600 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_graph.py > gpurun_out/r02b_gpu.log 2>&1; echo "gpu rc=$?"
tail -5 gpurun_out/r02b_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 --no_cpu_baseline > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r02b_bench.err
timeout 900 python bench.py --steps 20 --warmup 3 --no_cpu_baseline --no_graph --no_parity_leg --no_kernel_table > gpurun_out/r02b_bench_nograph.json 2>> gpurun_out/r02b_bench.err; echo "bench rc=$?"
python -c "
import json
for f in ('r02b_bench','r02b_bench_nograph'):
    d=json.load(open('gpurun_out/%s.json'%f))
    print(f, d['value'],d['ms_per_step'],d['e2e']['value'],d['gpu_launches'],d.get('launch_mode'), d['roofline']['step']['frac'] if d['roofline'] else None, (d['parity_mode'] or {}).get('value'))
"
