mkdir -p gpurun_out
nproc > gpurun_out/r02a_host.txt; free -g >> gpurun_out/r02a_host.txt; cat /sys/fs/cgroup/memory.max >> gpurun_out/r02a_host.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_x3.py -x -q -s > gpurun_out/r02a_x3.log 2>&1; echo "x3 rc=$?" 
tail -30 gpurun_out/r02a_x3.log
timeout 600 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_x3.py > gpurun_out/r02a_gpu.log 2>&1; echo "gpu rc=$?"
tail -5 gpurun_out/r02a_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r02a_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r02a_bench.json'))
print(d['value'],d['ms_per_step'],d['e2e']['value'])
print(json.dumps(d['roofline'],indent=0)[:3000])
print(json.dumps(d['parity_mode'])[:3000])
print(d['cpu_baseline'],d['reference_gpu'],d['epoch_ops'])
"
