mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r02f_gpu.log 2>&1; echo "gpu suite rc=$?"
tail -15 gpurun_out/r02f_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r02f_bench.err
timeout 900 python bench.py --workload nerfpp --steps 20 --warmup 3 > gpurun_out/r02f_bench_nerfpp.json 2> gpurun_out/r02f_bench_nerfpp.err; echo "bench nerfpp rc=$?"; tail -c 1200 gpurun_out/r02f_bench_nerfpp.err
python -c "
import json
d=json.load(open('gpurun_out/r02f_bench.json')); print('lego', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity_mode']['value'], d['cpu_baseline'], d['reference_gpu'])
d=json.load(open('gpurun_out/r02f_bench_nerfpp.json')); print('nerfpp', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['parity_mode'])
"
bash tools/gpu_profile_r02.sh
