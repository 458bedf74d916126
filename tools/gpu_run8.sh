mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02h_gpu.log 2>&1; echo "gpu suite rc=$?"; tail -8 gpurun_out/r02h_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02h_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02h_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r02h_bench.err
timeout 900 python bench.py --workload nerfpp --steps 20 --warmup 3 > gpurun_out/r02h_bench_nerfpp.json 2> gpurun_out/r02h_bench_nerfpp.err; echo "bench nerfpp rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r02h_bench.json')); print('lego', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity_mode']['value'], d['parity_mode']['ms_per_step'], d['reference_gpu'] and d['reference_gpu'].get('value'), d['cpu_baseline'] and d['cpu_baseline'].get('value'))
print({k:(round(v['ms'],3), round(v['tensor_frac_burst'],3)) for k,v in d['parity_mode']['kernels'].items()})
d=json.load(open('gpurun_out/r02h_bench_nerfpp.json')); print('nerfpp', d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity_mode']['value'])
"
PC_RES=200 PC_VIEWS=40 PC_ITERS=4000 PC_NRAND=1024 PC_SEEDS=$(seq -s, 58 105) PC_REF_SEEDS=0 PC_ARMS=bf16x3,bf16 timeout 2400 python tools/psnr_check.py > gpurun_out/r02h_psnr_48b.json 2> gpurun_out/r02h_psnr_48b.err; echo "psnr rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02h_psnr_48b.json')); print(json.dumps(d['delta_db']['test'],indent=0)); print({k:sum(v)/len(v) for k,v in d['seconds'].items()})"
