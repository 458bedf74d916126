mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02i_gpu.log 2>&1; echo "gpu suite rc=$?"; tail -6 gpurun_out/r02i_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02i_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02i_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r02i_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r02i_bench.json')); print('lego', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity_mode']['value'], d['parity_mode']['ms_per_step'], d['reference_gpu'] and d['reference_gpu'].get('value'), d['cpu_baseline'] and d['cpu_baseline'].get('value'))
print({k:(round(v['ms'],3), round(v['tensor_frac_burst'],3)) for k,v in d['roofline']['kernels'].items()}, d['roofline']['step']['frac'])
print({k:(round(v['exceeds_l2']['gbs']), round(v['step_size_l2_flushed']['us'],1)) for k,v in d['hbm_kernels'].items()})
"
sed -i 's/r02f/r02i/g' tools/gpu_profile_r02.sh
bash tools/gpu_profile_r02.sh
