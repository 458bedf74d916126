mkdir -p gpurun_out
for sc in weak strong; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 200 --warmup 5 --scaling $sc --no_kernel_table --no_parity_leg --no_cpu_baseline > gpurun_out/r02r_bench_2gpu_${sc}_200.json 2> gpurun_out/r02r_bench_2gpu_${sc}.err
echo "$sc rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/r02r_bench_2gpu_${sc}_200.json').read().strip().splitlines()[-1]); print(d['scaling'], d['value'], d['ms_per_step'], d['e2e']['value'])"
done
timeout 300 python bench.py --steps 200 --warmup 5 --no_kernel_table --no_parity_leg --no_cpu_baseline > gpurun_out/r02r_bench_1gpu_200.json 2>/dev/null; python -c "
import json
d=json.loads(open('gpurun_out/r02r_bench_1gpu_200.json').read().strip().splitlines()[-1]); print('1gpu', d['value'], d['ms_per_step'], d['e2e']['value'])"
timeout 300 python bench.py --steps 200 --warmup 5 --no_graph --no_kernel_table --no_parity_leg --no_cpu_baseline > gpurun_out/r02r_bench_1gpu_200_nograph.json 2>/dev/null; python -c "
import json
d=json.loads(open('gpurun_out/r02r_bench_1gpu_200_nograph.json').read().strip().splitlines()[-1]); print('1gpu nograph', d['value'], d['ms_per_step'], d['e2e']['value'])"
