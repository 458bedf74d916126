mkdir -p gpurun_out
for g in 1 0; do
FLNERF_GRAPH_DP=$g timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2956$g bench.py --gpus 8 --steps 20 --warmup 3 --no_kernel_table --no_cpu_baseline --no_parity_leg > gpurun_out/r02v_bench_8gpu_dp$g.json 2> gpurun_out/r02v_bench_8gpu_dp$g.err
echo "dp_graph=$g rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/r02v_bench_8gpu_dp$g.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['launch_mode'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
