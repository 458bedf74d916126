#!/bin/bash
# one node, 8 ranks: the lego line (4096 rays per rank) and the nerf++ line (1024 rays per rank = N_rand 8192, BASELINE configs[4])
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 50 --warmup 5 --no_kernel_table --no_cpu_baseline > gpurun_out/r02v_bench_8gpu.json 2> gpurun_out/r02v_bench_8gpu.err
echo "lego rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/r02v_bench_8gpu.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['parity_mode'] and d['parity_mode']['value'], d['clocks'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --steps 50 --warmup 5 --workload nerfpp > gpurun_out/r02v_bench_nerfpp_8gpu.json 2> gpurun_out/r02v_bench_nerfpp_8gpu.err
echo "nerfpp rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/r02v_bench_nerfpp_8gpu.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['parity_mode'] and d['parity_mode']['value'])"
