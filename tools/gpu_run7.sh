mkdir -p gpurun_out
# (1) two ranks: the step graph with NCCL inside, bench + the 2-rank tests
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02g_bench_2gpu.json 2> gpurun_out/r02g_bench_2gpu.err; echo "bench2 rc=$?"; tail -c 1500 gpurun_out/r02g_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --no_graph --no_parity_leg --no_kernel_table > gpurun_out/r02g_bench_2gpu_nograph.json 2>> gpurun_out/r02g_bench_2gpu.err; echo "bench2 nograph rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload nerfpp --steps 20 --warmup 3 > gpurun_out/r02g_bench_nerfpp_2gpu.json 2> gpurun_out/r02g_bench_nerfpp_2gpu.err; echo "bench2 nerfpp rc=$?"; tail -c 800 gpurun_out/r02g_bench_nerfpp_2gpu.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02g_ref_cpu.json 2> gpurun_out/r02g_ref_cpu.err; echo "ref rc=$?"; cat gpurun_out/r02g_ref_cpu.json | cut -c1-400
python -c "
import json
for f in ('r02g_bench_2gpu','r02g_bench_2gpu_nograph','r02g_bench_nerfpp_2gpu'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d.get('launch_mode'), (d.get('parity_mode') or {}).get('value'))
    except Exception as e: print(f, 'ERR', e)
"
# (2) tests touched since the last run
timeout 900 python -m pytest tests/test_gpu_eval.py tests/test_gpu_nerfpp.py tests/test_gpu_edge.py tests/test_gpu_kernels.py tests/test_gpu_driver.py -q > gpurun_out/r02g_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r02g_tests.log
# (3) experiment: weight-gradient terms of the split mode
for p in 1 2 3; do FLNERF_X3_WGRAD_PASSES=$p timeout 300 python tools/x3_wgrad_passes.py 2>&1 | tail -1; done | tee gpurun_out/r02g_wgrad_passes.log
