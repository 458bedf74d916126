mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_eval.py tests/test_gpu_driver.py -x -q > gpurun_out/r02d_eval_driver.log 2>&1; echo "eval/driver rc=$?"
tail -40 gpurun_out/r02d_eval_driver.log
PC_PERTURB=0 PC_RES=200 PC_VIEWS=40 PC_ITERS=4000 PC_NRAND=1024 PC_SEEDS=0,1,2 PC_REF_SEEDS=3 PC_ARMS=bf16x3,x3p timeout 2000 python tools/psnr_check.py > gpurun_out/r02d_psnr_det.json 2> gpurun_out/r02d_psnr_det.err; echo "psnr rc=$?"
grep partial gpurun_out/r02d_psnr_det.err; tail -3 gpurun_out/r02d_psnr_det.err
python -c "
import json; d=json.load(open('gpurun_out/r02d_psnr_det.json')); print(json.dumps(d['delta_db']['test'],indent=0))"
