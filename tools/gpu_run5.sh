mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_eval.py tests/test_gpu_driver.py tests/test_gpu_nerfpp.py -q -s > gpurun_out/r02e_tests.log 2>&1; echo "tests rc=$?"
grep -E "nerf\+\+|passed|failed|Error|error" gpurun_out/r02e_tests.log | head -40
PC_RES=200 PC_VIEWS=40 PC_ITERS=4000 PC_NRAND=1024 PC_SEEDS=$(seq -s, 10 57) PC_REF_SEEDS=0 PC_ARMS=bf16x3,bf16 timeout 2400 python tools/psnr_check.py > gpurun_out/r02e_psnr_48.json 2> gpurun_out/r02e_psnr_48.err; echo "psnr rc=$?"
tail -3 gpurun_out/r02e_psnr_48.err
python -c "
import json; d=json.load(open('gpurun_out/r02e_psnr_48.json')); print(json.dumps(d['delta_db']['test'],indent=0))"
