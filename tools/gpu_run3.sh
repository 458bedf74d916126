mkdir -p gpurun_out
PC_RES=200 PC_VIEWS=40 PC_ITERS=4000 PC_NRAND=1024 PC_SEEDS=0,1,2,3,4,5,6,7,8,9 PC_REF_SEEDS=2 timeout 2000 python tools/psnr_check.py > gpurun_out/r02c_psnr.json 2> gpurun_out/r02c_psnr.err; echo "psnr rc=$?"
grep partial gpurun_out/r02c_psnr.err; tail -3 gpurun_out/r02c_psnr.err
python -c "
import json; d=json.load(open('gpurun_out/r02c_psnr.json')); print(json.dumps(d['delta_db'],indent=1)); print(d['seconds'])"
