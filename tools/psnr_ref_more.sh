#!/bin/bash
# ten more paired seeds of the deterministic-depth protocol: bf16x3 vs the UNMODIFIED reference (profiles/r02d_*: seeds 0-2)
mkdir -p gpurun_out
PC_RES=200 PC_VIEWS=40 PC_ITERS=4000 PC_NRAND=1024 PC_PERTURB=0 PC_ARMS=bf16x3 PC_SEEDS=3,4,5,6,7,8,9,10,11,12 PC_REF_SEEDS=10 \
  timeout 2300 python tools/psnr_check.py 2> gpurun_out/r02t_psnr_det_more.err > gpurun_out/r02t_psnr_det_more.json
echo "rc=$?"; tail -c 600 gpurun_out/r02t_psnr_det_more.json; tail -3 gpurun_out/r02t_psnr_det_more.err
