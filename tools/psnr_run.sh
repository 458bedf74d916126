#!/bin/bash
# On the GPU box: encode test + PSNR-at-equal-iterations check (bf16 tensor-core path vs fp32 parity path, 3 seeds).
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_kernels.py::test_posenc_and_encode -x -q -m gpu 2>&1 | tail -3
PC_ITERS=${PC_ITERS:-600} PC_NRAND=${PC_NRAND:-512} PC_RES=${PC_RES:-80} PC_VIEWS=${PC_VIEWS:-30} timeout ${PC_TIMEOUT:-420} python tools/psnr_check.py 2> gpurun_out/psnr.err | tee gpurun_out/psnr.json
tail -3 gpurun_out/psnr.err
