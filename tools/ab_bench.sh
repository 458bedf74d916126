#!/bin/bash
# On the GPU box: interleaved A/B timing of two builds of libflnerf.so (tools/ab/libA.so, libB.so) on the SAME box --
# box-to-box variance (+-5 %) is larger than most single optimisations.
LIB=fast-learning-nerf_b200/flnerf_b200/libflnerf.so
cp $LIB /tmp/lib_keep.so
for rep in 1 2; do
  for v in ${VARIANTS:-A B}; do
    cp tools/ab/lib$v.so $LIB
    KB_TAG=$v timeout 200 python tools/kernel_bench.py 2>&1 | tail -1
  done
done
cp /tmp/lib_keep.so $LIB
