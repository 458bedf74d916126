#!/bin/bash
# Builds libflnerf.so (sm_100a only) in-tree next to the Python binding.  nvcc cross-compiles without a GPU.
set -e
cd "$(dirname "$0")"
OUT=../flnerf_b200/libflnerf.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall"
mkdir -p build
pids=()
for f in api rays composite train_ops nerfpp eval mlp_simt mlp_tc; do
  if [ ! -f build/$f.o ] || [ $f.cu -nt build/$f.o ] || [ -n "$(find . -maxdepth 1 \( -name '*.cuh' -o -name '*.h' \) -newer build/$f.o 2>/dev/null)" ] \
     || [ ../../include/flnerf.h -nt build/$f.o ]; then
    $NVCC $FLAGS ${FL_PTXAS_V:+-Xptxas -v} -c $f.cu -o build/$f.o &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o $OUT build/api.o build/rays.o build/composite.o build/train_ops.o build/nerfpp.o build/eval.o build/mlp_simt.o build/mlp_tc.o -lcudart
echo "built $OUT"
