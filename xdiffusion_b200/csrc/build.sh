#!/bin/bash
# Build libxdb200.so (sm_100a only).  nvcc cross-compiles without a GPU.
set -e
cd "$(dirname "$0")"
OUT=../libxdb200.so
NVCC=${NVCC:-nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC ${NVCC_EXTRA:-}"
mkdir -p ../../build/obj
pids=()
for f in abi gemm_tc dit_block simt norm attention attention_tc elementwise step; do
  $NVCC $FLAGS -c $f.cu -o ../../build/obj/$f.o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o $OUT ../../build/obj/*.o
echo "built $(realpath $OUT)"
