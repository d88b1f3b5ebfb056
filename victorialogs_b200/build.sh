#!/bin/bash
# Builds libvlscan.so (CUDA kernels + C ABI) for sm_100a. nvcc cross-compiles without a GPU.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xcudafe --diag_suppress=177 ${VL_NVCC_EXTRA}"
mkdir -p build
pids=()
for tu in vl_engine vl_gen vl_zstd; do
    $NVCC $FLAGS -c csrc/$tu.cu -o build/$tu.o &
    pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o libvlscan.so build/vl_engine.o build/vl_gen.o build/vl_zstd.o -ldl
echo built victorialogs_b200/libvlscan.so
