#!/bin/bash
# A/B builds of libbz2b200.so with extra -D flags:  tools/ab_build.sh NAME "-DBZ_SCAP=16 ..."  -> gpurun_ab/libbz2b200_NAME.so
# (run a variant with BZ2B200_LIB=gpurun_ab/libbz2b200_NAME.so python bench.py ...)
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_ab
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xcompiler -O3 --shared $2 \
    -o gpurun_ab/libbz2b200_$1.so bzip2_rust_b200/csrc/*.cu
