#!/bin/bash
# bash tools/build_variant.sh <name> <nvcc extra flags...>  ->  build_variants/libxm_<name>.so (A/B builds for tools/gpu_ab.sh)
set -e
NAME=$1; shift
mkdir -p build_variants
cd x-maps_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -rdc=true -Xcompiler -fPIC -shared "$@" -o ../../build_variants/libxm_$NAME.so xm_capi.cu -lcudadevrt
echo built build_variants/libxm_$NAME.so
