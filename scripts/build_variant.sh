#!/bin/bash
# build gpurun_scratch/libmvosr_<tag>.so from the working tree with extra nvcc flags:  scripts/build_variant.sh <tag> [-DFLAG ...]
set -e
tag=$1; shift
cd "$(dirname "$0")/../mvoscalerecovery_b200/csrc"
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
nvcc $F "$@" -c -o /tmp/api_$tag.o api.cu 2>&1 | grep -E "error|spill" || true
[ -f /tmp/fp5_api.o ] && [ /tmp/fp5_api.o -nt five_point_api.cu ] && [ /tmp/fp5_api.o -nt five_point.cuh ] && [ /tmp/fp5_api.o -nt five_point_kernel.cuh ] || nvcc $F -fmad=false -c -o /tmp/fp5_api.o five_point_api.cu
mkdir -p ../../gpurun_scratch
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../gpurun_scratch/libmvosr_$tag.so /tmp/api_$tag.o /tmp/fp5_api.o
echo built $tag
