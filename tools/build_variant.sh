#!/bin/bash
# build/libvar_<name>.so with extra nvcc flags (compile-time knobs): tools/build_variant.sh <name> -DNID_X=1 ...
name=$1; shift
cd "$(dirname "$0")/../nid-pose-estimation_b200" && mkdir -p ../build && \
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -ccbin /usr/bin/g++ \
  "$@" -shared -o ../build/libvar_$name.so csrc/nid_kernels.cu csrc/nid_sorted.cu csrc/nid_api.cu csrc/ref_shims.cu
