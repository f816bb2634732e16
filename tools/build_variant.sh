#!/bin/bash
# usage: tools/build_variant.sh <name> [-DNID_...=...]...   ->  build/libvar_<name>.so (same flags as the Makefile + the defines)
name=$1; shift
mkdir -p build
cd nid-pose-estimation_b200
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -ccbin /usr/bin/g++ \
  "$@" -shared -o ../build/libvar_$name.so csrc/nid_kernels.cu csrc/nid_sorted.cu csrc/nid_api.cu csrc/ref_shims.cu 2> ../build/var_$name.log \
  && echo "built build/libvar_$name.so" || (tail -20 ../build/var_$name.log; exit 1)
