#!/bin/bash
# build_variant.sh NAME [extra nvcc flags...]  ->  build/variants/libj3dg_NAME.so  (kernel tuning experiments)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O3,-fvisibility=hidden \
  --expt-relaxed-constexpr -Xptxas -v -shared -I include -I j3d_b200/csrc "$@" \
  -o build/variants/libj3dg_$name.so j3d_b200/csrc/*.cu 2> build/variants/$name.ptxas.log
grep -A2 "cast_kernelILb0" build/variants/$name.ptxas.log | grep -E "Used|spill" | sed "s/^/$name: /"
