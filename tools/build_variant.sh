#!/bin/bash
# Development aid: build a variant of libdelly_b200.so with extra -D flags for ONE source file (default long_needle.cu) into variants/NAME/
# (git-ignored, travels with gpurun). Select it at run time with DGPU_LIB=variants/NAME/libdelly_b200.so.
# usage: tools/build_variant.sh NAME "-DFOO -DBAR=1" [source.cu]
set -e
cd "$(dirname "$0")/.."
name=$1; flags=$2; src=${3:-long_needle.cu}
mkdir -p variants/$name
/usr/local/cuda/bin/nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC $flags -c -o variants/$name/${src%.cu}.o delly_b200/csrc/$src
objs=()
for o in build/*.o; do
  if [ "$(basename $o)" == "${src%.cu}.o" ]; then objs+=("variants/$name/${src%.cu}.o"); else objs+=("$o"); fi
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/$name/libdelly_b200.so "${objs[@]}" -lcudart
echo "built variants/$name/libdelly_b200.so"
