#!/bin/bash
# Builds libdelly_b200.so (sm_100a only) in-tree. Used by __graft_entry__.build().
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
SRCS=$(ls delly_b200/csrc/*.cu)
$NVCC -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a \
  -Xcompiler -fPIC -Xcompiler -Wall -shared ${DGPU_EXTRA_FLAGS} \
  -o delly_b200/libdelly_b200.so $SRCS -lcudart
echo "built delly_b200/libdelly_b200.so"
# C++ host mirror of the reference interface (+ flat test hooks), linked against the CUDA library
g++ -std=c++17 -O2 -fPIC -shared -Wall -Wno-sign-compare -o delly_b200/libdelly_b200_host.so delly_b200/host/capi.cpp \
  -Ldelly_b200 -l:libdelly_b200.so -Wl,-rpath,'$ORIGIN'
echo "built delly_b200/libdelly_b200_host.so"
