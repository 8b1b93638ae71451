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
