#!/bin/bash
# Builds libdelly_b200.so (sm_100a only) in-tree. Used by __graft_entry__.build().
# Each .cu is compiled to an object in parallel (build/ is git-ignored), then linked.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
mkdir -p build
pids=()
objs=()
for src in delly_b200/csrc/*.cu; do
  obj=build/$(basename "${src%.cu}").o
  objs+=("$obj")
  # rebuild when the source or any shared header is newer than the object
  if [ ! -f "$obj" ] || [ "$src" -nt "$obj" ] || [ -n "$(find delly_b200/csrc include -name '*.cuh' -newer "$obj" -o -name '*.h' -newer "$obj")" ] || [ -n "${DGPU_EXTRA_FLAGS}" ]; then
    $NVCC -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a \
      -Xcompiler -fPIC -Xcompiler -Wall ${DGPU_EXTRA_FLAGS} -c -o "$obj" "$src" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait "$p"; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o delly_b200/libdelly_b200.so "${objs[@]}" -lcudart
echo "built delly_b200/libdelly_b200.so"
# C++ host mirror of the reference interface (+ flat test hooks), linked against the CUDA library
g++ -std=c++17 -O2 -fPIC -shared -Wall -Wno-sign-compare -o delly_b200/libdelly_b200_host.so delly_b200/host/capi.cpp \
  -Ldelly_b200 -l:libdelly_b200.so -Wl,-rpath,'$ORIGIN'
echo "built delly_b200/libdelly_b200_host.so"
