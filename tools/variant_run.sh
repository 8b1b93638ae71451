#!/bin/bash
# Development aid: time a command with alternative builds of the CUDA library (variants/*.so) swapped in.
# usage: tools/variant_run.sh 'command'   (run on the GPU box)
cp delly_b200/libdelly_b200.so /tmp/lib_base.so
for v in variants/*.so; do
  echo "== $v"; cp "$v" delly_b200/libdelly_b200.so; bash -c "$1"
done
cp /tmp/lib_base.so delly_b200/libdelly_b200.so
echo "== base"; bash -c "$1"
