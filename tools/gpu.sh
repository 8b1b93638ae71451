#!/bin/bash
# Rebuild the in-tree .so files, then run a command on a B200 through gpurun.
# usage: tools/gpu.sh [--timeout S] [--gpus N] -- 'command'
set -e
cd "$(dirname "$0")/.."
./build.sh >/dev/null
make -C oracle -s all >/dev/null 2>&1 || true
exec /usr/local/graft/bin/gpurun "$@"
