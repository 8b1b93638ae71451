#!/bin/bash
# Round-1h: ncu full captures of the longNeedle kernels after the traceback / ring changes (K3 single-warp classes, K5 multi-warp CTA).
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ln_kernel -c 3 -f -o gpurun_out/r1h_ln python tools/prof_run.py k3 > gpurun_out/r1h_ncu_ln.log 2>&1
K5N=148 timeout 400 ncu --set full --clock-control none --import-source on -k regex:ln_kernel -c 1 -f -o gpurun_out/r1h_ln5 python tools/prof_run.py k5 > gpurun_out/r1h_ncu_ln5.log 2>&1
tail -3 gpurun_out/r1h_ncu_ln5.log
