#!/bin/bash
# Round-1e GPU pass: parity tests, bench line (pipelined e2e), msa capture at bench scale.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1e_pytest.log
timeout 500 python bench.py > gpurun_out/r1e_bench.json 2> gpurun_out/r1e_bench.err; echo "bench rc=$?" >> gpurun_out/r1e_bench.err
K2N=16384 timeout 300 ncu --set full --clock-control none --import-source on -k regex:msa_kernel -c 1 -f -o gpurun_out/r1e_msa python tools/prof_run.py k2 > gpurun_out/r1e_ncu_msa.log 2>&1
tail -5 gpurun_out/r1e_pytest.log; python -c "
import json; j=json.load(open('gpurun_out/r1e_bench.json')); print(j['value'], j['e2e'], j['ms_per_step'])"
