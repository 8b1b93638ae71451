#!/bin/bash
# Round-1d GPU pass: parity tests, bench line, ncu launch list, full captures of the top kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1d_pytest.log
timeout 500 python bench.py > gpurun_out/r1d_bench.json 2> gpurun_out/r1d_bench.err; echo "bench rc=$?" >> gpurun_out/r1d_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1d_launches.csv python tools/prof_run.py all > gpurun_out/r1d_launches.log 2>&1
K2N=1024 timeout 300 ncu --set full --clock-control none --import-source on -k regex:msa_kernel -c 1 -f -o gpurun_out/r1d_msa python tools/prof_run.py k2 > gpurun_out/r1d_ncu_msa.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ln_kernel -c 3 -f -o gpurun_out/r1d_ln python tools/prof_run.py k3 > gpurun_out/r1d_ncu_ln.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ed_small_kernel -c 3 -f -o gpurun_out/r1d_ed python tools/prof_run.py k1 > gpurun_out/r1d_ncu_ed.log 2>&1
tail -3 gpurun_out/r1d_pytest.log; cat gpurun_out/r1d_bench.json | head -c 1500
