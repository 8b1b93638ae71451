#!/bin/bash
# Round-1l GPU pass: all parity tests, bench line, ncu launch list, full captures (ed, ln single-warp, ln multi-warp, msa, cluster edges).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1l_pytest.log
tail -4 gpurun_out/r1l_pytest.log
timeout 600 python bench.py > gpurun_out/r1l_bench.json 2> gpurun_out/r1l_bench.err; echo "bench rc=$?" >> gpurun_out/r1l_bench.err; tail -3 gpurun_out/r1l_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1l_launches.csv python tools/prof_run.py all > gpurun_out/r1l_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ln_kernel -c 3 -f -o gpurun_out/r1l_ln python tools/prof_run.py k3 > gpurun_out/r1l_ncu_ln.log 2>&1
K5N=1184 timeout 400 ncu --set full --clock-control none --import-source on -k regex:ln_kernel -c 6 -f -o gpurun_out/r1l_ln5 python tools/prof_run.py k5 > gpurun_out/r1l_ncu_ln5.log 2>&1
python -c "
import json; j=json.load(open('gpurun_out/r1l_bench.json')); print(j['value'], j['e2e']['value'], j['ms_per_step'])
for f in j['families']: print(f.get('family','')[:40], f.get('value'), f.get('gcups'), f.get('kernel_ms'), (f.get('cpu_baseline') or {}).get('value'), f.get('error'))"
