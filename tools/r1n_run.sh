#!/bin/bash
# Round-1n (final) GPU pass: all parity tests, smoke(), bench line, ncu launch list.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1n_pytest.log
tail -4 gpurun_out/r1n_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1n_smoke.log 2>&1; tail -2 gpurun_out/r1n_smoke.log
timeout 600 python bench.py > gpurun_out/r1n_bench.json 2> gpurun_out/r1n_bench.err; echo "bench rc=$?" >> gpurun_out/r1n_bench.err; tail -2 gpurun_out/r1n_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1n_launches.csv python tools/prof_run.py all > gpurun_out/r1n_launches.log 2>&1
python -c "
import json; j=json.load(open('gpurun_out/r1n_bench.json')); print(j['value'], j['e2e']['value'], j['ms_per_step'])
for f in j['families']: print(f.get('family','')[:40], f.get('value'), f.get('gcups'), f.get('kernel_ms'), (f.get('cpu_baseline') or {}).get('value'), f.get('error'))"
