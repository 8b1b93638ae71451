#!/bin/bash
# Round-1g GPU pass: wavefront engine changes (L2 prefetch of partner rows, ring hand-off between warps) — parity, then timing of variants.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_long_needle.py tests/test_edit_path.py tests/test_host_split.py tests/test_host_genotype.py -m gpu -x -q > gpurun_out/r1g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1g_pytest.log
tail -4 gpurun_out/r1g_pytest.log
K5N=148 bash tools/variant_run.sh 'K5N=148 timeout 200 python tools/time_families.py k3 k5' > gpurun_out/r1g_variants.log 2>&1
cat gpurun_out/r1g_variants.log
