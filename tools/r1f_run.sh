#!/bin/bash
# Round-1f GPU pass: longNeedle warp-cooperative traceback — parity, then K3/K5 timing of launch-bounds variants.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_long_needle.py tests/test_host_split.py tests/test_host_genotype.py -m gpu -x -q > gpurun_out/r1f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1f_pytest.log
tail -4 gpurun_out/r1f_pytest.log
bash tools/variant_run.sh 'timeout 200 python tools/time_families.py k3 k5' > gpurun_out/r1f_variants.log 2>&1
cat gpurun_out/r1f_variants.log
