#!/bin/bash
# Development aid: parity tests of everything built on the wavefront engine, then kernel timing (also of variants/*.so).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_long_needle.py tests/test_edit_path.py tests/test_host_split.py tests/test_host_genotype.py -m gpu -x -q > gpurun_out/ln_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ln_pytest.log
tail -4 gpurun_out/ln_pytest.log
bash tools/variant_run.sh 'timeout 300 python tools/time_families.py k3 k3w' > gpurun_out/ln_time.log 2>&1; cat gpurun_out/ln_time.log
