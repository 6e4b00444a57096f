#!/bin/bash
# Round-2 GPU pass 28: the whole -m gpu suite on the final kernels (wide flash forward, cluster GroupNorm).
set -u
mkdir -p gpurun_out
T0=$(date +%s)
timeout 1700 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=15 > gpurun_out/r2ac_pytest.log 2>&1; echo "pytest=$? t=$(( $(date +%s) - T0 ))"
tail -30 gpurun_out/r2ac_pytest.log | cut -c1-300
