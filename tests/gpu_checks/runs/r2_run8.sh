#!/bin/bash
# Round-2 GPU pass 8: ncu full captures of the norm kernels at the 64x64 level (why ~10 us for a 10 MB pass?)
set -u
mkdir -p gpurun_out
T0=$(date +%s)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gn_stats_kernel|gn_apply_kernel|ln_fwd_kernel" --launch-skip 2 \
    --launch-count 2 -f -o gpurun_out/r2h_gn python tests/gpu_checks/kernel_cases.py --case perf_norms > gpurun_out/r2h_ncu_gn.log 2>&1
echo "ncu_gn=$? t=$(( $(date +%s) - T0 ))"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"ln_fwd_kernel" --launch-skip 1 \
    --launch-count 1 -f -o gpurun_out/r2h_ln python tests/gpu_checks/kernel_cases.py --case perf_norms > gpurun_out/r2h_ncu_ln.log 2>&1
echo "ncu_ln=$? t=$(( $(date +%s) - T0 ))"
