#!/bin/bash
set -u
mkdir -p gpurun_out
T0=$(date +%s)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r2k_launches.csv python tests/gpu_checks/profile_step.py --k 1 > gpurun_out/r2k_prof.log 2>&1
echo "launches=$? t=$(( $(date +%s) - T0 ))"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:flash_attn_fwd_kernel --launch-skip 2 \
    --launch-count 1 -f -o gpurun_out/r2k_flash python tests/gpu_checks/kernel_cases.py flash_perf_4096_m0 \
    > gpurun_out/r2k_ncu_flash.log 2>&1
echo "ncu_flash=$? t=$(( $(date +%s) - T0 ))"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:flash_attn_bwd_kernel --launch-skip 0 \
    --launch-count 1 -f -o gpurun_out/r2k_flash_bwd python tests/gpu_checks/kernel_cases.py --case attn_self_4096_d64 \
    > gpurun_out/r2k_ncu_flash_bwd.log 2>&1
echo "ncu_flash_bwd=$? t=$(( $(date +%s) - T0 ))"
