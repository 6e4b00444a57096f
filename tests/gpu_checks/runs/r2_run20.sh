#!/bin/bash
# Round-2 GPU pass 20: ncu --set full of the TMEM flash forward kernel at the bench's roofline shape.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:flash_attn_fwd_ts_kernel --launch-skip 2 \
    --launch-count 1 -f -o gpurun_out/r2u_flash_ts python tests/gpu_checks/kernel_cases.py --case flash_perf_4096_m0 \
    > gpurun_out/r2u_ncu_flash.log 2>&1
echo "ncu_flash=$? t=$(( $(date +%s) - T0 ))"
tail -3 gpurun_out/r2u_ncu_flash.log | cut -c1-300
