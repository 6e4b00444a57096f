#!/bin/bash
# Round-2 GPU pass 4: stand-alone probes + launch list + full ncu capture of the flash kernel (current build).
set -u
mkdir -p gpurun_out
T0=$(date +%s)
timeout 120 tests/gpu_checks/build/micro_probe > gpurun_out/r2d_micro_probe.txt 2>&1; echo "probe=$? t=$(( $(date +%s) - T0 ))"
tail -40 gpurun_out/r2d_micro_probe.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r2d_launches.csv python tests/gpu_checks/profile_step.py --k 1 > gpurun_out/r2d_prof.log 2>&1
echo "launches=$? t=$(( $(date +%s) - T0 ))"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:flash_attn_fwd_kernel --launch-skip 2 \
    --launch-count 1 -f -o gpurun_out/r2d_flash python tests/gpu_checks/kernel_cases.py flash_perf_4096_m0 \
    > gpurun_out/r2d_ncu_flash.log 2>&1
echo "ncu_flash=$? t=$(( $(date +%s) - T0 ))"
