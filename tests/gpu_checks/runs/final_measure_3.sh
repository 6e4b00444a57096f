#!/bin/bash
# ncu --set full captures of the kernels round 2 starts from (library-chosen 2-CTA conv, short-K GEMM, flash attention)
set -u
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 --launch-skip 4 --launch-count 1 -f \
    -o gpurun_out/final3_conv_2cta python tests/gpu_checks/gemm_cases.py perfauto_conv_64_320 > gpurun_out/final3_a.log 2>&1; echo "a=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 --launch-skip 4 --launch-count 1 -f \
    -o gpurun_out/final3_qkv python tests/gpu_checks/gemm_cases.py perfauto_qkv_320 > gpurun_out/final3_b.log 2>&1; echo "b=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:flash_attn_fwd_kernel --launch-skip 2 --launch-count 1 -f \
    -o gpurun_out/final3_flash python tests/gpu_checks/kernel_cases.py flash_perf_4096_m0 > gpurun_out/final3_c.log 2>&1; echo "c=$?"
