#!/bin/bash
# Round-end measurement on one B200 (run through gpurun from the repo root); everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/final_bench.log 2>&1; echo "bench=$?"
tail -1 gpurun_out/final_bench.log > gpurun_out/final_bench.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/final_bench_ref.log 2>&1; echo "ref=$?"
tail -1 gpurun_out/final_bench_ref.log > gpurun_out/final_bench_ref.json
# launch list of one iteration (k = 1: one CFG denoise step + the tail), serialised / cold cache
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/final_launches.csv python tests/gpu_checks/profile_step.py --k 1 > gpurun_out/final_prof.log 2>&1
echo "launches=$?"
# full-section capture of the dominant kernel (implicit-GEMM conv 320->320 @64x64, 4 samples)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 --launch-skip 4 \
    --launch-count 1 -f -o gpurun_out/final_conv python tests/gpu_checks/gemm_cases.py perf_conv_64_320 \
    > gpurun_out/final_ncu_conv.log 2>&1
echo "ncu_conv=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:flash_attn_fwd_kernel --launch-skip 2 \
    --launch-count 1 -f -o gpurun_out/final_flash python tests/gpu_checks/kernel_cases.py flash_perf_4096_m0 \
    > gpurun_out/final_ncu_flash.log 2>&1
echo "ncu_flash=$?"
cat gpurun_out/final_bench.json
cat gpurun_out/final_bench_ref.json
