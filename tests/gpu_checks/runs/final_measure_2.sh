set -u
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 --launch-skip 4 \
    --launch-count 1 -f -o gpurun_out/final2_conv python tests/gpu_checks/gemm_cases.py perf_conv_64_320 > gpurun_out/final2_ncu_conv.log 2>&1
echo "ncu_conv=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/final2_launches.csv python tests/gpu_checks/profile_step.py --k 1 > gpurun_out/final2_prof.log 2>&1
echo "launches=$?"
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/final2_bench.log 2>&1; echo "bench=$?"
tail -1 gpurun_out/final2_bench.log
