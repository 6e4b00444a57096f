#!/bin/bash
# Round-2 GPU pass 27: cluster GroupNorm, division-free fold + cluster size chosen by resident clusters: parity, A/B on
# the batch-2 and batch-4 workloads.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
timeout 150 python tests/gpu_checks/kernel_cases.py --case norms > gpurun_out/r2ab_norms.log 2>&1; RC=$?
echo "norms rc=$RC t=$(( $(date +%s) - T0 ))"; tail -1 gpurun_out/r2ab_norms.log | cut -c1-60
timeout 200 python tests/gpu_checks/kernel_cases.py --case perf_norms 2>&1 | grep RESULT | cut -c1-640
echo "perf t=$(( $(date +%s) - T0 ))"
if [ $RC -eq 0 ]; then
  for CFG in sd21 sd21_b4 sd15_c3lier; do
    for V in "LECO_GN_IMPL=v3" "LECO_GN_IMPL=v2"; do
      if [ $CFG = sd21 ] && [ $V = "LECO_GN_IMPL=v2" ]; then continue; fi
      env $V timeout 400 python bench.py --config $CFG --steps 4 --warmup 3 --no-cpu-baseline --no-kernel-rooflines > gpurun_out/r2ab_bench_${CFG}_$V.log 2>&1; echo "bench $CFG $V rc=$? t=$(( $(date +%s) - T0 ))"
      tail -1 gpurun_out/r2ab_bench_${CFG}_$V.log > gpurun_out/r2ab_bench_${CFG}_$V.json
      python -c "import json; d=json.load(open('gpurun_out/r2ab_bench_${CFG}_$V.json')); print(d['ms_per_step'], d['value'], d['loss'], d['gpu_launches'], d['phases']['denoise_step_ms'], d['phases']['tail_ms'])" 2>&1 | cut -c1-300
    done
  done
fi
