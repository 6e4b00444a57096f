#!/bin/bash
# Round-2 GPU pass 25: cluster GroupNorm with the loads actually in flight: parity, micro A/B, step A/B, timeline.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
timeout 150 python tests/gpu_checks/kernel_cases.py --case norms > gpurun_out/r2z_norms.log 2>&1; RC=$?
echo "norms rc=$RC t=$(( $(date +%s) - T0 ))"; tail -1 gpurun_out/r2z_norms.log | cut -c1-60
for V in "LECO_GN_IMPL=v2" "LECO_GN_IMPL=v3" "LECO_GN_CLUSTER=8"; do
  env $V timeout 200 python tests/gpu_checks/kernel_cases.py --case perf_norms 2>&1 | grep RESULT | cut -c1-700 | sed "s/^/$V /"
done
echo "perf t=$(( $(date +%s) - T0 ))"
if [ $RC -eq 0 ]; then
  for V in "LECO_GN_IMPL=v3" "LECO_GN_IMPL=v2"; do
    env $V timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-kernel-rooflines > gpurun_out/r2z_bench_$V.log 2>&1; echo "bench $V rc=$? t=$(( $(date +%s) - T0 ))"
    tail -1 gpurun_out/r2z_bench_$V.log > gpurun_out/r2z_bench_$V.json
    python -c "import json; d=json.load(open('gpurun_out/r2z_bench_$V.json')); print(d['ms_per_step'], d['value'], d['loss'], d['gpu_launches'], d['phases']['denoise_step_ms'], d['phases']['tail_ms'])" 2>&1 | cut -c1-300
  done
  timeout 300 python tests/gpu_checks/timeline_step.py --k 2 --out gpurun_out/r2z_timeline_sd21.md > gpurun_out/r2z_timeline_sd21.log 2>&1; echo "timeline rc=$? t=$(( $(date +%s) - T0 ))"
  grep -n "gn_\|span" gpurun_out/r2z_timeline_sd21.md | head -12
fi
