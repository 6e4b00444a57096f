#!/bin/bash
# Round-2 GPU pass 24: one-launch cluster GroupNorm (leco_group_norm_v3): parity, micro A/B, whole-step A/B.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
timeout 150 python tests/gpu_checks/kernel_cases.py --case norms > gpurun_out/r2y_norms.log 2>&1; RC=$?
echo "norms rc=$RC t=$(( $(date +%s) - T0 ))"; tail -1 gpurun_out/r2y_norms.log | python -c "
import sys, json
l = sys.stdin.read()
try:
    d = json.loads(l[7:]); print('ok', d['ok'], {k: v for k, v in d['parts'].items() if not v['ok']})
except Exception as e:
    print('unparsed', l[:1500])"
LECO_GN_CLUSTER=8 timeout 150 python tests/gpu_checks/kernel_cases.py --case norms > gpurun_out/r2y_norms_c8.log 2>&1; echo "norms cluster8 rc=$? t=$(( $(date +%s) - T0 ))"
tail -1 gpurun_out/r2y_norms_c8.log | cut -c1-80
for V in "LECO_GN_IMPL=v2" "LECO_GN_IMPL=v3" "LECO_GN_CLUSTER=8"; do
  env $V timeout 200 python tests/gpu_checks/kernel_cases.py --case perf_norms 2>&1 | grep RESULT | cut -c1-1200 | sed "s/^/$V /"
done
echo "perf t=$(( $(date +%s) - T0 ))"
if [ $RC -eq 0 ]; then
  for V in "LECO_GN_IMPL=v3" "LECO_GN_IMPL=v2" "LECO_GN_CLUSTER=8"; do
    env $V timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-kernel-rooflines > gpurun_out/r2y_bench_$V.log 2>&1; echo "bench $V rc=$? t=$(( $(date +%s) - T0 ))"
    tail -1 gpurun_out/r2y_bench_$V.log > gpurun_out/r2y_bench_$V.json
    python -c "import json; d=json.load(open('gpurun_out/r2y_bench_$V.json')); print(d['ms_per_step'], d['value'], d['loss'], d['gpu_launches'], d['phases']['denoise_step_ms'], d['phases']['tail_ms'])" 2>&1 | cut -c1-300
  done
  timeout 500 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "engine_fwd or full_size or norms or iteration_matches" > gpurun_out/r2y_pytest.log 2>&1; echo "pytest=$? t=$(( $(date +%s) - T0 ))"
  tail -4 gpurun_out/r2y_pytest.log | cut -c1-400
fi
