#!/bin/bash
# Round-2 GPU pass 22: flash forward for head dims 64 < d <= 192 (SD1.5 levels), in-graph kernel timelines, SD1.5 A/B.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
timeout 120 python tests/gpu_checks/kernel_cases.py --case flash_wide_d80_1024 > gpurun_out/r2w_first.log 2>&1; RC=$?
echo "first wide case rc=$RC t=$(( $(date +%s) - T0 ))"; tail -2 gpurun_out/r2w_first.log | cut -c1-600
if [ $RC -eq 0 ]; then
  timeout 700 python tests/gpu_checks/kernel_cases.py flash_wide flash_perf_wide > gpurun_out/r2w_cases.log 2>&1; echo "cases=$? t=$(( $(date +%s) - T0 ))"
  cut -c1-500 gpurun_out/r2w_cases.log | tail -14
fi
timeout 300 python tests/gpu_checks/timeline_step.py --k 2 --out gpurun_out/r2w_timeline_sd21.md > gpurun_out/r2w_timeline_sd21.log 2>&1; echo "timeline sd21 rc=$? t=$(( $(date +%s) - T0 ))"
head -40 gpurun_out/r2w_timeline_sd21.md | cut -c1-200; tail -3 gpurun_out/r2w_timeline_sd21.log | cut -c1-300
if [ $RC -eq 0 ]; then
  timeout 400 python -m pytest tests -m gpu -q -p no:cacheprovider -k "full_size_sd15 or tiny15" > gpurun_out/r2w_pytest.log 2>&1; echo "pytest=$? t=$(( $(date +%s) - T0 ))"
  tail -4 gpurun_out/r2w_pytest.log | cut -c1-400
fi
for W in 0 1; do
  LECO_FLASH_WIDE=$W timeout 400 python bench.py --config sd15_c3lier --steps 3 --warmup 3 --no-cpu-baseline --no-kernel-rooflines > gpurun_out/r2w_bench_sd15_w$W.log 2>&1; echo "bench sd15 wide=$W rc=$? t=$(( $(date +%s) - T0 ))"
  tail -1 gpurun_out/r2w_bench_sd15_w$W.log > gpurun_out/r2w_bench_sd15_w$W.json
  python -c "import json; d=json.load(open('gpurun_out/r2w_bench_sd15_w$W.json')); print(d['ms_per_step'], d['value'], d['loss'], d['phases'])" 2>&1 | cut -c1-400
done
timeout 300 python tests/gpu_checks/timeline_step.py --arch sd15 --batch 4 --rank 8 --c3lier --k 2 --out gpurun_out/r2w_timeline_sd15.md > gpurun_out/r2w_timeline_sd15.log 2>&1; echo "timeline sd15 rc=$? t=$(( $(date +%s) - T0 ))"
head -30 gpurun_out/r2w_timeline_sd15.md | cut -c1-200
