#!/bin/bash
# Round-2 GPU pass 23: wide flash forward after the pv_done phase fix (all cases, repeated), end-time kernel timelines.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
for R in 1 2 3; do
  timeout 400 python tests/gpu_checks/kernel_cases.py flash_wide flash_perf_wide > gpurun_out/r2x_cases_$R.log 2>&1; echo "cases run $R rc=$? t=$(( $(date +%s) - T0 ))"
  grep -v '"ok": true' gpurun_out/r2x_cases_$R.log | cut -c1-400 | tail -6
done
grep flash_perf gpurun_out/r2x_cases_3.log | cut -c1-500
timeout 300 python tests/gpu_checks/timeline_step.py --k 2 --out gpurun_out/r2x_timeline_sd21.md > gpurun_out/r2x_timeline_sd21.log 2>&1; echo "timeline sd21 rc=$? t=$(( $(date +%s) - T0 ))"
tail -3 gpurun_out/r2x_timeline_sd21.log | cut -c1-300
timeout 300 python tests/gpu_checks/timeline_step.py --arch sd15 --batch 4 --rank 8 --c3lier --k 2 --out gpurun_out/r2x_timeline_sd15.md > gpurun_out/r2x_timeline_sd15.log 2>&1; echo "timeline sd15 rc=$? t=$(( $(date +%s) - T0 ))"
timeout 400 python bench.py --config sd15_c3lier --steps 3 --warmup 3 --no-cpu-baseline --no-kernel-rooflines > gpurun_out/r2x_bench_sd15.log 2>&1; echo "bench sd15 rc=$? t=$(( $(date +%s) - T0 ))"
tail -1 gpurun_out/r2x_bench_sd15.log > gpurun_out/r2x_bench_sd15.json
python -c "import json; d=json.load(open('gpurun_out/r2x_bench_sd15.json')); print(d['ms_per_step'], d['value'], d['loss'], d['phases'])" 2>&1 | cut -c1-400
