#!/bin/bash
# Round-2 GPU pass 30: share of the flash-forward exponentials on the FMA pipe (0 / 1 / 2 / 3 quarters), same box.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
for R in 1 2; do
for P in 1 0 2 3; do
  LECO_FLASH_POLY=$P timeout 120 python tests/gpu_checks/kernel_cases.py --case flash_perf_4096_m0 2>&1 | grep RESULT | python -c "
import sys, json
d = json.loads(sys.stdin.read()[7:]); print('poly quarters $P run $R:', d['ok'], round(d['flash_ms'], 4), 'ms', round(d['flash_tflops'], 1), 'TF/s  rel', round(d['rel'], 5), ' sdpa', round(d['torch_sdpa_ms'], 4))"
done
done
echo "t=$(( $(date +%s) - T0 ))"
for P in 1 2; do
  LECO_FLASH_POLY=$P timeout 120 python tests/gpu_checks/kernel_cases.py --case flash_rescale_2048 2>&1 | grep RESULT | cut -c1-200
  LECO_FLASH_POLY=$P timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-kernel-rooflines > gpurun_out/r2ae_bench_p$P.log 2>&1
  tail -1 gpurun_out/r2ae_bench_p$P.log | python -c "import sys, json; d=json.loads(sys.stdin.read()); print('bench poly $P', d['ms_per_step'], d['value'], d['phases']['denoise_step_ms'])"
done
echo "t=$(( $(date +%s) - T0 ))"
