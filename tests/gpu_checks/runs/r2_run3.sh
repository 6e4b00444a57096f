#!/bin/bash
# Round-2 GPU pass 3: TMA-store epilogue + single-launch GroupNorm + ordered GN reductions: parity, then A/B timing.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python tests/gpu_checks/gemm_cases.py > gpurun_out/r2c_gemm.log 2>&1; echo "gemm=$? t=$(( $(date +%s) - T0 ))"
grep -v '"ok": true' gpurun_out/r2c_gemm.log | tail -15; tail -1 gpurun_out/r2c_gemm.log
cp gpurun_out/gemm_cases.json gpurun_out/r2c_gemm.json 2>/dev/null
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2c_pytest.log 2>&1; echo "pytest=$? t=$(( $(date +%s) - T0 ))"
tail -40 gpurun_out/r2c_pytest.log | cut -c1-600
for v in "1 1" "0 1" "1 0"; do
  set -- $v
  LECO_TMA_STORE=$1 LECO_GN_FUSED=$2 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-kernel-rooflines > gpurun_out/r2c_bench_tma$1_gn$2.log 2>&1
  echo "bench tma=$1 gn=$2 rc=$? t=$(( $(date +%s) - T0 ))"
  tail -1 gpurun_out/r2c_bench_tma$1_gn$2.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['loss'], d['gpu_launches'])"
done
