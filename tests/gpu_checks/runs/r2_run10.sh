#!/bin/bash
# Round-2 GPU pass 10: split-K with in-kernel finalize: parity cases, suite, bench A/B.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
timeout 600 python tests/gpu_checks/gemm_cases.py splitk fl_splitk fl_conv_8 conv_ ragged tiny bias_ lora > gpurun_out/r2j_gemm.log 2>&1; echo "gemm=$? t=$(( $(date +%s) - T0 ))"
cut -c1-300 gpurun_out/r2j_gemm.log | tail -24
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2j_pytest.log 2>&1; echo "pytest=$? t=$(( $(date +%s) - T0 ))"
tail -8 gpurun_out/r2j_pytest.log | cut -c1-600
for v in 1 0; do
  LECO_SPLITK_FUSED=$v timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-kernel-rooflines > gpurun_out/r2j_bench_sk$v.log 2>&1
  echo "bench sk=$v rc=$? t=$(( $(date +%s) - T0 ))"
  tail -1 gpurun_out/r2j_bench_sk$v.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['loss'], d['gpu_launches'], d['phases']['denoise_step_ms'], d['phases']['tail_ms'])"
done
