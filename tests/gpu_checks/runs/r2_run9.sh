#!/bin/bash
# Round-2 GPU pass 9: GroupNorm v2 + LayerNorm lane-group kernel: parity, perf triage, suite, bench.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
timeout 300 python tests/gpu_checks/kernel_cases.py norms > gpurun_out/r2i_norms.log 2>&1; echo "norms=$? t=$(( $(date +%s) - T0 ))"
cut -c1-600 gpurun_out/r2i_norms.log | tail -3
run_norm() { name=$1; shift
  env "$@" timeout 120 python tests/gpu_checks/kernel_cases.py --case perf_norms > gpurun_out/r2i_norm_$name.log 2>&1
  echo "norm_$name rc=$? t=$(( $(date +%s) - T0 ))"; grep RESULT gpurun_out/r2i_norm_$name.log | cut -c1-1000
}
run_norm v1 LECO_GN_IMPL=v1
run_norm v2 LECO_GN_IMPL=v2
run_norm v2_s128 LECO_GN_IMPL=v2 LECO_GN_ROWS=8 LECO_GN_SPLITS=128
run_norm v2_r32 LECO_GN_IMPL=v2 LECO_GN_ROWS=32
run_norm v2_apply16 LECO_GN_IMPL=v2 LECO_GN_APPLY_BPSM=16 LECO_GN_APPLY_RPT=2
run_norm v2_apply4 LECO_GN_IMPL=v2 LECO_GN_APPLY_BPSM=4 LECO_GN_APPLY_RPT=8
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2i_pytest.log 2>&1; echo "pytest=$? t=$(( $(date +%s) - T0 ))"
tail -8 gpurun_out/r2i_pytest.log | cut -c1-600
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_bench.log 2>&1; echo "bench rc=$? t=$(( $(date +%s) - T0 ))"
tail -1 gpurun_out/r2i_bench.log > gpurun_out/r2i_bench.json
python -c "import json; d=json.load(open('gpurun_out/r2i_bench.json')); print(d['ms_per_step'], d['value'], d['loss'], d['gpu_launches'], d['roofline']['ms'], d['roofline']['frac'], d['step_roofline']['frac'], d['phases'])"
