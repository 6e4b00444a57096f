#!/bin/bash
# Round-2 GPU pass 7: GroupNorm launch-shape triage, flash forward with tree max, parity suite, bench with phases.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
run_norm() { # name, env...
  name=$1; shift
  env "$@" timeout 120 python tests/gpu_checks/kernel_cases.py --case perf_norms > gpurun_out/r2g_norm_$name.log 2>&1
  echo "norm_$name rc=$? t=$(( $(date +%s) - T0 ))"; grep RESULT gpurun_out/r2g_norm_$name.log | cut -c1-900
}
run_norm A X=1
run_norm B LECO_GN_ROWS=8 LECO_GN_SPLITS=128
run_norm C LECO_GN_ROWS=32
run_norm D LECO_GN_APPLY_BPSM=16 LECO_GN_APPLY_RPT=2
run_norm E LECO_GN_APPLY_BPSM=4 LECO_GN_APPLY_RPT=8
run_norm F LECO_GN_FUSED=1
run_norm G LECO_GN_ROWS=8 LECO_GN_SPLITS=128 LECO_GN_APPLY_BPSM=16 LECO_GN_APPLY_RPT=2
timeout 300 python tests/gpu_checks/kernel_cases.py flash_perf > gpurun_out/r2g_flash.log 2>&1; echo "flash=$? t=$(( $(date +%s) - T0 ))"
cut -c1-420 gpurun_out/r2g_flash.log | tail -4
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2g_pytest.log 2>&1; echo "pytest=$? t=$(( $(date +%s) - T0 ))"
tail -8 gpurun_out/r2g_pytest.log | cut -c1-600
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench.log 2>&1; echo "bench rc=$? t=$(( $(date +%s) - T0 ))"
tail -1 gpurun_out/r2g_bench.log > gpurun_out/r2g_bench.json
python -c "import json; d=json.load(open('gpurun_out/r2g_bench.json')); print(d['ms_per_step'], d['value'], d['loss'], d['gpu_launches'], d['roofline']['ms'], d['roofline']['frac'], d['roofline_attention']['frac'], d['step_roofline']['frac'], d['phases'])"
