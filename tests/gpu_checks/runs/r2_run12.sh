#!/bin/bash
# Round-2 GPU pass 12: FL kernels with N-inner tile runs (T reuse), conv_out on tensor cores: parity, suite, bench.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
timeout 600 python tests/gpu_checks/gemm_cases.py fl_ triage_fl_perf splitk > gpurun_out/r2l_gemm.log 2>&1; echo "gemm=$? t=$(( $(date +%s) - T0 ))"
cut -c1-900 gpurun_out/r2l_gemm.log | grep -v '"ok": true' | tail -8; grep triage_fl gpurun_out/r2l_gemm.log | cut -c1-900; tail -1 gpurun_out/r2l_gemm.log
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2l_pytest.log 2>&1; echo "pytest=$? t=$(( $(date +%s) - T0 ))"
tail -8 gpurun_out/r2l_pytest.log | cut -c1-600
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-kernel-rooflines > gpurun_out/r2l_bench.log 2>&1; echo "bench rc=$? t=$(( $(date +%s) - T0 ))"
tail -1 gpurun_out/r2l_bench.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['loss'], d['gpu_launches'], d['phases']['denoise_step_ms'], d['phases']['tail_ms'])"
LECO_CONV_OUT_GEMM=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-kernel-rooflines > gpurun_out/r2l_bench_co0.log 2>&1; echo "bench_co0 rc=$? t=$(( $(date +%s) - T0 ))"
tail -1 gpurun_out/r2l_bench_co0.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['loss'], d['gpu_launches'], d['phases']['denoise_step_ms'], d['phases']['tail_ms'])"
