#!/bin/bash
# Round-2 GPU pass 2: full parity suite (no -x), fused-LoRA (second-UMMA) cases + triage, bench A/B of the LoRA paths.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python tests/gpu_checks/gemm_cases.py fl_ triage_fl_perf > gpurun_out/r2b_gemm_fl.log 2>&1; echo "gemm_fl=$? t=$(( $(date +%s) - T0 ))"
tail -15 gpurun_out/r2b_gemm_fl.log
cp gpurun_out/gemm_cases.json gpurun_out/r2b_gemm_fl.json 2>/dev/null
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2b_pytest.log 2>&1; echo "pytest=$? t=$(( $(date +%s) - T0 ))"
tail -40 gpurun_out/r2b_pytest.log
LECO_FUSED_LORA=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-kernel-rooflines > gpurun_out/r2b_bench_fl0.log 2>&1; echo "bench_fl0=$? t=$(( $(date +%s) - T0 ))"
tail -c 1500 gpurun_out/r2b_bench_fl0.log | head -c 700; echo
LECO_FUSED_LORA=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-kernel-rooflines > gpurun_out/r2b_bench_fl1.log 2>&1; echo "bench_fl1=$? t=$(( $(date +%s) - T0 ))"
tail -c 1500 gpurun_out/r2b_bench_fl1.log | head -c 700; echo
