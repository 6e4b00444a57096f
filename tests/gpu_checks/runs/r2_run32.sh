#!/bin/bash
# Round-2 GPU pass 32: GroupNorm backward with loads in flight + warp-parallel ordered fold: parity, tail cost, grads.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
timeout 150 python tests/gpu_checks/kernel_cases.py --case norms > gpurun_out/r2ah_norms.log 2>&1; RC=$?
echo "norms rc=$RC t=$(( $(date +%s) - T0 ))"; tail -1 gpurun_out/r2ah_norms.log | cut -c1-60
LECO_GN_IMPL=v2 timeout 150 python tests/gpu_checks/kernel_cases.py --case norms 2>&1 | tail -1 | cut -c1-60
timeout 300 python tests/gpu_checks/timeline_step.py --k 2 --out gpurun_out/r2ah_timeline_sd21.md > gpurun_out/r2ah_timeline.log 2>&1; echo "timeline rc=$? t=$(( $(date +%s) - T0 ))"
grep -n "gn_bwd\|span" gpurun_out/r2ah_timeline_sd21.md | head -8
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "grads or fullsize or iteration or determinism or norms" > gpurun_out/r2ah_pytest.log 2>&1; echo "pytest=$? t=$(( $(date +%s) - T0 ))"
tail -3 gpurun_out/r2ah_pytest.log | cut -c1-300
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-kernel-rooflines > gpurun_out/r2ah_bench.log 2>&1; echo "bench rc=$? t=$(( $(date +%s) - T0 ))"
tail -1 gpurun_out/r2ah_bench.log | python -c "import sys, json; d=json.loads(sys.stdin.read()); print('bench', d['ms_per_step'], d['value'], d['loss'], d['phases']['denoise_step_ms'], d['phases']['tail_ms'])"
