#!/bin/bash
# Round-2 GPU pass 31: split-K finalize with two vectors in flight: split-K parity cases, in-graph cost, bench.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
timeout 500 python tests/gpu_checks/gemm_cases.py splitk fl_splitk fl_conv_8 > gpurun_out/r2ag_gemm.log 2>&1; echo "gemm cases rc=$? t=$(( $(date +%s) - T0 ))"
grep -v '"ok": true' gpurun_out/r2ag_gemm.log | cut -c1-300 | tail -5
timeout 300 python tests/gpu_checks/timeline_step.py --k 2 --out gpurun_out/r2ag_timeline_sd21.md > gpurun_out/r2ag_timeline.log 2>&1; echo "timeline rc=$? t=$(( $(date +%s) - T0 ))"
grep -n "splitk\|span" gpurun_out/r2ag_timeline_sd21.md | head -8
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-kernel-rooflines > gpurun_out/r2ag_bench.log 2>&1; echo "bench rc=$? t=$(( $(date +%s) - T0 ))"
tail -1 gpurun_out/r2ag_bench.log | python -c "import sys, json; d=json.loads(sys.stdin.read()); print('bench', d['ms_per_step'], d['value'], d['loss'], d['phases']['denoise_step_ms'], d['phases']['tail_ms'])"
