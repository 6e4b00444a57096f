#!/bin/bash
# same-box A/B/C of GEMM epilogue variants (perf triage only)
for v in A B C; do
  cp ab/lib$v.so leco_b200/csrc/libleco_b200.so
  timeout 150 python tests/gpu_checks/gemm_cases.py triage_shape > gpurun_out/ab_$v.log 2>&1
  cp gpurun_out/gemm_cases.json gpurun_out/ab_triage_$v.json
done
for v in A C; do
  cp ab/lib$v.so leco_b200/csrc/libleco_b200.so
  timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ab_bench_$v.log 2>&1
  echo "bench $v: $(tail -1 gpurun_out/ab_bench_$v.log | cut -c1-160)"
done
cp ab/libC.so leco_b200/csrc/libleco_b200.so
timeout 300 python tests/gpu_checks/gemm_cases.py basic lora geglu conv_ splitk ragged tiny multi > gpurun_out/ab_C_cases.log 2>&1
tail -1 gpurun_out/ab_C_cases.log
grep '"ok": false' gpurun_out/ab_C_cases.log | cut -c1-200
