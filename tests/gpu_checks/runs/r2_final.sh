#!/bin/bash
# Round-2 final measurements: default bench line (the driver's command), the other BASELINE rows, ncu launch list of the
# same workload, ncu --set full of the roofline kernel (flash forward) and of the cluster GroupNorm, in-graph timeline.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2f_bench.log 2>&1; echo "bench default rc=$? t=$(( $(date +%s) - T0 ))"
tail -1 gpurun_out/r2f_bench.log > gpurun_out/r2f_bench.json
python -c "
import json; d=json.load(open('gpurun_out/r2f_bench.json'))
print(d['ms_per_step'], d['value'], d['e2e'], d['gpu_launches'], d['clocks'])
print('roofline', d['roofline']); print('attn', d.get('roofline_attention')); print('conv', d.get('roofline_conv')); print('step', d['step_roofline']); print('cpu', d.get('cpu_baseline')); print('random_k', d.get('random_k'))" 2>&1 | cut -c1-700
for CFG in sdxl sd21_b4 sd15_c3lier; do
  timeout 600 python bench.py --config $CFG --steps 4 --warmup 3 --no-cpu-baseline --no-kernel-rooflines > gpurun_out/r2f_bench_$CFG.log 2>&1; echo "bench $CFG rc=$? t=$(( $(date +%s) - T0 ))"
  tail -1 gpurun_out/r2f_bench_$CFG.log > gpurun_out/r2f_bench_$CFG.json
  python -c "import json; d=json.load(open('gpurun_out/r2f_bench_$CFG.json')); print(d['ms_per_step'], d['value'], d['step_roofline']['achieved'], d.get('peak_mem_gb'), d['phases']['denoise_step_ms'], d['phases']['tail_ms'])" 2>&1 | cut -c1-300
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2f_launches.csv python tests/gpu_checks/profile_step.py --k 1 > gpurun_out/r2f_profile_step.log 2>&1; echo "launch list rc=$? t=$(( $(date +%s) - T0 ))"
python tests/gpu_checks/summarize_launches.py gpurun_out/r2f_launches.csv > gpurun_out/r2f_launches_k1_iteration.md 2>&1
python tests/gpu_checks/summarize_launches.py gpurun_out/r2f_launches.csv guided_step > gpurun_out/r2f_launches_k1_denoise_forward.md 2>&1
head -14 gpurun_out/r2f_launches_k1_denoise_forward.md | cut -c1-150
timeout 300 ncu --set full --clock-control none --import-source on -k regex:flash_attn_fwd_ts --launch-skip 2 --launch-count 1 -f -o gpurun_out/r2f_ncu_flash_fwd python tests/gpu_checks/kernel_cases.py --case flash_perf_4096_m0 > gpurun_out/r2f_ncu_flash.log 2>&1; echo "ncu flash rc=$? t=$(( $(date +%s) - T0 ))"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gn_cluster --launch-skip 2 --launch-count 1 -f -o gpurun_out/r2f_ncu_gn_cluster python tests/gpu_checks/kernel_cases.py --case perf_norms > gpurun_out/r2f_ncu_gn.log 2>&1; echo "ncu gn rc=$? t=$(( $(date +%s) - T0 ))"
timeout 300 python tests/gpu_checks/timeline_step.py --k 2 --out gpurun_out/r2f_timeline_sd21.md > gpurun_out/r2f_timeline_sd21.log 2>&1; echo "timeline rc=$? t=$(( $(date +%s) - T0 ))"
rm -f gpurun_out/r2f_launches.csv.tmp; ls -la gpurun_out | grep r2f | cut -c30-120
