#!/bin/bash
# Round-2 GPU pass 19: TMEM flash forward with packed f32x2 softmax arithmetic and a quarter of the exponentials on the FMA pipe: attention cases + bench.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python tests/gpu_checks/kernel_cases.py attn_ flash_ > gpurun_out/r2t_attn.log 2>&1; echo "attn=$? t=$(( $(date +%s) - T0 ))"
cut -c1-420 gpurun_out/r2t_attn.log | grep -v '"ok": true' | tail; grep flash_perf gpurun_out/r2t_attn.log | cut -c1-420; tail -1 gpurun_out/r2t_attn.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2t_bench.log 2>&1; echo "bench rc=$? t=$(( $(date +%s) - T0 ))"
tail -1 gpurun_out/r2t_bench.log > gpurun_out/r2t_bench.json
python -c "import json; d=json.load(open('gpurun_out/r2t_bench.json')); print(d['ms_per_step'], d['value'], d['loss'], d['gpu_launches'], d['roofline']['ms'], d['roofline']['frac'], d['roofline_attention']['frac'], d['step_roofline']['frac'], d['phases']['denoise_step_ms'], d['phases']['tail_ms'])"
