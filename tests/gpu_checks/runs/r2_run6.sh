#!/bin/bash
# Round-2 GPU pass 6: flash forward with decoupled softmax warpgroups: attention cases, parity suite, bench.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python tests/gpu_checks/kernel_cases.py attn_ flash_ > gpurun_out/r2f_attn.log 2>&1; echo "attn=$? t=$(( $(date +%s) - T0 ))"
cut -c1-420 gpurun_out/r2f_attn.log | tail -30
cp gpurun_out/kernel_cases.json gpurun_out/r2f_attn.json 2>/dev/null
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2f_pytest.log 2>&1; echo "pytest=$? t=$(( $(date +%s) - T0 ))"
tail -12 gpurun_out/r2f_pytest.log | cut -c1-600
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench.log 2>&1; echo "bench rc=$? t=$(( $(date +%s) - T0 ))"
tail -1 gpurun_out/r2f_bench.log > gpurun_out/r2f_bench.json
python -c "import json; d=json.load(open('gpurun_out/r2f_bench.json')); print(d['ms_per_step'], d['value'], d['loss'], d['gpu_launches'], d['roofline']['ms'], d['roofline']['frac'], d['roofline_attention']['frac'], d['step_roofline']['frac'])"
