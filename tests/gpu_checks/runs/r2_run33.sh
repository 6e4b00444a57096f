#!/bin/bash
# Round-2 GPU pass 33: four-pixel conv_in; then the whole -m gpu suite and the default bench on the final build.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
timeout 150 python tests/gpu_checks/kernel_cases.py --case elementwise > gpurun_out/r2ai_elem.log 2>&1; RC=$?
echo "elementwise rc=$RC t=$(( $(date +%s) - T0 ))"; tail -1 gpurun_out/r2ai_elem.log | python -c "
import sys, json
l = sys.stdin.read()
try:
    d = json.loads(l[7:]); print('ok', d['ok'], {k: v for k, v in d['parts'].items() if k.startswith('conv_in') or not v['ok']})
except Exception:
    print(l[:1200])"
timeout 300 python tests/gpu_checks/timeline_step.py --k 2 --out gpurun_out/r2ai_timeline_sd21.md > gpurun_out/r2ai_timeline.log 2>&1; echo "timeline rc=$? t=$(( $(date +%s) - T0 ))"
grep -n "conv_in\|span" gpurun_out/r2ai_timeline_sd21.md | head -8
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2ai_pytest.log 2>&1; echo "pytest=$? t=$(( $(date +%s) - T0 ))"
tail -3 gpurun_out/r2ai_pytest.log | cut -c1-300
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2ai_bench.log 2>&1; echo "bench rc=$? t=$(( $(date +%s) - T0 ))"
tail -1 gpurun_out/r2ai_bench.log > gpurun_out/r2ai_bench.json
python -c "
import json; d=json.load(open('gpurun_out/r2ai_bench.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], d['clocks'], d['loss'])
print('roofline', d['roofline']['ms'], d['roofline']['achieved'], 'attn', d['roofline_attention']['achieved'], 'conv', d['roofline_conv']['achieved'], 'step', d['step_roofline']['achieved'], d['phases']['denoise_step_ms'], d['phases']['tail_ms'], 'random_k', d['random_k']['ms_per_step'])" 2>&1 | cut -c1-600
