#!/bin/bash
# Round-2 GPU pass 21: tensor-core tn_reduce (LoRA weight gradients): kernel cases, gradient parity, bench.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python tests/gpu_checks/kernel_cases.py training_kernels determinism engine_grads > gpurun_out/r2v_cases.log 2>&1; echo "cases=$? t=$(( $(date +%s) - T0 ))"
cut -c1-700 gpurun_out/r2v_cases.log | tail -12
python - <<'PY'
import json
d = json.load(open("gpurun_out/kernel_cases.json"))
print({k: v for k, v in d["training_kernels"]["parts"].items() if k.startswith("tn_")})
PY
LECO_TN_MMA=0 timeout 300 python tests/gpu_checks/kernel_cases.py --case training_kernels | cut -c1-3000 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('RESULT '):
        d = json.loads(l[7:]); print('FMA kernel:', d['parts']['tn_strided'])"
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "fullsize or iteration or grads" > gpurun_out/r2v_pytest.log 2>&1; echo "pytest=$? t=$(( $(date +%s) - T0 ))"
tail -5 gpurun_out/r2v_pytest.log | cut -c1-400
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2v_bench.log 2>&1; echo "bench rc=$? t=$(( $(date +%s) - T0 ))"
tail -1 gpurun_out/r2v_bench.log > gpurun_out/r2v_bench.json
python -c "import json; d=json.load(open('gpurun_out/r2v_bench.json')); print(d['ms_per_step'], d['value'], d['loss'], d['gpu_launches'], d['roofline']['ms'], d['roofline']['frac'], d['roofline_attention']['frac'], d['step_roofline']['frac'], d['phases']['denoise_step_ms'], d['phases']['tail_ms'])"
