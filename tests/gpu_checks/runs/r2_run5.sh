#!/bin/bash
# Round-2 GPU pass 5: fused attention backward (flash_attn_bwd): kernel cases first (each in its own subprocess), then the
# whole parity suite and a bench A/B of the backward implementations.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python tests/gpu_checks/kernel_cases.py attn_ flash_ > gpurun_out/r2e_attn.log 2>&1; echo "attn=$? t=$(( $(date +%s) - T0 ))"
cut -c1-400 gpurun_out/r2e_attn.log | tail -40
cp gpurun_out/kernel_cases.json gpurun_out/r2e_attn.json 2>/dev/null
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2e_pytest.log 2>&1; echo "pytest=$? t=$(( $(date +%s) - T0 ))"
tail -30 gpurun_out/r2e_pytest.log | cut -c1-600
for v in flash v0; do
  LECO_ATTENTION_BWD=$v timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-kernel-rooflines > gpurun_out/r2e_bench_$v.log 2>&1
  echo "bench bwd=$v rc=$? t=$(( $(date +%s) - T0 ))"
  tail -1 gpurun_out/r2e_bench_$v.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['loss'], d['gpu_launches'], d['peak_memory_bytes'])"
done
