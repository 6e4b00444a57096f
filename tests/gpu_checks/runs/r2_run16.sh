#!/bin/bash
# Round-2 GPU pass 16: text-encoder prologue (kernels, encoders at full size, checkpoint-directory training).
set -u
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python tests/gpu_checks/kernel_cases.py text_ > gpurun_out/r2q_text.log 2>&1; echo "text=$? t=$(( $(date +%s) - T0 ))"
cut -c1-1200 gpurun_out/r2q_text.log | tail -12
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -k "checkpoint_directory or text_" > gpurun_out/r2q_pytest.log 2>&1; echo "pytest=$? t=$(( $(date +%s) - T0 ))"
tail -25 gpurun_out/r2q_pytest.log | cut -c1-400
