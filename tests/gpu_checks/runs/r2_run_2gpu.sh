#!/bin/bash
# 2-GPU pass: NCCL data-parallel parity (LecoTrainer's DP branch) + bench at N = 2 (configs[4]-style batch 4 per GPU).
set -u
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "two_gpus" > gpurun_out/r2n_pytest_2gpu.log 2>&1; echo "pytest2=$? t=$(( $(date +%s) - T0 ))"
tail -5 gpurun_out/r2n_pytest_2gpu.log | cut -c1-500; cat gpurun_out/dp_check_w2.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2n_bench_2gpu.log 2>&1; echo "bench2=$? t=$(( $(date +%s) - T0 ))"
tail -1 gpurun_out/r2n_bench_2gpu.log > gpurun_out/r2n_bench_2gpu.json; cut -c1-1500 gpurun_out/r2n_bench_2gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --config sd21 --steps 5 --warmup 3 --no-kernel-rooflines > gpurun_out/r2n_bench_2gpu_b2.log 2>&1; echo "bench2_b2=$? t=$(( $(date +%s) - T0 ))"
tail -1 gpurun_out/r2n_bench_2gpu_b2.log > gpurun_out/r2n_bench_2gpu_b2.json; cut -c1-600 gpurun_out/r2n_bench_2gpu_b2.json
