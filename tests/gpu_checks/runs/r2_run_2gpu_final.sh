#!/bin/bash
# Final 2-GPU pass: NCCL data-parallel parity (ranks bit-identical) + bench at N = 2 on the final kernels.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "two_gpus or two_rank" > gpurun_out/r2aj_pytest_2gpu.log 2>&1; echo "pytest2=$? t=$(( $(date +%s) - T0 ))"
tail -3 gpurun_out/r2aj_pytest_2gpu.log | cut -c1-300; cut -c1-600 gpurun_out/dp_check_w2.json 2>/dev/null
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 --no-kernel-rooflines > gpurun_out/r2aj_bench_2gpu.log 2>&1; echo "bench2=$? t=$(( $(date +%s) - T0 ))"
tail -1 gpurun_out/r2aj_bench_2gpu.log > gpurun_out/r2aj_bench_2gpu.json; python -c "
import json; d=json.load(open('gpurun_out/r2aj_bench_2gpu.json')); print(d['n_gpus'], d['ms_per_step'], d['value'], d['config']['workload'][:60], d.get('allreduce'))" 2>&1 | cut -c1-400
