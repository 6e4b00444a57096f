#!/bin/bash
# Round-2 GPU pass 34 (last GPU seconds of the round): `train.precision: float32` — fp32 master adapters
# (leco_optim_flat_master, FlatState.master): the kernel case, the trainer iteration, the drop-in loop, the driver; smoke.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
timeout 100 python -m pytest tests -m gpu -q -p no:cacheprovider -x \
  -k "optimizers_master or float32" > gpurun_out/r2al_pytest.log 2>&1; echo "pytest=$? t=$(( $(date +%s) - T0 ))"
tail -25 gpurun_out/r2al_pytest.log | cut -c1-400
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2al_smoke.log 2>&1; echo "smoke=$? t=$(( $(date +%s) - T0 ))"
tail -2 gpurun_out/r2al_smoke.log | cut -c1-400
