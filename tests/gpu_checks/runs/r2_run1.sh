#!/bin/bash
# Round-2 first GPU pass (one B200, through gpurun from the repo root): parity suite incl. the full-size cases, smoke,
# every bench row, the CPU reference arm and the launch list of one iteration.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2a_pytest.log 2>&1; echo "pytest=$? t=$(( $(date +%s) - T0 ))"
tail -30 gpurun_out/r2a_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo "smoke=$? t=$(( $(date +%s) - T0 ))"
tail -3 gpurun_out/r2a_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.log 2>&1; echo "bench=$? t=$(( $(date +%s) - T0 ))"
tail -1 gpurun_out/r2a_bench.log > gpurun_out/r2a_bench.json
for c in sd21_b4 sd15_c3lier sdxl; do
  timeout 600 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_$c.log 2>&1; echo "bench_$c=$? t=$(( $(date +%s) - T0 ))"
  tail -1 gpurun_out/r2a_bench_$c.log > gpurun_out/r2a_bench_$c.json
  tail -c 600 gpurun_out/r2a_bench_$c.log
done
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2a_bench_ref.log 2>&1; echo "ref=$? t=$(( $(date +%s) - T0 ))"
tail -1 gpurun_out/r2a_bench_ref.log > gpurun_out/r2a_bench_ref.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r2a_launches.csv python tests/gpu_checks/profile_step.py --k 1 > gpurun_out/r2a_prof.log 2>&1
echo "launches=$? t=$(( $(date +%s) - T0 ))"
cat gpurun_out/r2a_bench.json
cat gpurun_out/r2a_bench_ref.json
