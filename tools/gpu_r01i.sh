#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list + full capture of the upstream kernels.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/r01i_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/r01i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r01i_pytest.log
tail -25 gpurun_out/r01i_pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r01i_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r01i_smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r01i_bench.json 2> gpurun_out/r01i_bench.err; echo "bench rc=$?"; cat gpurun_out/r01i_bench.json; tail -5 gpurun_out/r01i_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r01i_launches.csv \
  python bench.py --walkers 32 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r01i_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'resample_kernel|broaden_kernel|gp_predict_kernel|combine_kernel|rot_transfer_kernel' -c 10 \
  -o gpurun_out/r01i_upstream python bench.py --walkers 32 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r01i_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -12
