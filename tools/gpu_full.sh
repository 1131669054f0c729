#!/bin/bash
# Full single-GPU visit: every GPU test, smoke, the default bench line, the reference arm, an ncu launch list.
set -u
TAG=${1:-full}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; echo "reference rc=$?"; cat gpurun_out/${TAG}_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --walkers 32 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
