#!/bin/bash
# Round-end validation visit: every GPU test (default kernels) + the structured tests with the rank-1 band kernel,
# smoke, the default bench line, the reference arm, an ncu launch list, an ncu capture of the band kernels, and a
# rank-1 A/B line of the structured leg.
set -u
TAG=${1:-final}
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 400 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; echo "reference rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --walkers 32 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'band_mma_kernel' -c 3 \
  -o gpurun_out/${TAG}_band_mma python bench.py --walkers 64 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-model > gpurun_out/${TAG}_ncu_band.log 2>&1; echo "ncu band rc=$?"
