#!/bin/bash
# ncu capture of the int8 trailing-update kernel: launch list + one full capture of the first big trailing update
set -u
TAG=${1:-i8n}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'syrk_i8|oz_' -c 400 --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --solver dense_i8 --walkers 8 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-model --no-structured > gpurun_out/${TAG}_l.log 2>&1; echo "ncu launches rc=$?"
python tools/ncu_summary.py launches gpurun_out/${TAG}_launches.csv | head -12
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'syrk_i8_kernel' -s 8 -c 1 \
  -o gpurun_out/${TAG}_syrk_i8 python bench.py --solver dense_i8 --walkers 8 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-model --no-structured > gpurun_out/${TAG}_n.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/${TAG}_n.log
