#!/bin/bash
# ncu full capture of the int8 trailing-update kernel: the first big trailing-triangle launch (REST(0), K=1024) and
# one in-panel strip
set -u
TAG=${1:-i8n}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'syrk_i8' -s 7 -c 2 \
  -o gpurun_out/${TAG}_syrk_i8 python bench.py --solver dense_i8 --walkers 32 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-model --no-structured --no-configs --no-alt --no-frozen > gpurun_out/${TAG}_n.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/${TAG}_n.log
