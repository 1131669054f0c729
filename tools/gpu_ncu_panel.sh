#!/bin/bash
# ncu full capture of the panel kernels of the int8 mode: trsm_kernel<true> (with the slices) and potrf_diag2_kernel
set -u
TAG=${1:-pn}
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'trsm_kernel|potrf_diag2' -s 20 -c 4 \
  -o gpurun_out/${TAG}_panel python bench.py --solver dense_i8 --walkers 64 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-model --no-structured --no-configs --no-alt --no-frozen > gpurun_out/${TAG}_n.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/${TAG}_n.log
