#!/bin/bash
# One GPU-box visit: parity tests (incl. structured solver), smoke, bench, ncu of the band kernels.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r01l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r01l_pytest.log
grep -E "variant|passed|failed|FAILED|Error|rc=" gpurun_out/r01l_pytest.log | tail -30
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r01l_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r01l_smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r01l_bench.json 2> gpurun_out/r01l_bench.err; echo "bench rc=$?"; cat gpurun_out/r01l_bench.json; tail -5 gpurun_out/r01l_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'band_chol_kernel|band_build_kernel' -c 8 \
  -o gpurun_out/r01l_band python bench.py --walkers 64 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-model > gpurun_out/r01l_ncu_band.log 2>&1; echo "ncu band rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gp_predict_kernel|resample_kernel|broaden_kernel|combine_kernel' -c 8 \
  -o gpurun_out/r01l_upstream python bench.py --walkers 64 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-structured > gpurun_out/r01l_ncu_up.log 2>&1; echo "ncu upstream rc=$?"
ls -la gpurun_out | tail -8
