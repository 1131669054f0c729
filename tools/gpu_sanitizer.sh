#!/bin/bash
# compute-sanitizer passes (memcheck, racecheck, synccheck) on the small GPU tests, incl. the tcgen05 / cluster kernel
set -u
TAG=${1:-san}
mkdir -p gpurun_out
SEL='n256 or not_positive or cho_factor or odd_sizes or ragged or 384'
timeout 900 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py tests/test_gpu_ozaki.py -q -x -k "$SEL" > gpurun_out/${TAG}_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|Hazard" gpurun_out/${TAG}_memcheck.log | head -20
timeout 900 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py tests/test_gpu_ozaki.py -q -x -k "$SEL" > gpurun_out/${TAG}_racecheck.log 2>&1; echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|passed|failed|hazard|Hazard" gpurun_out/${TAG}_racecheck.log | head -20
timeout 600 compute-sanitizer --tool synccheck --print-limit 10 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_parity.py -q -x -k "$SEL" > gpurun_out/${TAG}_synccheck.log 2>&1; echo "synccheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Barrier|barrier" gpurun_out/${TAG}_synccheck.log | head -20
