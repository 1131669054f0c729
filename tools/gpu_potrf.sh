#!/bin/bash
# blocked potrf_diag: phase timing (experiments build prints it), parity tests, launch list
set -u
TAG=${1:-pd}
mkdir -p gpurun_out
SFB200_LIB=$PWD/starfish_b200/libsfb200_exp.so timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -s -k "cho_factor or stress" > gpurun_out/${TAG}_phase.log 2>&1; echo "phase rc=$?"; grep -E "potrf_diag2 cycles|passed|failed" gpurun_out/${TAG}_phase.log | sort | uniq -c | sort -rn | head -8
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py tests/test_gpu_ozaki.py tests/test_gpu_shared_factor.py -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'potrf|trsm' -c 300 --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --solver dense_i8 --walkers 32 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-model --no-structured --no-configs --no-alt --no-frozen > gpurun_out/${TAG}_l.log 2>&1; echo "ncu launches rc=$?"
python tools/ncu_summary.py launches gpurun_out/${TAG}_launches.csv 2>/dev/null | head -5
