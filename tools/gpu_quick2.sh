#!/bin/bash
# structured-solver iteration: parity tests, short bench with the symmetric and (A/B) the full-square window
set -u
TAG=${1:-q}
mkdir -p gpurun_out
SFB_BAND_LOOKAHEAD=1 timeout 300 python -m pytest tests/test_gpu_structured.py -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest (look-ahead) rc=$?"; tail -6 gpurun_out/${TAG}_pytest.log
for mode in base la; do
  if [ $mode = la ]; then export SFB_BAND_LOOKAHEAD=1; else unset SFB_BAND_LOOKAHEAD; fi
  timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_$mode.json 2> gpurun_out/${TAG}_bench_$mode.err; echo "bench $mode rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_$mode.json"))
s=d["structured"]
print("$mode: structured", s["value"], "ms", s["ms_per_step"], "model", s["e2e_model"]["value"], "diff", s["max_rel_diff_vs_dense_lnL"], s["window_classes"], "chol ms", s["roofline"]["ms"], "build ms", s["roofline"]["band_build"]["ms"])
PY
  tail -2 gpurun_out/${TAG}_bench_$mode.err
done
export SFB_BAND_LOOKAHEAD=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'band_chol_kernel' -c 3 \
  -o gpurun_out/${TAG}_band python bench.py --walkers 64 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-model > gpurun_out/${TAG}_ncu_band.log 2>&1; echo "ncu band rc=$?"
