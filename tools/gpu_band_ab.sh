#!/bin/bash
# A/B of the band kernels in one box visit: default (rank-4 DMMA, unrolled) vs SFB_BAND_MMA_ROLLED=1 (one copy of the
# block body) vs SFB_BAND_MMA_LA=1 (panel look-ahead) vs SFB_BAND_RANK1=1.  The tests are run with -x per variant; a
# variant that fails them is simply dropped.  For every variant: the structured parity tests, a short bench line (structured
# leg only matters), and the per-class kernel durations from an ncu launch list.
#   gpurun --timeout 600 -- 'bash tools/gpu_band_ab.sh r03a'
set -u
TAG=${1:-ab}
mkdir -p gpurun_out
for V in default rolled rank1 la; do
  case $V in
    default) unset SFB_BAND_MMA_ROLLED SFB_BAND_MMA_LA SFB_BAND_RANK1 ;;
    rolled)  unset SFB_BAND_MMA_LA SFB_BAND_RANK1; export SFB_BAND_MMA_ROLLED=1 ;;
    la)      unset SFB_BAND_MMA_ROLLED SFB_BAND_RANK1; export SFB_BAND_MMA_LA=1 ;;
    rank1)   unset SFB_BAND_MMA_ROLLED SFB_BAND_MMA_LA; export SFB_BAND_RANK1=1 ;;
  esac
  timeout 200 python -m pytest tests/test_gpu_structured.py -q -x > gpurun_out/${TAG}_${V}_pytest.log 2>&1; echo "$V pytest rc=$?"; tail -1 gpurun_out/${TAG}_${V}_pytest.log
  timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-e2e --no-model > gpurun_out/${TAG}_${V}_bench.json 2> gpurun_out/${TAG}_${V}_bench.err; echo "$V bench rc=$?"
  python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_${V}_bench.json")); s = d["structured"]
print("$V", "structured evals/s", round(s["value"]), "ms/step", round(s["ms_per_step"], 2), "chol ms (serial)", round(s["roofline"]["ms"], 2),
      "max rel diff vs dense", s["max_rel_diff_vs_dense_lnL"], "not PD", s["not_positive_definite"])
PY
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"band_" -c 40 --csv \
    --log-file gpurun_out/${TAG}_${V}_band_launches.csv python bench.py --walkers 64 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-model > /dev/null 2>&1
  python tools/ncu_summary.py launches gpurun_out/${TAG}_${V}_band_launches.csv | head -8
done
