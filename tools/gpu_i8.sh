#!/bin/bash
# int8 (Ozaki) trailing update: parity tests, short bench of the int8 mode (and optionally the fp64 mode), ncu launch list
set -u
TAG=${1:-i8}
MODES=${2:-dense_i8}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ozaki.py -q -x -s > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/${TAG}_pytest.log
for S in $MODES; do
  timeout 600 python bench.py --solver $S --steps 3 --warmup 1 --no-cpu-baseline --no-e2e --no-model --no-structured > gpurun_out/${TAG}_bench_$S.json 2> gpurun_out/${TAG}_bench_$S.err; echo "bench $S rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench_$S.json"))
    r = d["roofline"]
    print("$S", "evals/s", round(d["value"], 2), "ms/step", round(d["ms_per_step"], 1), "syrk TFLOP/s(fp64-equiv)", round(r["achieved"], 2),
          "share", round(r["share_of_step"], 3), "clocks", d["clocks"], "other", json.dumps(r["other_kernels"]))
except Exception as e:
    print("$S bench unreadable:", e)
PY
  tail -3 gpurun_out/${TAG}_bench_$S.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --solver dense_i8 --walkers 32 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-model --no-structured > gpurun_out/${TAG}_l.log 2>&1; echo "ncu launches rc=$?"
python tools/ncu_summary.py launches gpurun_out/${TAG}_launches.csv | head -12
