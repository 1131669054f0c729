#!/bin/bash
# int8 (Ozaki) trailing update: probe, parity tests, short bench A/B against the fp64 DMMA mode.
set -u
TAG=${1:-i8}
mkdir -p gpurun_out
timeout 120 ./tools/exp/umma_i8_probe > gpurun_out/${TAG}_probe.txt 2>&1; echo "probe rc=$?"; tail -8 gpurun_out/${TAG}_probe.txt
timeout 900 python -m pytest tests/test_gpu_ozaki.py -q -x -s > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/${TAG}_pytest.log
for S in dense_i8 dense; do
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
