#!/bin/bash
# Quick structured-solver iteration: its parity tests, a short bench, one ncu capture of the band kernels.
set -u
TAG=${1:-q}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_structured.py -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print("dense", d["value"], "model", d["e2e_model"]["value"], "upstream ms", d["e2e_model"].get("upstream_ms_per_step"))
print("structured", json.dumps(d["structured"])[:1500])
PY
tail -3 gpurun_out/${TAG}_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'band_chol_kernel|band_build_kernel' -c 6 \
  -o gpurun_out/${TAG}_band python bench.py --walkers 64 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-model > gpurun_out/${TAG}_ncu_band.log 2>&1; echo "ncu band rc=$?"
