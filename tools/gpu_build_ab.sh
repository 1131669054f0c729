#!/bin/bash
# cov_build iteration: every GPU test, the write-only bandwidth of the box, a short dense bench, one ncu capture of cov_build.
set -u
TAG=${1:-b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
timeout 120 tools/write_peak > gpurun_out/${TAG}_write_peak.txt 2>&1; cat gpurun_out/${TAG}_write_peak.txt
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-model --no-structured > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print("dense", d["value"], "build", d["roofline"]["other_kernels"]["build"])
PY
tail -3 gpurun_out/${TAG}_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cov_build_kernel -c 1 \
  -o gpurun_out/${TAG}_cov_build python bench.py --walkers 48 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-model --no-structured > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"
