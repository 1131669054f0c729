#!/bin/bash
# panel-chain kernels (potrf_diag blocked, balanced trsm): parity tests, ncu launch list, short bench
set -u
TAG=${1:-pn}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py tests/test_gpu_ozaki.py tests/test_gpu_shared_factor.py tests/test_gpu_fullsize.py -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --solver dense_i8 --walkers 32 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-model --no-structured --no-configs --no-alt --no-frozen > gpurun_out/${TAG}_l.log 2>&1; echo "ncu launches rc=$?"
python tools/ncu_summary.py launches gpurun_out/${TAG}_launches.csv 2>/dev/null | head -9
for S in dense_i8 dense; do
timeout 600 python bench.py --solver $S --steps 3 --warmup 1 --no-cpu-baseline --no-e2e --no-model --no-structured --no-configs --no-alt --no-frozen > gpurun_out/${TAG}_bench_$S.json 2> gpurun_out/${TAG}_bench_$S.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench_$S.json"))
print("$S", "evals/s", round(d["value"], 2), "ms/step", round(d["ms_per_step"], 1), "clk", d["clocks"]["sm_mhz"], "kernels ms", {k: v["ms"] for k, v in d["kernels"].items()})
PY
done
