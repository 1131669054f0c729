#!/bin/bash
# int8 (Ozaki) trailing update: parity tests, short bench of the int8 mode, launch list
set -u
TAG=${1:-i8}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ozaki.py -q -x -s > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/${TAG}_pytest.log
run_bench () {  # name, extra env
  timeout 600 env $2 python bench.py --solver dense_i8 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e --no-model --no-structured --no-configs --no-alt --no-frozen > gpurun_out/${TAG}_bench_$1.json 2> gpurun_out/${TAG}_bench_$1.err; echo "bench $1 rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench_$1.json"))
    r = d["roofline"]
    print("$1", "evals/s", round(d["value"], 2), "ms/step", round(d["ms_per_step"], 1), "syrk TFLOP/s(fp64-equiv)", round(r["achieved_fp64_equivalent_tflops"], 2),
          "share", round(r["share_of_step"], 3), "issued_frac", r.get("int8_mma_issued_frac"), "clocks", d["clocks"], "kernels", json.dumps(d["kernels"]))
except Exception as e:
    print("$1 bench unreadable:", e)
PY
  tail -3 gpurun_out/${TAG}_bench_$1.err
}
run_bench default "SFB_NOP=1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --solver dense_i8 --walkers 32 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-model --no-structured --no-configs --no-alt --no-frozen > gpurun_out/${TAG}_l.log 2>&1; echo "ncu launches rc=$?"
python tools/ncu_summary.py launches gpurun_out/${TAG}_launches.csv | head -9
