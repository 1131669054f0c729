#!/bin/bash
# shared-factor tests + experiments build A/B of the int8 schedule knobs (tiles per CTA, outer block width, SS/TS)
set -u
TAG=${1:-exp}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_shared_factor.py tests/test_gpu_parity.py tests/test_gpu_ozaki.py -q -x -s > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "dlnL|passed|failed|Error|error" gpurun_out/${TAG}_pytest.log | tail -20
run_bench () {  # name, extra env
  timeout 600 env SFB200_LIB=$PWD/starfish_b200/libsfb200_exp.so $2 python bench.py --solver dense_i8 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e --no-model --no-structured --no-configs --no-alt --no-frozen > gpurun_out/${TAG}_bench_$1.json 2> gpurun_out/${TAG}_bench_$1.err; echo "bench $1 rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench_$1.json"))
    r = d["roofline"]
    print("$1", "evals/s", round(d["value"], 2), "ms/step", round(d["ms_per_step"], 1), "syrk TF/s-eq", round(r["achieved_fp64_equivalent_tflops"], 2),
          "clk", d["clocks"]["sm_mhz"], "W", d["clocks"]["power_w_max"], "kernels ms", {k: v["ms"] for k, v in d["kernels"].items()})
except Exception as e:
    print("$1 bench unreadable:", e)
PY
  tail -2 gpurun_out/${TAG}_bench_$1.err
}
run_bench base "SFB_NOP=1"
run_bench tpc4 "SFB_OZ_TPC=4"
run_bench tpc2 "SFB_OZ_TPC=2"
run_bench tpc16 "SFB_OZ_TPC=16"
run_bench ot16 "SFB_OUTER_TILES=16"
run_bench ot4 "SFB_OUTER_TILES=4"
run_bench nohi "SFB_DEBUG_MODE=1"
run_bench lanes3 "SFB_LANES=3"
run_bench lanes4 "SFB_LANES=4"
run_bench lanes1 "SFB_LANES=1"
