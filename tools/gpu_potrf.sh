#!/bin/bash
# blocked potrf_diag: parity tests, bench A/B against the one-sweep kernel (experiments build), launch list
set -u
TAG=${1:-pd}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py tests/test_gpu_ozaki.py tests/test_gpu_shared_factor.py -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
run_bench () {  # name, extra env, solver
  timeout 600 env SFB200_LIB=$PWD/starfish_b200/libsfb200_exp.so $2 python bench.py --solver $3 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e --no-model --no-structured --no-configs --no-alt --no-frozen > gpurun_out/${TAG}_bench_$1.json 2> gpurun_out/${TAG}_bench_$1.err; echo "bench $1 rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench_$1.json"))
    print("$1", "evals/s", round(d["value"], 2), "ms/step", round(d["ms_per_step"], 1), "clk", d["clocks"]["sm_mhz"], "kernels ms", {k: v["ms"] for k, v in d["kernels"].items()})
except Exception as e:
    print("$1 bench unreadable:", e)
PY
  tail -2 gpurun_out/${TAG}_bench_$1.err
}
run_bench i8_blocked "SFB_POTRF_BLOCKED=1" dense_i8
run_bench i8_onesweep "SFB_POTRF_BLOCKED=0" dense_i8
run_bench f64_blocked "SFB_POTRF_BLOCKED=1" dense
run_bench f64_onesweep "SFB_POTRF_BLOCKED=0" dense
timeout 300 python -m pytest tests/test_gpu_fullsize.py -q -x > gpurun_out/${TAG}_pytest_full.log 2>&1; echo "pytest fullsize rc=$?"; tail -3 gpurun_out/${TAG}_pytest_full.log
