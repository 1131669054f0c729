#!/bin/bash
# every GPU test, smoke, the default bench line and the reference arm
set -u
TAG=${1:-full}
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x --timeout 120 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 400 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "model", d["e2e_model"]["value"], "parity", d["parity"])
print("roofline", {k: d["roofline"][k] for k in ("achieved", "peak", "frac", "achieved_fp64_equivalent_tflops", "share_of_step") if k in d["roofline"]})
print("kernels", d["kernels"])
alt = d.get("fp64_dmma") or d.get("int8_tensor")
print("alt", alt["solver"], alt["value"], alt["max_rel_diff_vs_headline_lnL"])
print("structured", d["structured"]["value"], d["structured"]["max_rel_diff_vs_dense_lnL"])
for k, c in (d.get("configs") or {}).items():
    for s in ("dense_i8", "dense"):
        print(k, s, round(c[s]["value"], 2), "frac of fp64 floor", round(c[s]["frac_of_fp64_floor"], 3), c[s].get("kernels"), c[s].get("max_rel_diff_vs_dense_i8"))
print("cpu", d["cpu_baseline"]["sample"])
print("clocks", d["clocks"])
PY
