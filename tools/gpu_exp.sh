#!/bin/bash
# A/B on the GPU box: the committed band.cu first, then every tools/exp/*.cu.txt swapped in and rebuilt there.
set -u
TAG=${1:-exp}
mkdir -p gpurun_out
run() {
  local tag=$1
  timeout 600 python -m pytest tests/test_gpu_structured.py -q -x > gpurun_out/${tag}_pytest.log 2>&1; echo "$tag pytest rc=$?"; tail -2 gpurun_out/${tag}_pytest.log
  timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "$tag bench rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_bench.json")); s=d["structured"]
print("$tag: structured", s["value"], "ms", s["ms_per_step"], "model", s["e2e_model"]["value"], "diff", s["max_rel_diff_vs_dense_lnL"], "chol ms", s["roofline"]["ms"], "build ms", s["roofline"]["band_build"]["ms"])
PY
}
run ${TAG}_base
cp starfish_b200/csrc/band.cu /tmp/band_committed.cu
for f in tools/exp/*.cu.txt; do
  v=$(basename $f .cu.txt)
  cp $f starfish_b200/csrc/band.cu && python -m starfish_b200.build --force > gpurun_out/${TAG}_${v}_build.log 2>&1; echo "$v rebuild rc=$?"
  run ${TAG}_${v}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:'band_chol_kernel<\(int\)128' -c 1 \
    -o gpurun_out/${TAG}_${v} python bench.py --walkers 64 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-model > gpurun_out/${TAG}_${v}_ncu.log 2>&1; echo "$v ncu rc=$?"
done
cp /tmp/band_committed.cu starfish_b200/csrc/band.cu
