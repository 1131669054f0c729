#!/bin/bash
# multi-GPU validation (run with gpurun --gpus 2): NCCL through the C ABI (two ranks, no torch.distributed) and the
# bench under torchrun
set -u
TAG=${1:-mg}
NG=${2:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_gpu_comm.py -q -s > gpurun_out/${TAG}_pytest_comm.log 2>&1; echo "pytest comm rc=$?"; tail -4 gpurun_out/${TAG}_pytest_comm.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $NG --steps 3 --warmup 3 --no-structured --no-frozen > gpurun_out/${TAG}_bench_${NG}gpu.json 2> gpurun_out/${TAG}_bench_${NG}gpu.err; echo "bench ${NG}gpu rc=$?"; tail -5 gpurun_out/${TAG}_bench_${NG}gpu.err
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench_${NG}gpu.json"))
print("n_gpus", d["n_gpus"], "value", d["value"], "e2e", d["e2e"]["value"], "model", d["e2e_model"]["value"])
alt = d.get("fp64_dmma"); print("fp64", alt["value"], alt["max_rel_diff_vs_headline_lnL"])
for k, c in (d.get("configs") or {}).items():
    print(k, {s: round(c[s]["value"], 2) for s in ("dense_i8", "dense")})
print("clocks", d["clocks"])
PY
