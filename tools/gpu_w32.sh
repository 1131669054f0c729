#!/bin/bash
# per-GPU share of the 8-GPU configuration (32 walkers per GPU) on one GPU: predicts the scaling efficiency; lane count A/B
set -u
TAG=${1:-w32}
mkdir -p gpurun_out
LIB=$PWD/starfish_b200/libsfb200_exp.so
run () {  # name, walkers, env...
  name=$1; w=$2; shift; shift
  timeout 300 env SFB200_LIB=$LIB "$@" python bench.py --solver dense_i8 --walkers $w --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-model --no-structured --no-configs --no-alt --no-frozen > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/${TAG}_$name.json') if l.startswith('{')][-1]); print('$name', 'walkers', $w, 'evals/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), 'kernels', {k: v['ms'] for k, v in d['kernels'].items()}, 'clk', d['clocks']['sm_mhz'])"
}
run w32_l2 32 SFB_LANES=2
run w32_l4 32 SFB_LANES=4
run w32_l3 32 SFB_LANES=3
run w64_l2 64 SFB_LANES=2
run w64_l4 64 SFB_LANES=4
