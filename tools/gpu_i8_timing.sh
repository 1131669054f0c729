#!/bin/bash
# experiments build: where the int8 update's MMA warp spends its time (SFB_OZ_TIMING), for a few schedule knobs
set -u
TAG=${1:-i8t}
mkdir -p gpurun_out
LIB=$PWD/starfish_b200/libsfb200_exp.so
run () {  # name, env...
  name=$1; shift
  timeout 300 env SFB200_LIB=$LIB SFB_OZ_TIMING=1 "$@" python bench.py --solver dense_i8 --walkers 64 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-model --no-structured --no-configs --no-alt --no-frozen > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err
  echo "== $name rc=$?"; grep "int8 MMA warp" gpurun_out/${TAG}_$name.err | tail -2
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/${TAG}_$name.json') if l.startswith('{')][-1]); print(' evals/s', round(d['value'],1), 'syrk ms', d['kernels']['syrk']['ms'], 'clk', d['clocks']['sm_mhz'])"
}
shift
for v in "$@"; do
  run $(echo $v | tr '= ' '__') $v
done
