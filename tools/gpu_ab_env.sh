#!/bin/bash
# Same-box A/B over an environment variable: bash tools/gpu_ab_env.sh <tag> <VAR> <value a> <value b> ...
set -u
TAG=$1; VAR=$2; shift; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
for rep in 1 2; do
  for F in "$@"; do
    env $VAR=$F timeout 300 python bench.py --steps 30 --no-cpu-baseline > $OUT/bench_${F}_$rep.json 2>> $OUT/bench.err
    python - <<PY
import json
d = json.load(open("$OUT/bench_${F}_$rep.json"))
print("$VAR=$F rep $rep: layer-call %.2f us, stage1 %.2f, stage2 %.2f" % (d["us_per_layer_call"], d["us_stage1"], d["us_stage2"]))
PY
  done
done
