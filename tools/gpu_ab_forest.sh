#!/bin/bash
# Same-box A/B on the forest (batched trees) workload: bash tools/gpu_ab_forest.sh <tag> <VAR> <values...>
set -u
TAG=$1; VAR=$2; shift; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
for F in "$@"; do
  env $VAR=$F timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --trees-per-gpu ${TREES:-32} > $OUT/bench_$F.json 2>> $OUT/bench.err
  python - <<PY
import json
d = json.load(open("$OUT/bench_$F.json"))
print("$VAR=$F: step %.2f ms, per tree per layer %.2f us, stage1 %.1f us, stage2 %.1f, frac %.3f, e2e %.1f ms" % (d["ms_per_step"], d["us_per_layer_call"] / d["config"]["trees_per_gpu"], d["us_stage1"], d["us_stage2"], d["roofline"]["frac"], d["e2e"]["ms_per_step"]))
PY
done
