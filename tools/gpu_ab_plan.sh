#!/bin/bash
# Same-box sweep of the plan's cost-model constants: bash tools/gpu_ab_plan.sh <tag> <workload> ["gather costs"] ["job consts"]
set -u
TAG=$1; W=${2:-cfg2}
GS=${3:-"1.2 1.5 1.8"}; CS=${4:-"1.0 2.0 3.0"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for G in $GS; do
  for C in $CS; do
    DEFT_PLAN_GATHER_COST=$G DEFT_PLAN_JOB_CONST=$C timeout 300 python bench.py --workload $W --steps 30 --no-cpu-baseline 2>> $OUT/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$W gather=$G const=$C: layer-call %.2f us, stage1 %.2f, stage2 %.2f' % (d['us_per_layer_call'], d['us_stage1'], d['us_stage2']))"
  done
done
