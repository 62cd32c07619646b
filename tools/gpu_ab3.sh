#!/bin/bash
# same-box A/B of stage-1 variants: DEFT_EXPERIMENT values given as arguments after the tag
set -u
OUT=gpurun_out/${1:-ab}; shift; mkdir -p $OUT
EXPS="${@:-0 16}"
for W in cfg2 cfg4 cfg3; do
  for E in $EXPS; do
    DEFT_EXPERIMENT=$E timeout 300 python bench.py --workload $W --steps 20 --no-cpu-baseline --no-cfg5 --e2e-static > $OUT/bench_${W}_e$E.json 2>> $OUT/bench.err
    python - <<PY
import json
try:
    b=json.load(open("$OUT/bench_${W}_e$E.json"))
    print("$W exp=$E: call %.2f us  stage1 %.2f  stage2 %.2f  frac %.3f (%s)  hbm %.3f tensor %.3f" % (b["us_per_layer_call"], b["us_stage1"], b["us_stage2"], b["roofline"]["frac"], b["roofline"]["bound"], b["roofline"]["hbm_frac"], b["roofline"]["tensor_frac"]))
except Exception as e:
    print("$W exp=$E: FAILED", e)
PY
  done
done
for E in $EXPS; do
  DEFT_EXPERIMENT=$E timeout 300 python bench.py --trees-per-gpu 64 --steps 10 --no-cpu-baseline --e2e-static > $OUT/bench_forest64_e$E.json 2>> $OUT/bench.err
  python - <<PY
import json
try:
    b=json.load(open("$OUT/bench_forest64_e$E.json"))
    print("forest64 exp=$E: call %.2f us  stage1 %.2f  stage2 %.2f  frac %.3f  clocks %s" % (b["us_per_layer_call"], b["us_stage1"], b["us_stage2"], b["roofline"]["frac"], b["clocks"]))
except Exception as e:
    print("forest64 exp=$E: FAILED", e)
PY
done
tail -5 $OUT/bench.err
