#!/bin/bash
# same-box A/B of plan variants (env settings given as "NAME=VALUE" arguments, "-" = defaults)
set -u
OUT=gpurun_out/${1:-ab}; shift; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
for V in "$@"; do
  for W in cfg2 cfg4 cfg3 cfg3b; do
    env ${V/-/X=0} timeout 300 python bench.py --workload $W --steps 20 --no-cpu-baseline --no-cfg5 --e2e-static > $OUT/b.json 2>> $OUT/bench.err
    python - <<PY
import json
try:
    b=json.load(open("$OUT/b.json"))
    print("$W [$V]: call %.2f us  stage1 %.2f  stage2 %.2f  hbm %.3f tensor %.3f" % (b["us_per_layer_call"], b["us_stage1"], b["us_stage2"], b["roofline"]["hbm_frac"], b["roofline"]["tensor_frac"]))
except Exception as e:
    print("$W [$V]: FAILED", e)
PY
  done
  env ${V/-/X=0} timeout 300 python bench.py --trees-per-gpu 64 --steps 10 --no-cpu-baseline --e2e-static > $OUT/b.json 2>> $OUT/bench.err
  python - <<PY
import json
try:
    b=json.load(open("$OUT/b.json"))
    print("forest64 [$V]: call %.2f us  stage1 %.2f  stage2 %.2f  frac %.3f  clocks %s" % (b["us_per_layer_call"], b["us_stage1"], b["us_stage2"], b["roofline"]["frac"], b["clocks"]))
except Exception as e:
    print("forest64 [$V]: FAILED", e)
PY
done
tail -5 $OUT/bench.err
