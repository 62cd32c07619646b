#!/bin/bash
# Kernel iteration round: parity tests, then same-box A/B of the stage-1 generations (DEFT_EXPERIMENT=16 = previous one)
set -u
TAG=${1:-r2b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest.log
for W in cfg2 cfg4 cfg3 cfg1; do
  for E in 0 16; do
    DEFT_EXPERIMENT=$E timeout 300 python bench.py --workload $W --steps 20 --no-cpu-baseline --no-cfg5 --e2e-static > $OUT/bench_${W}_e$E.json 2>> $OUT/bench.err
    python - <<PY
import json
try:
    b=json.load(open("$OUT/bench_${W}_e$E.json"))
    print("$W exp=$E: call %.2f us  stage1 %.2f  stage2 %.2f  frac %.3f (%s)  e2e %.3f ms" % (b["us_per_layer_call"], b["us_stage1"], b["us_stage2"], b["roofline"]["frac"], b["roofline"]["bound"], b["e2e"]["ms_per_step"]))
except Exception as e:
    print("$W exp=$E: FAILED", e)
PY
  done
done
for E in 0 16; do
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
TRACE_TABLE= timeout 120 python tools/trace_stage1.py cfg2 2 > $OUT/trace_cfg2.txt 2>&1; tail -60 $OUT/trace_cfg2.txt
