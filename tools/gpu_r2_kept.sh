#!/bin/bash
set -u
TAG=${1:-r2s}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
for W in cfg2 cfg4 cfg3; do PYTHONPATH=. timeout 200 python tools/kept_order_ab.py $W 60 2>&1 | tail -3 | tee -a $OUT/kept_order_ab.txt; done
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -3 $OUT/bench.err
python - $OUT/bench.json <<'PY'
import json, sys
l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: l[k] for k in ("value", "ms_per_step", "us_stage1")}, "e2e", l["e2e"]["ms_per_step"], l["e2e"]["ms_per_step_serial"], l["e2e"]["graph_captures"])
print("cfg5", {k: l["cfg5"][k] for k in ("trees_per_s", "us_stage1", "roofline_frac_stage1")}, l["cfg5"]["e2e"]["ms_per_step"], l["cfg5"]["e2e"]["trees_per_s"])
PY
timeout 300 python bench.py --workload cfg4 --steps 20 --no-cpu-baseline --no-cfg5 > $OUT/bench_cfg4.json 2>> $OUT/bench.err
python - $OUT/bench_cfg4.json <<'PY'
import json, sys
l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("cfg4", {k: l[k] for k in ("value", "ms_per_step", "us_stage1")}, "e2e", l["e2e"]["ms_per_step"], l["e2e"]["ms_per_step_serial"], l["e2e"]["graph_captures"])
PY
