#!/bin/bash
# Host-side profile of the pipelined end-to-end step + GPU suite after the native tree mirror.
set -u
TAG=${1:-r2r}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cfg5 --no-cpu-baseline --profile-e2e $OUT/e2e_profile.txt > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -3 $OUT/bench.err
python - $OUT/bench.json <<'PY'
import json, sys
l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: l[k] for k in ("value", "ms_per_step", "us_stage1")}, "e2e", l["e2e"]["ms_per_step"], l["e2e"]["ms_per_step_serial"], l["e2e"]["graph_captures"])
PY
head -60 $OUT/e2e_profile.txt
DEFT_NATIVE_TREE=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cfg5 --no-cpu-baseline > $OUT/bench_nomirror.json 2>> $OUT/bench.err
python - $OUT/bench_nomirror.json <<'PY'
import json, sys
l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("no mirror: e2e", l["e2e"]["ms_per_step"], l["e2e"]["ms_per_step_serial"])
PY
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_cfg5.json 2>> $OUT/bench.err
python - $OUT/bench_cfg5.json <<'PY'
import json, sys
l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("cfg5 e2e", l["cfg5"]["e2e"]["ms_per_step"], l["cfg5"]["e2e"]["trees_per_s"])
PY
