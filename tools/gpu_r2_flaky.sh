#!/bin/bash
set -u
OUT=gpurun_out/${1:-r2u}; mkdir -p $OUT
for i in 1 2 3; do
  timeout 300 python bench.py --workload cfg4 --steps 20 --no-cpu-baseline --no-cfg5 > $OUT/cfg4_$i.json 2>> $OUT/err.txt
  python - $OUT/cfg4_$i.json <<'PY'
import json, sys
l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("cfg4 e2e", round(l["e2e"]["ms_per_step"], 3), "serial", round(l["e2e"]["ms_per_step_serial"], 3), l["e2e"]["host_wall_per_step"], l["e2e"]["graph_captures"])
PY
done
for i in 1 2; do
  timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-cfg5 > $OUT/cfg2_$i.json 2>> $OUT/err.txt
  python - $OUT/cfg2_$i.json <<'PY'
import json, sys
l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("cfg2 e2e", round(l["e2e"]["ms_per_step"], 3), "serial", round(l["e2e"]["ms_per_step_serial"], 3), l["e2e"]["host_wall_per_step"], l["e2e"]["graph_captures"])
PY
done
tail -3 $OUT/err.txt
