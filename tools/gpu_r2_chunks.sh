#!/bin/bash
set -u
OUT=gpurun_out/${1:-r2w}; mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for CH in 8 4 2 16; do
  timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-cfg5 --e2e-chunk $CH > $OUT/cfg2_ch$CH.json 2>> $OUT/err.txt
  python - $OUT/cfg2_ch$CH.json $CH <<'PY'
import json, sys
l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("cfg2 chunk", sys.argv[2], "e2e", round(l["e2e"]["ms_per_step"], 3), "serial", round(l["e2e"]["ms_per_step_serial"], 3), l["e2e"]["host_wall_per_step"], l["e2e"]["graph_captures"])
PY
done
for CH in 4 2; do
  timeout 300 python bench.py --workload cfg4 --steps 20 --no-cpu-baseline --no-cfg5 --e2e-chunk $CH > $OUT/cfg4_ch$CH.json 2>> $OUT/err.txt
  python - $OUT/cfg4_ch$CH.json $CH <<'PY'
import json, sys
l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("cfg4 chunk", sys.argv[2], "e2e", round(l["e2e"]["ms_per_step"], 3), "serial", round(l["e2e"]["ms_per_step_serial"], 3), l["e2e"]["host_wall_per_step"], l["e2e"]["graph_captures"])
PY
done
tail -3 $OUT/err.txt
