#!/bin/bash
# Pipelined end-to-end step (DecodeStepPipeline): GPU parity suite + the bench line + N = 1 cfg3/cfg4 lines.
set -u
TAG=${1:-r2p}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
echo "== bench"
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -3 $OUT/bench.err
for W in cfg3 cfg4; do
  timeout 300 python bench.py --workload $W --steps 20 --no-cpu-baseline --no-cfg5 > $OUT/bench_$W.json 2>> $OUT/bench.err; cat $OUT/bench_$W.json
done
timeout 300 python bench.py --mode node --steps 20 --no-cpu-baseline --no-cfg5 > $OUT/bench_node.json 2>> $OUT/bench.err; cat $OUT/bench_node.json
