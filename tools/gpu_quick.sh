#!/bin/bash
# Quick GPU check: the tcgen05 unit tests first (short timeout: a hang must not eat the box), then everything.
set -u
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 240 python -m pytest tests/test_umma_gpu.py -x -q -s -m gpu > $OUT/umma.log 2>&1; echo "umma rc=$?"
tail -40 $OUT/umma.log
timeout 600 python -m pytest tests -x -q -m gpu > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest.log
timeout 300 python bench.py --steps 20 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -5 $OUT/bench.err
