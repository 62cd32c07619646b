#!/bin/bash
set -u
OUT=gpurun_out/${1:-r2d}; mkdir -p $OUT
TRACE_JOBS=1 TRACE_TREES=64 timeout 200 python tools/trace_stage1.py cfg2 1 > $OUT/jobs_forest64.txt 2>&1; tail -50 $OUT/jobs_forest64.txt
TRACE_JOBS=1 timeout 200 python tools/trace_stage1.py cfg2 1 > $OUT/jobs_cfg2.txt 2>&1; tail -12 $OUT/jobs_cfg2.txt
TRACE_JOBS=1 timeout 200 python tools/trace_stage1.py cfg4 1 > $OUT/jobs_cfg4.txt 2>&1; tail -12 $OUT/jobs_cfg4.txt
timeout 600 ncu --set full --clock-control none -k regex:stage1 -s 34 -c 1 -f -o $OUT/prof_forest64 \
   python bench.py --trees-per-gpu 64 --steps 2 --warmup 3 --no-cpu-baseline --e2e-static > $OUT/ncu_full_forest.log 2>&1; echo "ncu forest rc=$?"
