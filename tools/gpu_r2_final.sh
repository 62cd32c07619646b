#!/bin/bash
# Round-2 closing GPU call: parity tests, smoke, the bench line (cfg2 + cfg5 block), the other workloads, ncu launch
# list and `--set full` captures of stage 1 (cfg2, cfg4, 64-tree forest), per-job trace, e2e breakdown.
set -u
TAG=${1:-r2z}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee $OUT/smoke.log | tail -2
echo "== bench"
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; grep -i "elapsed\|error" $OUT/bench.err | tail -3
for W in cfg1 cfg3 cfg3b cfg4; do
  timeout 300 python bench.py --workload $W --steps 20 --no-cpu-baseline --no-cfg5 > $OUT/bench_$W.json 2>> $OUT/bench.err; cat $OUT/bench_$W.json
done
for W in cfg2 cfg4; do
  DEFT_PLAN_REGROUP=0 timeout 300 python bench.py --workload $W --steps 20 --no-cpu-baseline --no-cfg5 --e2e-static > $OUT/bench_${W}_noregroup.json 2>> $OUT/bench.err; cat $OUT/bench_${W}_noregroup.json
done
DEFT_STAGE1_IMPL=1 timeout 300 python bench.py --workload cfg1 --steps 20 --no-cpu-baseline --no-cfg5 --e2e-static > $OUT/bench_cfg1_fma.json 2>> $OUT/bench.err; cat $OUT/bench_cfg1_fma.json
for M in node node_chunk seq; do
  timeout 300 python bench.py --mode $M --steps 20 --no-cpu-baseline --no-cfg5 > $OUT/bench_$M.json 2>> $OUT/bench.err; cat $OUT/bench_$M.json
done
timeout 300 python bench.py --trees-per-gpu 64 --steps 10 --no-cpu-baseline > $OUT/bench_forest64.json 2>> $OUT/bench.err; cat $OUT/bench_forest64.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2>> $OUT/bench.err; cat $OUT/bench_reference.json
timeout 120 python tools/e2e_breakdown.py > $OUT/e2e_breakdown.txt 2>&1; tail -4 $OUT/e2e_breakdown.txt
META_PARTS=1 timeout 120 python tools/e2e_breakdown.py > $OUT/e2e_meta_parts.txt 2>&1; tail -1 $OUT/e2e_meta_parts.txt
TRACE_JOBS=1 TRACE_TREES=64 timeout 200 python tools/trace_stage1.py cfg2 1 > $OUT/jobs_forest64.txt 2>&1; grep -v "^cta" $OUT/jobs_forest64.txt | tail -5
TRACE_JOBS=1 timeout 200 python tools/trace_stage1.py cfg2 1 > $OUT/jobs_cfg2.txt 2>&1; grep -v "^cta" $OUT/jobs_cfg2.txt | tail -4
timeout 200 python tools/trace_stage1.py cfg2 2 > $OUT/trace_cfg2.txt 2>&1; head -3 $OUT/trace_cfg2.txt
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cfg5 > $OUT/ncu_launch_bench.log 2>&1; echo "ncu list rc=$?"
echo "== ncu full cfg2 / cfg4 / forest64"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stage1 -s 40 -c 3 -f -o $OUT/prof \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cfg5 > $OUT/ncu_full_bench.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:stage1 -s 34 -c 2 -f -o $OUT/prof_cfg4 \
   python bench.py --workload cfg4 --steps 2 --warmup 3 --no-cpu-baseline --no-cfg5 > $OUT/ncu_full_cfg4.log 2>&1; echo "ncu cfg4 rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:stage1 -s 34 -c 2 -f -o $OUT/prof_forest64 \
   python bench.py --trees-per-gpu 64 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_forest.log 2>&1; echo "ncu forest rc=$?"
ls $OUT
