#!/bin/bash
# Round 2, first GPU call: parity tests, the reference Triton probe, the new bench line (cfg2 + cfg5 block),
# and `ncu --set full` captures of stage 1 for cfg2, cfg4 and the 64-tree forest.
set -u
TAG=${1:-r2a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -x -q -m gpu > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest.log
echo "== probe"
timeout 1500 python tools/ref_triton_probe.py > $OUT/probe.log 2>&1; echo "probe rc=$?"; tail -30 $OUT/probe.log
cp gpurun_out/ref_probe.json $OUT/ 2>/dev/null
echo "== bench"
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -5 $OUT/bench.err
for W in cfg1 cfg3 cfg3b cfg4; do
  timeout 300 python bench.py --workload $W --steps 20 --no-cpu-baseline --no-cfg5 > $OUT/bench_$W.json 2>> $OUT/bench.err; cat $OUT/bench_$W.json
done
timeout 300 python bench.py --trees-per-gpu 64 --steps 10 --no-cpu-baseline > $OUT/bench_forest64.json 2>> $OUT/bench.err; cat $OUT/bench_forest64.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cfg5 > $OUT/ncu_launch_bench.log 2>&1; echo "ncu list rc=$?"
echo "== ncu full cfg2 / cfg4 / forest64"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stage1 -s 40 -c 3 -f -o $OUT/prof \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cfg5 > $OUT/ncu_full_bench.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stage1 -s 34 -c 2 -f -o $OUT/prof_cfg4 \
   python bench.py --workload cfg4 --steps 2 --warmup 3 --no-cpu-baseline --no-cfg5 > $OUT/ncu_full_cfg4.log 2>&1; echo "ncu cfg4 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stage1 -s 34 -c 2 -f -o $OUT/prof_forest64 \
   python bench.py --trees-per-gpu 64 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_forest.log 2>&1; echo "ncu forest rc=$?"
ls -la $OUT
