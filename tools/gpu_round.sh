#!/bin/bash
# One GPU-box round: parity tests, smoke, bench, ncu launch list, ncu full capture of the top kernel.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh <tag> [kernel-regex]
set -u
TAG=${1:-r1}
KREGEX=${2:-stage1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu" | tee $OUT/pytest.log
timeout 900 python -m pytest tests -x -q -m gpu >> $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest.log
tail -5 $OUT/pytest.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee $OUT/smoke.log | tail -3
echo "== bench"
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -3 $OUT/bench.err
for W in cfg1 cfg3 cfg4; do
  timeout 300 python bench.py --workload $W --steps 20 --no-cpu-baseline > $OUT/bench_$W.json 2>> $OUT/bench.err; cat $OUT/bench_$W.json
done
for M in node node_chunk; do
  timeout 300 python bench.py --mode $M --steps 20 --no-cpu-baseline > $OUT/bench_$M.json 2>> $OUT/bench.err; cat $OUT/bench_$M.json
done
timeout 300 python bench.py --trees-per-gpu 64 --steps 10 --no-cpu-baseline > $OUT/bench_forest64.json 2>> $OUT/bench.err; cat $OUT/bench_forest64.json
timeout 300 python bench.py --mode seq --steps 10 --no-cpu-baseline > $OUT/bench_seq.json 2>> $OUT/bench.err; cat $OUT/bench_seq.json
timeout 120 python tools/e2e_breakdown.py > $OUT/e2e_breakdown.txt 2>&1; tail -2 $OUT/e2e_breakdown.txt
echo "== bench --impl reference"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2>> $OUT/bench.err; cat $OUT/bench_reference.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launch_bench.log 2>&1; echo "ncu list rc=$?"
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 40 -c 3 -f -o $OUT/prof \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_bench.log 2>&1; echo "ncu full rc=$?"
ls -la $OUT
