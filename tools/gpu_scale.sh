#!/bin/bash
# Weak-scaling run on N GPUs of one box (trees shard by rank, no data-path collective): bash tools/gpu_scale.sh <tag> <N> [trees per gpu]
set -u
TAG=$1; N=$2; T=${3:-64}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for TT in 1 $T; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --trees-per-gpu $TT > $OUT/scale_n${N}_t${TT}.json 2>> $OUT/scale.err
  python -c "
import json; d=[json.loads(l) for l in open('$OUT/scale_n${N}_t${TT}.json') if l.startswith('{')][-1]; print('N=$N trees/gpu=$TT: %.0f tokens/s, %.0f trees/s, %.1f us/layer-call, e2e %.2f ms' % (d['value'], d['trees_per_s'], d['us_per_layer_call'], d['e2e']['ms_per_step']))"
done
