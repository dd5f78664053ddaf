#!/bin/bash
# the default bench line at N ranks of one box (what the driver's scaling run launches)
TAG=${1:-scale}
NP=${2:-8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NP --steps 20 --warmup 5 > $OUT/bench_n${NP}.json 2> $OUT/bench_n${NP}.err
grep "^\[bench\]" $OUT/bench_n${NP}.err | grep -v timing | cut -c1-900; tail -2 $OUT/bench_n${NP}.err | cut -c1-200; cut -c1-300 $OUT/bench_n${NP}.json
