#!/bin/bash
# 8-GPU pass: the default bench line (halo planes stored from inside pass 2, both sides) and the 41^6 block again with
# the hybrid transport (one side fused, the other through the copy engines)
TAG=${1:-r02s}
NP=${2:-8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29512 bench.py --gpus $NP --steps 20 --warmup 5 > $OUT/bench_n${NP}.json 2> $OUT/bench_n${NP}.err
grep "^\[bench\]" $OUT/bench_n${NP}.err | grep -v timing | cut -c1-700; tail -2 $OUT/bench_n${NP}.err | cut -c1-200
timeout 600 $TR --master-port 29513 bench.py --gpus $NP --steps 20 --warmup 5 --fused hybrid --blocks dubins6d --e2e-steps 0 > $OUT/bench_n${NP}_hybrid.json 2> $OUT/bench_n${NP}_hybrid.err
grep "^\[bench\]" $OUT/bench_n${NP}_hybrid.err | grep -v timing | cut -c1-700; tail -2 $OUT/bench_n${NP}_hybrid.err | cut -c1-200
