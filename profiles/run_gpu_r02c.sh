#!/bin/bash
# 2-GPU pass: multi-process slab parity (tests/slab_rank_worker.py) and the bench line with its workloads blocks
# over both halo transports.  usage (under gpurun --gpus 2): bash profiles/run_gpu_r02c.sh <tag> [nproc]
TAG=${1:-r02c}
NP=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/slab_rank_worker.py > $OUT/worker.log 2>&1
echo "worker rc=$?" >> $OUT/worker.log
grep '^{' $OUT/worker.log > $OUT/worker_cases.jsonl
timeout 800 $TR --master-port 29512 bench.py --gpus $NP --steps 10 --warmup 3 > $OUT/bench_n${NP}_peer.json 2> $OUT/bench_n${NP}_peer.err
timeout 800 $TR --master-port 29513 bench.py --gpus $NP --steps 10 --warmup 3 --transport p2p > $OUT/bench_n${NP}_p2p.json 2> $OUT/bench_n${NP}_p2p.err
tail -3 $OUT/worker.log; cat $OUT/worker_cases.jsonl | cut -c1-250
tail -5 $OUT/bench_n${NP}_peer.err; cat $OUT/bench_n${NP}_peer.json
tail -5 $OUT/bench_n${NP}_p2p.err; cat $OUT/bench_n${NP}_p2p.json
