#!/bin/bash
# multi-GPU pass: parity tests of the split / slab paths, the multi-process slab worker (tests/slab_rank_worker.py) and
# the bench line with its workloads blocks.  usage (under gpurun --gpus N): bash profiles/run_gpu_r02c.sh <tag> N [p2p]
TAG=${1:-r02c}
NP=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node"
if [ "$NP" -le 2 ]; then
  timeout 600 python -m pytest tests/test_gpu_split.py tests/test_gpu_slab.py -x -q > $OUT/pytest.txt 2>&1; tail -3 $OUT/pytest.txt
fi
WNP=$NP; if [ "$WNP" -gt 4 ]; then WNP=0; fi
[ "$WNP" -gt 0 ] && timeout 600 $TR $WNP --master-addr 127.0.0.1 --master-port 29511 tests/slab_rank_worker.py > $OUT/worker.log 2>&1
echo "worker rc=$?" >> $OUT/worker.log
grep '^{' $OUT/worker.log > $OUT/worker_cases_n$WNP.jsonl
[ "$WNP" -gt 0 ] && tail -3 $OUT/worker.log; [ "$WNP" -gt 0 ] && python - <<PY
import json
rows=[json.loads(l) for l in open("$OUT/worker_cases_n$WNP.jsonl")]
print(len(rows), "cases;", sum(r["ok"] for r in rows), "ok;", sum(r["bit_identical"] for r in rows), "bit-identical")
for r in rows:
    if not r["ok"]: print(r)
PY
timeout 900 $TR $NP --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NP --steps 20 --warmup 5 > $OUT/bench_n${NP}_peer.json 2> $OUT/bench_n${NP}_peer.err
tail -3 $OUT/bench_n${NP}_peer.err | cut -c1-300; cat $OUT/bench_n${NP}_peer.json
if [ "$3" = "p2p" ]; then
  timeout 900 $TR $NP --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $NP --steps 20 --warmup 5 --transport p2p --blocks dubins6d > $OUT/bench_n${NP}_p2p.json 2> $OUT/bench_n${NP}_p2p.err
  tail -3 $OUT/bench_n${NP}_p2p.err | cut -c1-300; cat $OUT/bench_n${NP}_p2p.json
fi
