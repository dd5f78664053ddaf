#!/bin/bash
# 4-GPU pass: the multi-process slab parity worker at 4 ranks (3-4 planes of dim 0 per rank in the small cases) and the
# default bench line at N=4
TAG=${1:-n4}
OUT=gpurun_out/$TAG
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/slab_rank_worker.py > $OUT/worker.log 2>&1
echo "worker rc=$?" >> $OUT/worker.log
grep '^{' $OUT/worker.log > $OUT/worker_cases_n4.jsonl
tail -2 $OUT/worker.log | cut -c1-200
python - <<PY
import json
rows=[json.loads(l) for l in open("$OUT/worker_cases_n4.jsonl")]
print(len(rows), "cases;", sum(r["ok"] for r in rows), "ok;", sum(r["bit_identical"] for r in rows), "bit-identical")
for r in rows:
    if not r["ok"]: print(r)
PY
timeout 600 $TR --master-port 29512 bench.py --gpus 4 --steps 20 --warmup 5 > $OUT/bench_n4.json 2> $OUT/bench_n4.err
grep "^\[bench\]" $OUT/bench_n4.err | grep -v timing | cut -c1-900; tail -2 $OUT/bench_n4.err | cut -c1-200; cut -c1-300 $OUT/bench_n4.json
