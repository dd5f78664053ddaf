#!/bin/bash
TAG=${1:-r02r}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_split.py tests/test_gpu_slab.py -x -q > $OUT/pytest.txt 2>&1; tail -3 $OUT/pytest.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/slab_rank_worker.py > $OUT/worker.log 2>&1
echo "worker rc=$?" >> $OUT/worker.log
grep '^{' $OUT/worker.log > $OUT/worker_cases_n2.jsonl
tail -2 $OUT/worker.log | cut -c1-200
python - <<PY
import json
rows=[json.loads(l) for l in open("$OUT/worker_cases_n2.jsonl")]
print(len(rows), "cases;", sum(r["ok"] for r in rows), "ok;", sum(r["bit_identical"] for r in rows), "bit-identical")
PY
