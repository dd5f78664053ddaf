#!/bin/bash
# 1-GPU pass after the ghost-warp kernels: parity tests of the paths they touch, then the default bench line (with
# its workloads blocks; writes the 41^6 per-plane checksums that the N>1 verify compares with).
TAG=${1:-r02d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_split.py tests/test_gpu_batch.py tests/test_gpu_slab.py tests/test_gpu_parity.py -x -q > $OUT/pytest.txt 2>&1
tail -15 $OUT/pytest.txt
timeout 600 python bench.py --steps 20 --warmup 5 --write-golden $OUT/bench_dubins6d_41_checksums.json > $OUT/bench_n1.json 2> $OUT/bench_n1.err
tail -5 $OUT/bench_n1.err; cat $OUT/bench_n1.json
