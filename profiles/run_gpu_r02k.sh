#!/bin/bash
TAG=${1:-r02k}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_driver.py tests/test_gpu_operators.py tests/test_restrict_rk2.py -x -q > $OUT/pytest.txt 2>&1; tail -15 $OUT/pytest.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --block-n 21 > $OUT/bench_n2_dev.json 2> $OUT/bench_n2_dev.err
grep "^\[bench\]" $OUT/bench_n2_dev.err | cut -c1-1200; tail -2 $OUT/bench_n2_dev.err | cut -c1-300
