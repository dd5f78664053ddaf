#!/bin/bash
TAG=${1:-r02l}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_driver.py tests/test_gpu_operators.py tests/test_eno_schemes.py tests/test_gpu_parity.py tests/test_gpu_slab.py -x -q -k "not full_horizon" > $OUT/pytest.txt 2>&1; tail -8 $OUT/pytest.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_n1.json 2> $OUT/bench_n1.err
grep "^\[bench\]" $OUT/bench_n1.err | grep -v timing | cut -c1-900; tail -2 $OUT/bench_n1.err | cut -c1-300; cut -c1-1500 $OUT/bench_n1.json
timeout 300 python bench.py --steps 10 --warmup 3 --weno intended --no-cpu --e2e-steps 0 > $OUT/bench_intended.json 2> $OUT/bench_intended.err
tail -2 $OUT/bench_intended.err | cut -c1-300; cut -c1-400 $OUT/bench_intended.json
