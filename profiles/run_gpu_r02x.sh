#!/bin/bash
# e2e chunk-height sweep of the pipelined host-buffer step + its bit-identity test
TAG=${1:-r02x}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 240 python tools/e2e_chunk_sweep.py > $OUT/e2e_chunk_sweep.jsonl 2> $OUT/e2e_chunk_sweep.err
tail -3 $OUT/e2e_chunk_sweep.err | cut -c1-300; cut -c1-200 $OUT/e2e_chunk_sweep.jsonl
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pipelined or ode_cfl3" > $OUT/pytest_pipelined.txt 2>&1; tail -3 $OUT/pytest_pipelined.txt
