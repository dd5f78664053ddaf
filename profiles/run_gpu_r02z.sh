#!/bin/bash
# 2 x 2-nodes-per-thread plane-ring kernel (hj_quad_kernel.cuh) in the tuning harness: bits vs production, ms per stage;
# then the pipelined host-buffer step after the ghost-plane fix
TAG=${1:-r02z}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 120 tools/tune_tma 512 10 6550 quad > $OUT/tune_quad.txt 2>&1; echo "rc=$?" >> $OUT/tune_quad.txt
timeout 120 tools/tune_tma 512 10 6550 quad_R8_7_b2 103 > $OUT/tune_quad_cz103.txt 2>&1; echo "rc=$?" >> $OUT/tune_quad_cz103.txt
timeout 120 tools/tune_tma 512 10 6550 quad_R8_7_b2 64 > $OUT/tune_quad_cz64.txt 2>&1; echo "rc=$?" >> $OUT/tune_quad_cz64.txt
cat $OUT/tune_quad.txt $OUT/tune_quad_cz103.txt $OUT/tune_quad_cz64.txt | cut -c1-230
timeout 240 python tools/e2e_chunk_sweep.py > $OUT/e2e_chunk_sweep.jsonl 2> $OUT/e2e_chunk_sweep.err
tail -3 $OUT/e2e_chunk_sweep.err | cut -c1-300; cut -c1-200 $OUT/e2e_chunk_sweep.jsonl
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slab.py -x -q -m gpu -k "pipelined or ode_cfl3 or stage_range" > $OUT/pytest_pipelined.txt 2>&1; tail -3 $OUT/pytest_pipelined.txt
