#!/bin/bash
# Final tree, 2 GPUs: the tests the 1-GPU pass skips (multi-process slab parity over real NCCL / CUDA IPC)
TAG=${1:-r03d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 500 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q -rs > $OUT/pytest_multirank.txt 2>&1; echo "rc=$?" >> $OUT/pytest_multirank.txt
tail -6 $OUT/pytest_multirank.txt | cut -c1-300
