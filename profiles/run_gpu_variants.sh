#!/bin/bash
# developer pass: per-kernel times of the split path / Flock batch for every library variant built by
# tools/build_variants.py.  usage: bash profiles/run_gpu_variants.sh <tag> [time_split args]
TAG=${1:-var}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
for so in levelsetpy_b200/_hjb200.so levelsetpy_b200/_hjb200_*.so; do
  timeout 300 python tools/time_split.py --lib $so "$@" 2> $OUT/err_$(basename $so).txt | tail -1 | tee -a $OUT/times.jsonl
done
