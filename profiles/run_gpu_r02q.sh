#!/bin/bash
TAG=${1:-r02q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for so in levelsetpy_b200/_hjb200.so levelsetpy_b200/_hjb200_*.so; do
  timeout 300 python tools/time_split.py --lib $so --what 6d,4d --planes0 20 2> $OUT/err_$(basename $so).txt | tail -1 | tee -a $OUT/times.jsonl
done
bash profiles/run_gpu_sanitize.sh $TAG
