#!/bin/bash
# One GPU-box pass: parity tests, bench, ncu launch list, ncu full capture of the stage kernels.
# usage (from the repo root, under gpurun): bash profiles/run_gpu_round.sh <tag>
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; tail -2 $OUT/bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 60 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > $OUT/launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stage_tma -s 9 -c 3 -f -o $OUT/prof_stage \
    python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 0 > $OUT/prof_bench.log 2>&1
ls -la $OUT
