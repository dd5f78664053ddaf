#!/bin/bash
# ncu pass of the bench command: launch list (gpu__time_duration per kernel) + one --set full capture of the three stage
# kernels of the 512^3 step.  usage (from the repo root, under gpurun): bash profiles/run_gpu_profile.sh <tag>
TAG=${1:-r01d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 60 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > $OUT/launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stage_tma -s 9 -c 3 -f -o $OUT/prof_stage \
    python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 0 > $OUT/prof_bench.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_stage -c 12 --csv --log-file $OUT/launches_6d.csv \
    python bench.py --workload dubins6d --steps 1 --warmup 3 --no-cpu --e2e-steps 0 > $OUT/launches_6d.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_stage -c 12 --csv --log-file $OUT/launches_4d.csv \
    python bench.py --workload dint4d --steps 1 --warmup 3 --no-cpu --e2e-steps 0 > $OUT/launches_4d.log 2>&1
for w in dint4d dubins6d flockbatch; do timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --e2e-steps 0 > $OUT/bench_$w.json 2> $OUT/bench_$w.err; done
ls -la $OUT
