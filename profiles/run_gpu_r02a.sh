#!/bin/bash
# Round-2 baseline pass (code as of round-1 end): today's numbers for configs 2-5 on one GPU and ncu --set full
# captures of the two kernels of the dimension-split path (6-D pair, 8 x 41^5 sub-grid) and of the intended-WENO stage
# kernels (256^3).  usage (repo root, under gpurun): bash profiles/run_gpu_r02a.sh <tag>
TAG=${1:-r02a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/smi.txt 2>&1
for w in dubins6d dint4d flockbatch; do
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --e2e-steps 0 > $OUT/bench_$w.json 2> $OUT/bench_$w.err
done
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > $OUT/bench_air3d.json 2> $OUT/bench_air3d.err
timeout 300 python bench.py --weno intended --steps 5 --warmup 3 --no-cpu --e2e-steps 0 > $OUT/bench_intended.json 2> $OUT/bench_intended.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stage -s 18 -c 6 -f -o $OUT/prof_split6d \
    python bench.py --workload dubins6d --planes0 8 --steps 1 --warmup 3 --no-cpu --e2e-steps 0 > $OUT/prof_split6d.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stage_tma -s 9 -c 3 -f -o $OUT/prof_intended \
    python bench.py --weno intended --n 256 --steps 1 --warmup 3 --no-cpu --e2e-steps 0 > $OUT/prof_intended.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_stage -s 18 -c 6 -f -o $OUT/prof_split4d \
    python bench.py --workload dint4d --steps 1 --warmup 3 --no-cpu --e2e-steps 0 > $OUT/prof_split4d.log 2>&1
for r in split6d intended split4d; do
  python profiles/ncu_summary.py $OUT/prof_$r.ncu-rep > $OUT/ncu_$r.txt 2>&1
  ncu -i $OUT/prof_$r.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $OUT/ncu_${r}_source.csv.gz
  rm -f $OUT/prof_$r.ncu-rep
done
ls -la $OUT
cat $OUT/bench_*.json | cut -c1-400
