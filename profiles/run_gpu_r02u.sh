#!/bin/bash
# final 1-GPU pass: the default bench line, its ncu launch list, and ncu --set full of the three 512^3 stage kernels
TAG=${1:-r02u}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/smi.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_n1.json 2> $OUT/bench_n1.err
tail -2 $OUT/bench_n1.err | cut -c1-200; cut -c1-600 $OUT/bench_n1.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cut -c1-300 $OUT/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stage_tma -s 9 -c 3 -f -o $OUT/prof_stage \
   python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 0 --blocks none > $OUT/prof_stage.log 2>&1
python profiles/ncu_summary.py $OUT/prof_stage.ncu-rep > $OUT/ncu_stage_summary.txt 2>&1
rm -f $OUT/prof_stage.ncu-rep
grep -E "^==|gpu__time_duration|dram__bytes|dram_throughput|l1tex__throughput|pipe_fp64_cycles|issue_active" $OUT/ncu_stage_summary.txt | cut -c1-140
timeout 300 python bench.py --steps 10 --warmup 3 --weno intended --no-cpu --e2e-steps 0 --blocks none > $OUT/bench_intended.json 2> $OUT/bench_intended.err; cut -c1-330 $OUT/bench_intended.json
