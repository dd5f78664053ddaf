#!/bin/bash
# developer pass: slab tests of the pieces protocol, variant timings, ncu of the two 6-D kernels (default library)
TAG=${1:-r02f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_slab.py -x -q > $OUT/pytest.txt 2>&1; tail -5 $OUT/pytest.txt
for so in levelsetpy_b200/_hjb200.so levelsetpy_b200/_hjb200_*.so; do
  timeout 300 python tools/time_split.py --lib $so --what 6d,fb 2> $OUT/err_$(basename $so).txt | tail -1 | tee -a $OUT/times.jsonl
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stage -s 8 -c 2 -f -o $OUT/prof_split6d \
    python tools/time_split.py --what 6d --reps 1 > $OUT/prof.log 2>&1
python profiles/ncu_summary.py $OUT/prof_split6d.ncu-rep > $OUT/ncu_split6d.txt 2>&1
ncu -i $OUT/prof_split6d.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $OUT/ncu_split6d_source.csv.gz
rm -f $OUT/prof_split6d.ncu-rep
grep -E "^==|gpu__time_duration|dram__bytes|l1tex__throughput|bank_conflicts|pipe_fp64_cycles|issue_active|warps_active|stalled_(long|short|wait|math|not_sel|barrier|branch)" $OUT/ncu_split6d.txt | cut -c1-150
