#!/bin/bash
# developer pass: pass-2 bound finding at a representative dim-0 extent (20 planes: 5 tiles, 2 of them at the boundary),
# ncu of the pass-2 kernel, and the new parity tests (512^3 sub-slab vs oracle)
TAG=${1:-r02h}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for so in levelsetpy_b200/_hjb200.so levelsetpy_b200/_hjb200_*.so; do
  timeout 300 python tools/time_split.py --lib $so --what 6d --planes0 20 2> $OUT/err_$(basename $so).txt | tail -1 | tee -a $OUT/times.jsonl
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stage_vec -s 5 -c 1 -f -o $OUT/prof_p2 \
    python tools/time_split.py --what 6d --planes0 20 --reps 1 > $OUT/prof.log 2>&1
python profiles/ncu_summary.py $OUT/prof_p2.ncu-rep > $OUT/ncu_p2.txt 2>&1
ncu -i $OUT/prof_p2.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $OUT/ncu_p2_source.csv.gz
rm -f $OUT/prof_p2.ncu-rep
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -k "subslab or config0" > $OUT/pytest.txt 2>&1; tail -5 $OUT/pytest.txt
