#!/bin/bash
TAG=${1:-r02m}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for p0 in 6 5; do
for so in levelsetpy_b200/_hjb200.so levelsetpy_b200/_hjb200_*.so; do
  timeout 300 python tools/time_split.py --lib $so --what 6d --planes0 $p0 2> $OUT/err_$(basename $so).txt | tail -1 | tee -a $OUT/times.jsonl
done; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_intended.csv \
   python bench.py --weno intended --n 256 --steps 2 --warmup 3 --no-cpu --e2e-steps 0 --blocks none > $OUT/ncu_int.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("$OUT/launches_intended.csv")) if len(r)>10]
hdr=rows[0]; ik=hdr.index("Kernel Name"); iv=hdr.index("Metric Value")
agg={}
for r in rows[1:]:
    k=r[ik][:60]; agg.setdefault(k,[]).append(float(r[iv].replace(",","")))
for k,v in agg.items(): print("%-62s n=%3d avg %.1f us"%(k,len(v),sum(v)/len(v)/1e3))
PY
timeout 600 python -m pytest tests/test_eno_schemes.py tests/test_gpu_driver.py -x -q -k "eno or stop_conditions" > $OUT/pytest.txt 2>&1; tail -3 $OUT/pytest.txt
