#!/bin/bash
# hj_deriv_range through the plane-ring stage kernel (dt = 0, dead output buffer): parity of the generic path again,
# memcheck, cost of the two-pass stage at 512^3
TAG=${1:-r03b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python -m pytest tests/test_generic_dynsys.py -q -m gpu > $OUT/pytest_generic.txt 2>&1; tail -25 $OUT/pytest_generic.txt | cut -c1-250
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_cases.py --generic-only > $OUT/sanitizer_memcheck_generic.log 2>&1
echo "memcheck exit code $?" >> $OUT/sanitizer_memcheck_generic.log; tail -5 $OUT/sanitizer_memcheck_generic.log
timeout 200 python tools/time_generic.py 512 10 > $OUT/time_generic.jsonl 2> $OUT/time_generic.err; tail -3 $OUT/time_generic.err; cat $OUT/time_generic.jsonl
