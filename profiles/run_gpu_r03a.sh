#!/bin/bash
# genericHam / genericPartial over the device dynSys: parity tests vs the reference-generated golden,
# memcheck of the small generic cases, and the cost of the two-pass stage at 512^3 beside DubinsVehicleRel
TAG=${1:-r03a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python -m pytest tests/test_generic_dynsys.py tests/test_gpu_operators.py -q -m gpu > $OUT/pytest_generic.txt 2>&1; tail -25 $OUT/pytest_generic.txt | cut -c1-250
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_cases.py --generic-only > $OUT/sanitizer_memcheck_generic.log 2>&1
echo "memcheck exit code $?" >> $OUT/sanitizer_memcheck_generic.log; tail -5 $OUT/sanitizer_memcheck_generic.log
timeout 200 python tools/time_generic.py 512 10 > $OUT/time_generic.jsonl 2> $OUT/time_generic.err; tail -3 $OUT/time_generic.err; cat $OUT/time_generic.jsonl
