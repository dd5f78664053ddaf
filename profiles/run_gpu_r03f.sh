#!/bin/bash
# rest of the -m gpu suite on the final library (the suites run_gpu_r03e.sh did not cover)
TAG=${1:-r03f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 225 python -m pytest tests/test_gpu_operators.py tests/test_eno_schemes.py tests/test_gpu_batch.py tests/test_gpu_slab.py tests/test_gpu_split.py tests/test_gpu_fullsize.py -q -m gpu -x > $OUT/pytest.txt 2>&1; echo "rc=$?" >> $OUT/pytest.txt
tail -6 $OUT/pytest.txt | cut -c1-250
