#!/bin/bash
# odeCFL2 over the generic hooks (hj_stage 4 + hj_deriv_range 4), plus the existing RK2 / restrict / driver tests that touch hj_stage
TAG=${1:-r03e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 400 python -m pytest tests/test_generic_dynsys.py tests/test_restrict_rk2.py tests/test_gpu_driver.py tests/test_gpu_parity.py -q -m gpu -x > $OUT/pytest.txt 2>&1; echo "rc=$?" >> $OUT/pytest.txt
tail -8 $OUT/pytest.txt | cut -c1-250
