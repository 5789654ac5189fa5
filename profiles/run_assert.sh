#!/bin/bash
# The parity, drop-in and fuzz suites against a library built with -DSSD_ASSERT (device-side invariant checks of the scans):
#   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -DSSD_ASSERT \
#        -o profiles/v_assert.so split-seq_demultiplexing_b200/csrc/ssd_b200.cu        (in the authoring container)
#   gpurun -- 'bash profiles/run_assert.sh'
o=gpurun_out
SSD_B200_LIB=/root/repo/profiles/v_assert.so python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -m gpu -x -q > $o/assert_tests.log 2>&1; echo "assert build: tests rc=$?"; tail -3 $o/assert_tests.log; grep -c SSD_ASSERT $o/assert_tests.log
