o=gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -m gpu -x -q -k "not full_size" > $o/r2t_tests.log 2>&1; echo "tests rc=$?"; tail -15 $o/r2t_tests.log
