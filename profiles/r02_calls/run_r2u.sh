o=gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $o/r2u_multi.log 2>&1; echo "multi tests rc=$?"; tail -15 $o/r2u_multi.log
python -m pytest tests/test_gpu_dropin.py -m gpu -x -q -k "split" > $o/r2u_split.log 2>&1; echo "split tests rc=$?"; tail -3 $o/r2u_split.log
