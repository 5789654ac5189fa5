o=gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cell_umi" > $o/r2n_tests.log 2>&1; echo "tests rc=$?"; tail -5 $o/r2n_tests.log
SSD_IO_TRACE=1 python profiles/files_probe.py > $o/r2n_files.log 2>&1; cat $o/r2n_files.log | tail -60
