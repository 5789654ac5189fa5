o=gpurun_out
g++ -O2 -pthread -o /tmp/host_io_probe profiles/host_io_probe.cpp && /tmp/host_io_probe /dev/shm/probe.bin 3000000000 > $o/r2o_ioprobe.log 2>&1; rm -f /dev/shm/probe.bin; cat $o/r2o_ioprobe.log
SSD_IO_TRACE=1 python profiles/files_probe.py > $o/r2o_files.log 2>&1; cat $o/r2o_files.log | tail -24
python -m pytest tests/test_gpu_dropin.py -m gpu -x -q > $o/r2o_tests.log 2>&1; echo "tests rc=$?"; tail -3 $o/r2o_tests.log
