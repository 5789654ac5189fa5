o=gpurun_out
nproc > $o/r2h_host.txt; free -g >> $o/r2h_host.txt; lscpu | head -20 >> $o/r2h_host.txt
g++ -O2 -pthread -o /tmp/host_io_probe profiles/host_io_probe.cpp && /tmp/host_io_probe /dev/shm/probe.bin 3000000000 > $o/r2h_ioprobe.log 2>&1; rm -f /dev/shm/probe.bin
python -m pytest tests -m gpu -x -q > $o/r2h_tests.log 2>&1; echo "tests rc=$?"; tail -3 $o/r2h_tests.log
python bench.py > $o/r2h_bench.json 2> $o/r2h_bench.err; echo "bench rc=$?"; tail -3 $o/r2h_bench.err
