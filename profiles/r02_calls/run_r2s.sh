o=gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $o/r2s_multi.log 2>&1; echo "multi tests rc=$?"; tail -5 $o/r2s_multi.log
SSD_SHARD_PLAN=host python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "cli" > $o/r2s_multi_host.log 2>&1; echo "host-plan rc=$?"; tail -2 $o/r2s_multi_host.log
