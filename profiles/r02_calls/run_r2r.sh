o=gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -m gpu -x -q -k "not full_size" > $o/r2r_tests.log 2>&1; echo "tests rc=$?"; tail -15 $o/r2r_tests.log
python bench.py --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 0 --no-other-configs --no-e2e-files --no-parity-check 2>&1 | python -c "
import sys, json
l = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(l['ms_per_step'], 3), {k: round(v, 3) for k, v in l['roofline']['stage_ms_per_step'].items()})"
