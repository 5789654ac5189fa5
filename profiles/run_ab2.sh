# A/B of prebuilt libraries profiles/v_*.so: per-stage milliseconds (device-resident leg only); "T" after a name = also run the parity tests with it
o=gpurun_out
for v in "$@"; do
  name=${v%T}
  SSD_B200_LIB=/root/repo/profiles/v_$name.so python bench.py --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 0 --no-other-configs --no-e2e-files --no-parity-check 2>&1 | python -c "
import sys, json
l = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$name', round(l['ms_per_step'], 3), {k: round(v, 3) for k, v in l['roofline']['stage_ms_per_step'].items()})"
  if [ "$v" != "$name" ]; then SSD_B200_LIB=/root/repo/profiles/v_$name.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not full_size" 2>&1 | tail -1; fi
done
