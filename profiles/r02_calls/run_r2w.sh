o=gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -m gpu -x -q -k "split or per_round or full_size" > $o/r2w_tests.log 2>&1; echo "tests rc=$?"; tail -3 $o/r2w_tests.log
python bench.py --config 5 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e-files --no-other-configs > $o/r2w_bench5.json 2> $o/r2w_bench5.err; echo "bench rc=$?"; tail -2 $o/r2w_bench5.err
python - <<'PY'
import json
for l in open('gpurun_out/r2w_bench5.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['stage_ms_per_step'])
PY
