o=gpurun_out
bash profiles/run_assert.sh
python bench.py --config 3 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e-files --no-other-configs --e2e-steps 0 > $o/r2x_bench3.json 2> $o/r2x_bench3.err; echo "bench rc=$?"; tail -2 $o/r2x_bench3.err
python - <<'PY'
import json
for l in open('gpurun_out/r2x_bench3.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['value'], d['roofline']['stage_ms_per_step'])
PY
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config3 or fuzz or long_reads or rule_cases or adversarial" 2>&1 | tail -2
