o=gpurun_out
python -m pytest tests -m gpu -x -q > $o/r2q_tests.log 2>&1; echo "tests rc=$?"; tail -3 $o/r2q_tests.log
python bench.py > $o/r2q_bench.json 2> $o/r2q_bench.err; echo "bench rc=$?"; tail -3 $o/r2q_bench.err
python - <<'PY'
import json
for l in open('gpurun_out/r2q_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['value']); print({k:v for k,v in d['e2e'].items() if 'note' not in k})
PY
