o=gpurun_out
python bench.py --config 4 --steps 2 --warmup 3 --no-other-configs --no-cpu-baseline > $o/r2y_c4.json 2> $o/r2y_c4.err; echo "bench rc=$?"; tail -2 $o/r2y_c4.err
python - <<'PY'
import json
for l in open('gpurun_out/r2y_c4.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['value'], d['result'])
PY
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "batches" 2>&1 | tail -2
