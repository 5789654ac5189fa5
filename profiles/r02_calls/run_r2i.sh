o=gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $o/r2i_multi.log 2>&1; echo "multi tests rc=$?"; tail -3 $o/r2i_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3 --no-other-configs > $o/r2i_bench2.json 2> $o/r2i_bench2.err; echo "bench rc=$?"; tail -3 $o/r2i_bench2.err
python - <<'PY'
import json
for l in open('gpurun_out/r2i_bench2.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e'].get('serial_value')); print(d['roofline']['stage_ms_per_step']); print(d.get('parity_check'))
PY
