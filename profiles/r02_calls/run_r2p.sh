o=gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 profiles/multi_breakdown.py > $o/r2p_breakdown_n$N.log 2>&1; tail -1 $o/r2p_breakdown_n$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 3 --no-other-configs > $o/r2p_bench$N.json 2> $o/r2p_bench$N.err; echo "bench rc=$?"; tail -3 $o/r2p_bench$N.err
python - $N <<'PY'
import json,sys
for l in open('gpurun_out/r2p_bench%s.json'%sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e'].get('serial_value')); print(d['roofline']['stage_ms_per_step']); print(d.get('parity_check',{}).get('ok'))
PY
