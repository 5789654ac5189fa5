o=gpurun_out
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config 4 --steps 2 --warmup 3 --no-other-configs --no-parity-check > $o/c4_n$N.json 2> $o/c4_n$N.err; echo "bench rc=$?"; tail -2 $o/c4_n$N.err
python - $N <<'PY'
import json,sys
for l in open('gpurun_out/c4_n%s.json'%sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['value'], d['result'])
PY
