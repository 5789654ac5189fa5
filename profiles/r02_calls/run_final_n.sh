o=gpurun_out
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 profiles/multi_breakdown.py > $o/final_breakdown_n$N.log 2>&1; tail -1 $o/final_breakdown_n$N.log
if [ "$N" = 8 ]; then python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 profiles/pcie_duplex.py > $o/final_pcie_n$N.log 2>&1; grep gpus_copying $o/final_pcie_n$N.log; fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 3 > $o/final_bench_n$N.json 2> $o/final_bench_n$N.err; echo "bench rc=$?"; tail -2 $o/final_bench_n$N.err
python - $N <<'PY'
import json,sys
for l in open('gpurun_out/final_bench_n%s.json'%sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e'].get('serial_value')); print(d['roofline']['stage_ms_per_step']); print(d.get('parity_check',{}).get('ok'))
        for k,v in d.get('configs',{}).items(): print(k, v.get('value'), v.get('ms_per_step'), (v.get('e2e') or {}).get('value'), v.get('error'))
PY
