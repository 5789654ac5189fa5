o=gpurun_out
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 profiles/multi_breakdown.py > $o/r2l_breakdown_n$N.log 2>&1; tail -2 $o/r2l_breakdown_n$N.log
