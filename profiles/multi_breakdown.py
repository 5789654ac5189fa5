"""Where a multi-GPU step spends its time: run under torchrun, prints per-phase milliseconds on rank 0.
CUDA events on the step's stream between the phases of ssd_b200.distributed.global_filter (no synchronisation inside the step,
so the phases overlap with the host exactly as in bench.py); the sum of the phases is the step."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "split-seq_demultiplexing_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import fastq_util as fu  # noqa: E402
import ssd_b200  # noqa: E402
from ssd_b200 import distributed as D  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
wl = fu.load_whitelist()
n = 10_000_000
dm = ssd_b200.Demultiplexer(wl, errors=1, min_count=10, device=local, record_bytes_hint=160)
spec = ssd_b200.synth_spec(wl, n, seed=0x5EED0001, first_index=rank * n + 1)
d1, d2 = ssd_b200.synth_device(dm, spec)
stream = torch.cuda.Stream()
out = torch.empty(int(n * 330), dtype=torch.uint8, device="cuda")
cap = D.UNDECIDED_CAPACITY
names = ["demux", "distinct_local", "pack + all-reduce(sum, 7 MB)", "undecided_keys", "all-gather", "count_distinct + adopt", "build_filter (sync)", "emit"]
acc = [0.0] * len(names)
und = 0
STEPS, WARM = 13, 3
with torch.cuda.stream(stream):
    dm.set_stream(stream.cuda_stream)
    for it in range(STEPS):
        torch.cuda.synchronize()
        dist.barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
        k = 0
        ev[k].record(); k += 1
        dm.demux_device(d1.data_ptr(), d1.numel(), d2.data_ptr(), d2.numel())
        ev[k].record(); k += 1
        dm.distinct_local()
        ev[k].record(); k += 1
        packed = dm.pack_decision()
        dist.all_reduce(packed)
        ev[k].record(); k += 1
        mine = dm.undecided_keys(packed, cap)
        ev[k].record(); k += 1
        everything = torch.empty((world * (cap + 1), 2), dtype=torch.int64, device=mine.device)
        dist.all_gather_into_tensor(everything, mine)
        ev[k].record(); k += 1
        dm.count_distinct(everything[:0, 0], everything)
        dm.adopt_decision(packed)
        ev[k].record(); k += 1
        st = dm.build_filter()
        ev[k].record(); k += 1
        dm.emit_device(ssd_b200.EMIT_PASSING, out.data_ptr(), out.numel())
        ev[k].record(); k += 1
        torch.cuda.synchronize()
        if it >= WARM:
            for j in range(len(names)):
                acc[j] += ev[j].elapsed_time(ev[j + 1])
            und = int(everything.view(world, cap + 1, 2)[:, 0, 1].max())
res = torch.tensor(acc, device="cuda") / (STEPS - WARM)
dist.all_reduce(res, op=dist.ReduceOp.MAX)
if rank == 0:
    print({"n_gpus": world, "phases_ms_max_over_ranks": {a: round(float(b), 3) for a, b in zip(names, res.tolist())}, "sum_ms": round(float(res.sum()), 3),
           "undecided_keys_max_per_rank": und, "capacity": cap})
dist.destroy_process_group()
