"""Where a multi-GPU step spends its time: run under torchrun, prints per-phase milliseconds on rank 0."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "split-seq_demultiplexing_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import fastq_util as fu  # noqa: E402
import ssd_b200  # noqa: E402
from ssd_b200 import distributed as D  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
wl = fu.load_whitelist()
n = 10_000_000
dm = ssd_b200.Demultiplexer(wl, errors=1, min_count=10, device=local, record_bytes_hint=160)
spec = ssd_b200.synth_spec(wl, n, seed=0x5EED0001, first_index=rank * n + 1)
d1, d2 = ssd_b200.synth_device(dm, spec)
stream = torch.cuda.Stream()
out = torch.empty(int(n * 330), dtype=torch.uint8, device="cuda")
acc = {}


def tick(name, t0):
    torch.cuda.synchronize()
    acc[name] = acc.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
    return time.perf_counter()


with torch.cuda.stream(stream):
    dm.set_stream(stream.cuda_stream)
    for it in range(8):
        if it == 3:
            acc.clear()
        torch.cuda.synchronize()
        dist.barrier()
        t = time.perf_counter()
        dm.demux_device(d1.data_ptr(), d1.numel(), d2.data_ptr(), d2.numel())
        t = tick("demux", t)
        dist.all_reduce(dm.reads_histogram())
        t = tick("allreduce hist", t)
        keys, offsets, irr = dm.partition_keys(world)
        t = tick("partition", t)
        kc = torch.tensor([offsets[k + 1] - offsets[k] for k in range(world)], dtype=torch.int64, device=keys.device)
        rk, ri = D.exchange_owned(keys, kc, irr)
        t = tick("exchange", t)
        dm.count_distinct_owned(rk, ri, rank, world)
        t = tick("count distinct", t)
        dist.all_reduce(dm.distinct_histogram())
        t = tick("allreduce distinct", t)
        st = dm.build_filter()
        t = tick("filter", t)
        dm.emit_device(ssd_b200.EMIT_PASSING, out.data_ptr(), out.numel())
        t = tick("emit", t)
if rank == 0:
    print({k: round(v / 5, 3) for k, v in acc.items()}, "sum", round(sum(acc.values()) / 5, 3))
dist.destroy_process_group()
