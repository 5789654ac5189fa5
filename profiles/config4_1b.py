#!/usr/bin/env python
"""BASELINE.json configs[3]: 10^9 synthetic read pairs sharded over the GPUs of one box, -m 10 decided globally
(all-reduce of the read histogram, exact distinct-UMI exchange by owner rank), every rank writing its own records.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        profiles/config4_1b.py [--pairs 1000000000] [--batch 100000000]

A rank's shard (pairs / N) is generated on the device batch by batch (S1B seed 0x5EED0003) and goes through
ssd_b200.distributed.demux_batches: two passes over the batches.  Prints one JSON line (rank 0): wall time of the two
passes, the time spent generating input inside them (measured with CUDA events, reported and subtracted), pairs/s, and
the checks: sum of the global histogram == matched reads, retained == reads of the passing cells, output bytes ==
planned bytes, 4 newlines per retained record."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "split-seq_demultiplexing_b200")):
    sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=1_000_000_000)
    ap.add_argument("--batch", type=int, default=100_000_000)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import ssd_b200
    from ssd_b200.distributed import demux_batches
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl = [l.strip() for l in open(os.path.join(ROOT, "tests", "golden", "barcodes_8mer.txt")) if l.strip()]
    dm = ssd_b200.Demultiplexer(wl, errors=1, min_count=10, device=local, record_bytes_hint=160)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    dm.set_stream(stream.cuda_stream)
    per_rank = args.pairs // world
    first_rank = 1 + rank * per_rank
    gen_ms = [0.0]

    def batches():
        done = 0
        while done < per_rank:
            n = min(args.batch, per_rank - done)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            d = ssd_b200.synth_device(dm, ssd_b200.synth_spec(wl, n, seed=0x5EED0003, first_index=first_rank + done, cell_pool=1_000_000))
            e1.record()
            e1.synchronize()
            gen_ms[0] += e0.elapsed_time(e1)
            yield d
            del d
            done += n

    checks = {"bytes": 0, "newlines": 0, "planned": 0}

    def sink(out, st):
        checks["bytes"] += out.numel()
        checks["planned"] += st["bytes_passing"]
        for lo in range(0, out.numel(), 1 << 31):
            checks["newlines"] += int((out[lo:lo + (1 << 31)] == 10).sum())

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # warm-up on a small shard: NCCL connections, module loading (the big buffers still grow inside the timed run)
    def small():
        yield ssd_b200.synth_device(dm, ssd_b200.synth_spec(wl, 1_000_000, seed=1, first_index=1 + rank * 1_000_000))
    demux_batches(dm, small, world=world)
    sync()
    gen_ms[0] = 0.0
    timing = {}
    t0 = time.perf_counter()
    total = demux_batches(dm, batches, world=world, sink=sink, timing=timing)
    sync()
    dt = time.perf_counter() - t0
    hist = dm.reads_histogram().to(torch.int64)       # the global tables are still in place after the last batch
    distinct = dm.distinct_histogram().to(torch.int64)
    passing = distinct >= 10
    t = torch.tensor([dt, gen_ms[0] / 1e3, total["n_retained"], total["n_matched"], checks["bytes"], checks["newlines"], checks["planned"],
                      total["n_records_r1"]], dtype=torch.float64, device="cuda")
    tmax = t.clone()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    if rank == 0:
        retained, matched, nbytes, newlines, planned, records = (int(x) for x in t[2:].tolist())
        wall, gen = float(tmax[0]), float(tmax[1])
        line = {"workload": "config 4: %d synthetic read pairs (150 bp), e=1, -m 10 distinct UMIs, %d GPU(s), batches of <= %d pairs per rank, "
                            "two passes" % (args.pairs, world, args.batch),
                "n_gpus": world, "pairs": records, "wall_s": wall, "generator_s_inside": gen, "rank0_phases_s": timing,
                "pairs_per_s_two_passes": records / max(1e-9, wall - gen),
                "n_matched": matched, "n_retained": retained, "n_cells": total["n_cells"], "n_passing_cells": total["n_passing_cells"],
                "output_bytes": nbytes,
                "checks": {"hist_sum_equals_matched": int(hist.sum()) == matched,
                           "retained_equals_reads_of_passing_cells": int(hist[passing].sum()) == retained,
                           "cells": int((hist > 0).sum()) == total["n_cells"] and int(passing.sum()) == total["n_passing_cells"],
                           "bytes_equal_planned": nbytes == planned, "four_lines_per_record": newlines == 4 * retained,
                           "all_records_seen": records == per_rank * world}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
