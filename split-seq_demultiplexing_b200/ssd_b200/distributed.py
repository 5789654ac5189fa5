"""Multi-GPU orchestration: one process per GPU (torch.distributed), reads sharded by record range.

The per-cell READ counts are additive, so they are all-reduced (NCCL over NVLink).  The reference's -m, however,
counts DISTINCT UMIs per cell (V2:411), which is not additive across shards: every rank groups its (cell, UMI) keys
by the rank that OWNS their cell (one partition pass on the device, duplicates kept), the groups are exchanged
all-to-all, every rank sorts what it received and counts the distinct keys of the cells it owns, and those counts
(zero elsewhere) are all-reduced so that all ranks hold the same, exact, global table.  (exchange_keys, the first
version, exchanges locally sorted + deduplicated keys by contiguous cell range; it is kept for hosts that want the
unique lists.)

The functions below only use torch ops + collectives on whatever device the tensors live on, so the exchange logic
is exercised on CPU with the gloo backend in tests/test_distributed_gloo.py.
"""
from __future__ import annotations

import os
from typing import List, Tuple

IRR_CELL_SHIFT = 36  # irregular key: hi = cell << 36 | len << 32 | bytes[0..4)  (include/ssd_b200.h)


def cell_range_edges(n_cells_total: int, world: int, device):
    import torch
    per = (n_cells_total + world - 1) // world
    return torch.arange(0, world + 1, device=device, dtype=torch.int64) * per


def split_points(keys, irr, n_cells_total: int, world: int, umi_bits: int):
    """Where the sorted key arrays have to be cut so that piece k holds the cells rank k owns.
    keys: int64[n] sorted, cell = key >> umi_bits.  irr: int64[m, 2] sorted by (hi, lo), cell = hi >> 36."""
    import torch
    edges = cell_range_edges(n_cells_total, world, keys.device)
    kb = torch.searchsorted(keys, edges << umi_bits) if keys.numel() else torch.zeros(world + 1, dtype=torch.int64, device=keys.device)
    ib = (torch.searchsorted(irr[:, 0].contiguous(), edges << IRR_CELL_SHIFT) if irr.numel()
          else torch.zeros(world + 1, dtype=torch.int64, device=keys.device))
    return kb, ib


def exchange_keys(keys, irr, n_cells_total: int, world: int, umi_bits: int, group=None):
    """All-to-all of the locally unique keys by cell range.  Returns (received regular keys, received irregular
    keys): everything any rank holds for the cells this rank owns (duplicates across source ranks included)."""
    import torch
    import torch.distributed as dist
    kb, ib = split_points(keys, irr, n_cells_total, world, umi_bits)
    send = torch.stack([kb[1:] - kb[:-1], ib[1:] - ib[:-1]], dim=1).contiguous()  # [world, 2]
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    send_l, recv_l = send.cpu().tolist(), recv.cpu().tolist()
    rk = torch.empty(sum(r[0] for r in recv_l), dtype=torch.int64, device=keys.device)
    dist.all_to_all_single(rk, keys.contiguous(), [r[0] for r in recv_l], [s[0] for s in send_l], group=group)
    ri = torch.empty((sum(r[1] for r in recv_l), 2), dtype=torch.int64, device=keys.device)
    dist.all_to_all_single(ri, irr.contiguous(), [r[1] for r in recv_l], [s[1] for s in send_l], group=group)
    return rk, ri


def owner_of_cells(cells, shift: int, n_buckets: int, world: int):
    """Owner rank of each cell id (int64 tensor): ((cell >> shift) * world) // n_buckets, as the library partitions."""
    return ((cells >> shift) * world) // n_buckets


def group_by_owner(rows, cells, shift: int, n_buckets: int, world: int):
    """Reorder `rows` (first dimension) so that the rows of owner 0 come first, then owner 1, ...; returns the
    reordered rows and the per-owner counts (int64[world])."""
    import torch
    owner = owner_of_cells(cells, shift, n_buckets, world)
    order = torch.argsort(owner, stable=True)
    counts = torch.bincount(owner, minlength=world)[:world] if owner.numel() else torch.zeros(world, dtype=torch.int64, device=rows.device)
    return rows[order].contiguous(), counts


def exchange_raw(keys, key_counts, irr, irr_counts, group=None):
    """All-to-all of keys already grouped by owner rank.  keys: int64[n], key_counts: int64[world]; irr: int64[m, 2],
    irr_counts: int64[world].  Returns what the peers sent this rank (regular, irregular)."""
    import torch
    import torch.distributed as dist
    send = torch.stack([key_counts, irr_counts], dim=1).contiguous()  # [world, 2]
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    send_l, recv_l = send.cpu().tolist(), recv.cpu().tolist()
    rk = torch.empty(sum(r[0] for r in recv_l), dtype=torch.int64, device=keys.device)
    dist.all_to_all_single(rk, keys.contiguous(), [r[0] for r in recv_l], [s[0] for s in send_l], group=group)
    ri = torch.empty((sum(r[1] for r in recv_l), 2), dtype=torch.int64, device=keys.device)
    dist.all_to_all_single(ri, irr.contiguous(), [r[1] for r in recv_l], [s[1] for s in send_l], group=group)
    return rk, ri


def exchange_owned(keys, key_counts, irr, group=None):
    """The exchange of the GPU path: regular keys (grouped by owner rank, key_counts[k] of them for rank k; a list or a tensor) go
    all-to-all; the few irregular keys go to everybody (all-gather, padded with -1 rows to the longest list), the
    receiver keeps those of its own cells.  Returns (regular keys received, all irregular keys)."""
    import torch
    import torch.distributed as dist
    send_keys = [int(x) for x in (key_counts.tolist() if hasattr(key_counts, "tolist") else key_counts)]  # host-side counts
    world = len(send_keys)
    send = torch.tensor([[k, irr.shape[0]] for k in send_keys], dtype=torch.int64, device=keys.device)  # [world, 2]
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    recv_l = recv.cpu().tolist()
    rk = torch.empty(sum(r[0] for r in recv_l), dtype=torch.int64, device=keys.device)
    dist.all_to_all_single(rk, keys.contiguous(), [r[0] for r in recv_l], send_keys, group=group)
    longest = max(r[1] for r in recv_l)
    if longest == 0:
        return rk, torch.empty((0, 2), dtype=torch.int64, device=keys.device)
    pad = torch.full((longest, 2), -1, dtype=torch.int64, device=keys.device)
    pad[: irr.shape[0]] = irr
    everything = torch.empty((world * longest, 2), dtype=torch.int64, device=keys.device)
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(everything, pad, group=group)
    else:
        dist.all_gather(list(everything.view(world, longest, 2).unbind(0)), pad, group=group)
    return rk, everything


def _adopt_torch_stream(dm):
    """Point the library context at torch's current CUDA stream (no-op for the CPU stand-ins of the gloo tests), so that the
    library's kernels and torch's collectives are ordered against each other.  Torch reports the legacy default stream as
    handle 0, which ssd_set_stream reads as "the context's own stream"; CUDA's explicit handle for it is cudaStreamLegacy (1)."""
    import torch
    if not hasattr(dm, "set_stream") or not torch.cuda.is_available():
        return
    handle = torch.cuda.current_stream().cuda_stream
    dm.set_stream(handle if handle != 0 else 1)


UNDECIDED_CAPACITY = int(os.environ.get("SSD_UNDECIDED_CAPACITY", 1 << 18))  # first guess: keys per rank and step in global_filter's exchange (4 MiB)


def global_filter(dm, world: int, group=None, exact_table: bool = False, capacity: int = None) -> dict:
    """After dm.demux_device(...) on every rank: make the -m decision global and exact, return dm's stats.

    reads mode: one all-reduce of the read histogram.  distinct-UMI mode (V2:411), without shipping every key: every rank counts
    the distinct UMIs of ITS OWN records per cell; one all-reduce(SUM) of a 64-bit word per cell carries both the reads of the
    cell and how many ranks count >= m by themselves.  A cell passes as soon as one rank does, and fails when all ranks together
    hold < m reads of it (distinct <= reads).  Only the cells in between are undecided, and every rank holds fewer than m
    distinct UMIs of each of them: their keys travel as one fixed-size packed list per rank (all-gather; a few hundred thousand
    keys instead of every rank's millions), everybody counts them exactly, done.  Two collectives, no owner partition and NO
    host round trip: the sizes never come back to the host before the step's own synchronisation (build_filter).  The list
    size follows what the previous steps needed (the headers of the gathered lists, identical on all ranks); should a rank
    hold more undecided keys than the list takes, the step is repeated with the exchange by owner rank, which has no such
    limit.  Afterwards the distinct table holds the exact global count for the undecided cells, m for the cells some rank
    decided alone and 0 elsewhere: a lower bound that decides -m like the exact table, identical on all ranks.
    exact_table=True runs the exchange by owner rank in the first place and leaves the exact global distinct count of every
    cell (what a per-cell table wants)."""
    if world <= 1:
        return dm.filter_single()
    import torch
    import torch.distributed as dist
    # torch's collectives are ordered on torch's current stream only: the library must enqueue on that very stream, or it
    # reads what NCCL has not delivered yet (and NCCL sends tables the kernels have not finished).  Idempotent.
    _adopt_torch_stream(dm)
    if dm.filter_mode != "distinct_umi":
        dist.all_reduce(dm.reads_histogram(), group=group)
        return dm.build_filter()
    if exact_table:
        return _global_filter_by_owner(dm, world, group)
    cap = int(capacity or getattr(dm, "_undecided_capacity", None) or UNDECIDED_CAPACITY)
    dm.distinct_local()
    packed = dm.pack_decision()
    dist.all_reduce(packed, group=group)
    mine = dm.undecided_keys(packed, cap)
    everything = torch.empty((world * (cap + 1), 2), dtype=torch.int64, device=mine.device)
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(everything, mine, group=group)
    else:
        dist.all_gather(list(everything.view(world, cap + 1, 2).unbind(0)), mine.clone(), group=group)
    most = everything.view(world, cap + 1, 2)[:, 0, 1].max()  # the largest list any rank wanted to send
    dm.count_distinct(everything[:0, 0], everything)  # exact global counts of the undecided cells (zero elsewhere)
    dm.adopt_decision(packed)
    st = dm.build_filter()  # synchronises: reading `most` now costs no extra wait
    most = int(most.item())
    if capacity is None:  # next time: room for half as much again, in steps of 64 Ki keys (the same number on every rank)
        dm._undecided_capacity = max(1 << 16, -(-(most + most // 2) // 65536) * 65536)
    if most > cap:
        return _global_filter_by_owner(dm, world, group, reads_are_global=True)
    return st


def _global_filter_by_owner(dm, world: int, group=None, reads_are_global: bool = False) -> dict:
    """The -m decision with the exact global distinct count of EVERY cell: all (cell, UMI) keys go to the rank that owns their cell
    (all-to-all, duplicates kept; irregular keys to everybody), every rank counts the cells it owns, the table is all-reduced."""
    import torch.distributed as dist
    rank = dist.get_rank(group)
    pending = None if reads_are_global else dist.all_reduce(dm.reads_histogram(), group=group, async_op=True)  # overlaps the key exchange
    keys, offsets, irr = dm.partition_keys(world)
    rk, ri = exchange_owned(keys, [offsets[k + 1] - offsets[k] for k in range(world)], irr, group=group)
    dm.count_distinct_owned(rk, ri, rank, world)
    dist.all_reduce(dm.distinct_histogram(), group=group)
    if pending is not None:
        pending.wait()
    return dm.build_filter()


def two_pass(dm, phase1, emit, world: int = 1, group=None, timing: dict = None, rounds: int = None) -> dict:
    """The two passes over batches that a global -m decision needs (see demux_batches).  `phase1` is a callable returning an
    iterator; every step of the iterator has run phase 1 of the next batch on dm (ssd_demux_device, ssd_demux_loaded,
    ssd_annotated_loaded ...) and yields whatever `emit` needs to know about the batch.  It is gone through twice.
    emit(batch, stats) is called in pass 2 once the filter of the batch is built from the GLOBAL tables; it writes the
    batch's output (ssd_emit_device, ssd_emit_file ...).  Returns the statistics summed over the batches (cell counts are
    global).  With N > 1 every batch of pass 1 is one round of collectives, so all ranks must go through the same number
    of rounds: `rounds` = the largest batch count of any rank (ranks with fewer batches, or none, send empty messages)."""
    import time
    import torch
    import torch.distributed as dist
    distinct_mode = dm.filter_mode == "distinct_umi"
    rank = dist.get_rank(group) if world > 1 else 0
    dev = dm.reads_histogram().device  # "cuda" with the library; the gloo tests drive this function with a CPU stand-in
    sync = torch.cuda.synchronize if dev.type == "cuda" else (lambda: None)
    trim = torch.cuda.empty_cache if dev.type == "cuda" else (lambda: None)
    hist, kept_k, kept_i = None, [], []
    if world > 1 and dev.type == "cuda":
        _adopt_torch_stream(dm)
    t_start = time.perf_counter()
    for batch in phase1():
        h = dm.reads_histogram()
        hist = h.clone() if hist is None else hist.add_(h)
        if distinct_mode:
            if world > 1:
                keys, offsets, irr = dm.partition_keys(world)
                k, i = exchange_owned(keys, [offsets[j + 1] - offsets[j] for j in range(world)], irr, group=group)
            else:
                # one GPU: the batch's keys as they are (one compaction pass over the annotation words, duplicates kept);
                # the per-cell sets at the end do not need them sorted or unique
                k, _, i = dm.partition_keys(1)
            kept_k.append(k.clone())
            kept_i.append(i.clone())
        sync()
        dm._keep = []  # the batch may go before the next one is produced (two of them need not fit)
        del batch
    if distinct_mode and world > 1:
        if rounds is None:  # every rank goes through as many exchange rounds as the rank with the most batches
            r = torch.tensor([len(kept_k)], dtype=torch.int64, device=dev)
            dist.all_reduce(r, op=dist.ReduceOp.MAX, group=group)
            rounds = int(r.item())
        for _ in range(len(kept_k), rounds):  # this rank has fewer batches than another: empty rounds
            k, i = exchange_owned(torch.empty(0, dtype=torch.int64, device=dev), [0] * world,
                                  torch.empty((0, 2), dtype=torch.int64, device=dev), group=group)
            kept_k.append(k.clone())
            kept_i.append(i.clone())
    if hist is None:
        if world == 1:
            return {}
        hist = torch.zeros_like(dm.reads_histogram())
    t_pass1 = time.perf_counter()
    distinct = None
    if distinct_mode:
        all_k = torch.cat(kept_k) if kept_k else torch.empty(0, dtype=torch.int64, device=dev)
        all_i = torch.cat(kept_i) if kept_i else torch.empty((0, 2), dtype=torch.int64, device=dev)
        del kept_k, kept_i
        trim()  # the library allocates with cudaMalloc: blocks torch has cached are out of its reach
        # the library may run on a stream of its own (no set_stream): hand tensors over only when both sides are done
        sync()
        if world > 1:
            dm.count_distinct_owned(all_k, all_i, rank, world)
            sync()
            dist.all_reduce(dm.distinct_histogram(), group=group)
        else:
            dm.count_distinct(all_k, all_i)
        sync()
        distinct = dm.distinct_histogram().clone()
        del all_k, all_i
    if world > 1:
        dist.all_reduce(hist, group=group)
    sync()
    t_count = time.perf_counter()
    total = {}
    for batch in phase1():
        dm.reads_histogram().copy_(hist)
        if distinct is not None:
            dm.distinct_histogram().copy_(distinct)
        sync()
        st = dm.build_filter()
        emit(batch, st)
        for key, v in st.items():
            total[key] = v if key in ("n_cells", "n_passing_cells") else total.get(key, 0) + v
        sync()
        dm._keep = []
        del batch
        trim()  # batches of tens of GB: do not let the caching allocator fragment what is left
    if timing is not None:
        timing.update(pass1_s=t_pass1 - t_start, count_s=t_count - t_pass1, pass2_s=time.perf_counter() - t_count)
    return total


def demux_batches(dm, batches, world: int = 1, group=None, which: int = None, sink=None, timing: dict = None) -> dict:
    """Inputs too large for one call (BASELINE config 4: 10^9 read pairs; 64 GiB per file is the per-call limit, the HBM of one
    GPU the practical one): the -m decision needs every batch's per-cell read counts and (cell, UMI) keys before the first
    record can be written, so the batches are gone through twice, the way the reference goes through its chunk files
    (CLI3:56-64) and then through MergedCells_1.fastq (V2:389-454).

      pass 1  every batch: scan + match (dm.demux_device); its read histogram is added up, its (cell, UMI) keys are kept
              (one GPU: the batch's sorted unique list; N GPUs: grouped by owner rank and sent there right away).
              Then one exact distinct count over everything kept, and the all-reduces of the two tables when N > 1.
      pass 2  every batch again: scan + match, the GLOBAL tables put in place of the batch's own, filter, write.

    `batches` is a callable returning an iterator of (r1, r2) uint8 CUDA tensors (called twice; every rank passes its own
    shard's batches).  sink(out_tensor, stats) receives each batch's output in order (the tensor is only valid during the
    call); without a sink the outputs are dropped.  Returns the statistics summed over the batches (cell counts are global).
    `timing`, when given, receives the wall-clock seconds of pass 1, of the global count and of pass 2 (device synchronised)."""
    import torch
    from . import _lib as L
    which = L.EMIT_PASSING if which is None else which

    def phase1():
        for r1, r2 in batches():
            dm.demux_device(r1.data_ptr(), r1.numel(), r2.data_ptr(), r2.numel(), keep=(r1, r2))
            dev = r1.device
            del r1, r2
            yield dev

    held = {"out": None}  # one output buffer for all batches, grown when a batch needs more: tens of GB are not given back and
                          # asked for again batch after batch

    def emit(dev, st):
        need = st["bytes_matched"] if which == L.EMIT_MATCHED else st["bytes_passing"]
        if held["out"] is None or held["out"].numel() < need + 16:
            held["out"] = None
            held["out"] = torch.empty(need + 16 + need // 64, dtype=torch.uint8, device=dev)
        out = held["out"]
        n = dm.emit_device(which, out.data_ptr(), out.numel())
        if sink is not None:
            sink(out[:n], st)

    return two_pass(dm, phase1, emit, world=world, group=group, timing=timing)


def shard_record_ranges(n_records: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous record ranges per rank (the reference cuts line-aligned chunks the same way, LIB:6-97)."""
    per = (n_records + world - 1) // world
    return [(min(k * per, n_records), min((k + 1) * per, n_records)) for k in range(world)]


def deal_evenly(n_items: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous runs of n_items over `world` ranks whose lengths differ by at most one (the first n % world ranks get the
    extra item): no rank idles while another holds two more batches than it."""
    base, extra = divmod(n_items, world)
    out, lo = [], 0
    for k in range(world):
        hi = lo + base + (1 if k < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out
