"""Files that do not fit one call: cut into batches of whole records and driven batch by batch.

ssd_run_files / ssd_annotated_file hold a whole file (pair) on the GPU; that stops at 64 GiB per file and, before that,
at what the GPU's memory holds next to the output.  The reference has the same problem in RAM and solves it with line
bins (numReadsBin, V2:80-81, 207-262) and chunk files (LIB:6-97).  Here the host finds record-aligned cut points
(ssd_fastq_record_cuts: newline counting, no '@' heuristics), the mate file is cut at the same RECORDS, and every batch
goes through ssd_load_file_range -> ssd_demux_loaded / ssd_annotated_loaded -> filter -> ssd_emit_file (append).

  * V2.demultiplex_fun writes every matched pair (V2:367): batches are independent, one pass.
  * V2.calc_demux_results decides per cell over the WHOLE file (V2:389-412) before it copies a record (V2:415-454): two
    passes (ssd_b200.distributed.two_pass): per-cell tables over all batches first, then filter + write per batch.

SSD_BATCH_BYTES (default 16 GiB per file and batch) is the cut target; the drop-in modules switch to this path for files
above it.  The output is byte-identical to the single call's (tests/test_gpu_dropin.py forces tiny batches on the bundled
files and checks the reference's golden digests)."""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

from . import _lib as L

DEFAULT_BATCH_BYTES = 16 << 30


def batch_bytes() -> int:
    return max(1, int(os.environ.get("SSD_BATCH_BYTES", DEFAULT_BATCH_BYTES)))


def needs_batches(*paths: str) -> bool:
    limit = batch_bytes()
    return any(os.path.exists(p) and os.path.getsize(p) > limit for p in paths)


def record_cuts(path: str, target_bytes: Optional[int] = None, records: Optional[Sequence[int]] = None) -> Tuple[List[int], List[int], int]:
    """(record indices, byte offsets, complete records in the file).  By size (target_bytes) or at given records."""
    lib = L.load()
    size = os.path.getsize(path)  # FileNotFoundError like the reference's open()
    if records is None:
        cap = size // max(1, int(target_bytes)) + 3
        rin, n_in = None, 0
    else:
        cap = len(records) + 1
        rin, n_in = (C.c_uint64 * len(records))(*records), len(records)
    while True:
        rec, off = (C.c_uint64 * cap)(), (C.c_uint64 * cap)()
        n, total = C.c_uint64(), C.c_uint64()
        rc = lib.ssd_fastq_record_cuts(os.fsencode(path), int(target_bytes or 0), rin, n_in, rec, off, cap, C.byref(n), C.byref(total))
        if rc == L.ERANGE:
            cap *= 2
            continue
        if rc != L.OK:
            raise OSError("cannot read %s" % path)
        return list(rec[: n.value]), list(off[: n.value]), total.value


def pair_batches(path_r1: str, path_r2: str, target_bytes: Optional[int] = None):
    """[(offset 1, length 1, offset 2, length 2)]: the same record range of both files per batch; read 1 is cut by size, read
    2 at the same records (the last batch takes whatever is left of both files, so surplus records still meet the mate
    check of V2:349-350)."""
    target = batch_bytes() if target_bytes is None else int(target_bytes)
    rec, off1, _ = record_cuts(path_r1, target_bytes=target)
    _, off2, _ = record_cuts(path_r2, records=rec)
    off1[-1], off2[-1] = os.path.getsize(path_r1), os.path.getsize(path_r2)
    return [(off1[k], off1[k + 1] - off1[k], off2[k], off2[k + 1] - off2[k]) for k in range(len(rec) - 1)] or [(0, off1[-1], 0, off2[-1])]


def _truncate(path: str, append: bool):
    if not append:
        open(path, "wb").close()


def _emit_file(dm, which: int, out_path: str) -> int:
    n = C.c_size_t()
    dm._check(dm.lib.ssd_emit_file(dm._ctx, which, os.fsencode(out_path), 1, C.byref(n)))
    return n.value


def run_files_batched(dm, path_r1: str, path_r2: str, out_path: str, which: int = L.EMIT_PASSING, append: bool = True,
                      target_bytes: Optional[int] = None):
    """Demultiplexer.run_files for files of any size.  Returns (bytes written, stats summed over the batches)."""
    from .distributed import two_pass
    batches = pair_batches(path_r1, path_r2, target_bytes)
    p1, p2 = os.fsencode(path_r1), os.fsencode(path_r2)
    _truncate(out_path, append)
    written = [0]

    def phase1():
        for o1, n1, o2, n2 in batches:
            dm._check(dm.lib.ssd_load_file_range(dm._ctx, 0, p1, o1, n1))
            dm._check(dm.lib.ssd_load_file_range(dm._ctx, 1, p2, o2, n2))
            dm._check(dm.lib.ssd_demux_loaded(dm._ctx))
            yield None

    if which == L.EMIT_MATCHED:  # no decision across batches: one pass
        # (V2.demultiplex_fun never looks at -m: only the output sizes are needed, so the per-batch filter is built from the
        # read histogram alone and the per-cell counts, which are per batch here, are not reported)
        total = {}
        for _ in phase1():
            st = dm.build_filter()
            written[0] += _emit_file(dm, which, out_path)
            for key, v in st.items():
                if key not in ("n_cells", "n_passing_cells", "n_retained", "bytes_passing"):
                    total[key] = total.get(key, 0) + v
        return written[0], total

    def emit(_, st):
        written[0] += _emit_file(dm, which, out_path)

    total = two_pass(dm, phase1, emit)
    return written[0], total


def filter_annotated_file_batched(dm, path_in: str, out_path: str, append: bool = True, target_bytes: Optional[int] = None):
    """Demultiplexer.filter_annotated_file (V2.calc_demux_results) for a MergedCells_1.fastq of any size."""
    from .distributed import two_pass
    target = batch_bytes() if target_bytes is None else int(target_bytes)
    _, off, _ = record_cuts(path_in, target_bytes=target)
    off[-1] = os.path.getsize(path_in)
    ranges = [(off[k], off[k + 1] - off[k]) for k in range(len(off) - 1)] or [(0, off[-1])]
    p = os.fsencode(path_in)
    _truncate(out_path, append)
    written = [0]

    def phase1():
        for o, n in ranges:
            dm._check(dm.lib.ssd_load_file_range(dm._ctx, 0, p, o, n))
            dm._check(dm.lib.ssd_annotated_loaded(dm._ctx))
            yield None

    def emit(_, st):
        written[0] += _emit_file(dm, L.EMIT_PASSING, out_path)

    total = two_pass(dm, phase1, emit)
    return written[0], total


def plan_shards(path_r1: str, path_r2: str, world: int, target_bytes: Optional[int] = None):
    """What rank 0 decides for run_files_sharded: the pair of files cut into batches of whole records (at most SSD_BATCH_BYTES
    of read 1 each, and small enough that there are about `world` of them when the files allow), dealt to the ranks as
    contiguous runs in file order.  Returns one list of (offset 1, length 1, offset 2, length 2) per rank (possibly empty)."""
    from .distributed import deal_evenly
    target = target_bytes
    if target is None:
        # as many batches as a multiple of `world` allows under the per-batch limit, all about the same size
        size = max(1, os.path.getsize(path_r1))
        per_rank = -(-size // (world * batch_bytes()))  # batches every rank gets
        target = max(1, -(-size // (world * per_rank)))
    batches = pair_batches(path_r1, path_r2, target)
    return [batches[lo:hi] for lo, hi in deal_evenly(len(batches), world)]


TILE_LOG2 = 20  # the shard phase counts newlines per MiB


def device_tile_counts(dm, slot: int):
    """count(path, offset, length) -> uint32 numpy array: newlines per 2^TILE_LOG2-byte tile of that byte range, counted on the
    device: the range is loaded into the context's input buffer `slot` (pinned rings, H2D) and ssd_loaded_newlines counts it."""
    import numpy as np

    def count(path: str, offset: int, length: int):
        if length <= 0:
            return np.zeros(0, dtype=np.uint32)
        dm._check(dm.lib.ssd_load_file_range(dm._ctx, slot, os.fsencode(path), int(offset), int(length)))
        tiles = (length + (1 << TILE_LOG2) - 1) >> TILE_LOG2
        out = np.zeros(tiles, dtype=np.uint32)
        n = C.c_uint64()
        dm._check(dm.lib.ssd_loaded_newlines(dm._ctx, slot, TILE_LOG2, out.ctypes.data_as(C.c_void_p), tiles, C.byref(n)))
        return out[: n.value]
    return count


def _nth_newline_end(path: str, offset: int, length: int, tile_counts, j: int) -> int:
    """File offset just behind the j-th (1-based) newline of the byte range [offset, offset + length), given the range's
    per-tile newline counts: the tile comes from their prefix sums, the position inside it from the tile's bytes."""
    import numpy as np
    pre = np.cumsum(tile_counts, dtype=np.int64)
    t = int(np.searchsorted(pre, j, side="left"))
    before = int(pre[t - 1]) if t else 0
    lo = offset + (t << TILE_LOG2)
    with open(path, "rb") as fh:
        fh.seek(lo)
        buf = np.frombuffer(fh.read(min(1 << TILE_LOG2, offset + length - lo)), dtype=np.uint8)
    return lo + int(np.flatnonzero(buf == 10)[j - before - 1]) + 1


def plan_shards_device(path_r1: str, path_r2: str, world: int, rank: int, count1, count2, allgather, target_bytes: Optional[int] = None):
    """The batch plan of run_files_sharded WITHOUT anybody reading a whole file ("shard phase"; the reference
    aligns its chunks by counting lines, LIB:6-97): both files are cut into world x k byte ranges of equal size, every rank
    counts the newlines of ITS ranges (count1 / count2: device_tile_counts), the per-range totals of all ranks are gathered
    (allgather: list of int64 per rank -> the lists of all ranks), and line indices follow by prefix sums.  The v-th batch starts
    at the first record that begins behind read 1's v-th byte cut, record r being line 4 r of BOTH files; whoever holds the
    range containing that line in a file reports the byte offset (all ranks learn all offsets in a second gather).  No '@'
    heuristics, like ssd_fastq_record_cuts.  Returns the same structure as plan_shards: one list of (offset 1, length 1,
    offset 2, length 2) per rank."""
    import numpy as np
    s1, s2 = os.path.getsize(path_r1), os.path.getsize(path_r2)
    limit = batch_bytes() if target_bytes is None else max(1, int(target_bytes))
    per_rank = max(1, -(-max(s1, s2, 1) // (world * limit)))
    v_all = world * per_rank
    cut1 = [v * s1 // v_all for v in range(v_all + 1)]
    cut2 = [v * s2 // v_all for v in range(v_all + 1)]
    mine = range(rank * per_rank, (rank + 1) * per_rank)
    tiles1 = {v: count1(path_r1, cut1[v], cut1[v + 1] - cut1[v]) for v in mine}
    tiles2 = {v: count2(path_r2, cut2[v], cut2[v + 1] - cut2[v]) for v in mine}
    totals = allgather([int(tiles1[v].sum()) for v in mine] + [int(tiles2[v].sum()) for v in mine])
    n1 = np.array([t for part in totals for t in part[:per_rank]], dtype=np.int64)
    n2 = np.array([t for part in totals for t in part[per_rank:]], dtype=np.int64)
    before1, before2 = np.concatenate([[0], np.cumsum(n1)]), np.concatenate([[0], np.cumsum(n2)])  # newlines before each byte cut
    # batch v starts at record r_v = the first record whose first line begins behind byte cut v of read 1: line 4 r_v, which
    # begins right behind newline number 4 r_v (counted from 1) of either file; past the last such newline: the end of the file
    found = []
    for v in range(1, v_all):
        target = 4 * (int(before1[v]) // 4 + 1)
        for which, (before, cut, tiles, path) in enumerate(((before1, cut1, tiles1, path_r1), (before2, cut2, tiles2, path_r2))):
            u = int(np.searchsorted(before, target, side="left")) - 1  # the range holding newline number `target`
            if 0 <= u < v_all and u in tiles:
                found.append((v, which, _nth_newline_end(path, cut[u], cut[u + 1] - cut[u], tiles[u], target - int(before[u]))))
    off = np.full((2, v_all + 1), -1, dtype=np.int64)
    off[:, 0] = 0
    for part in allgather([x for f in found for x in f]):
        for k in range(0, len(part), 3):
            off[part[k + 1], part[k]] = part[k + 2]
    off[0, v_all], off[1, v_all] = s1, s2
    for which, size in ((0, s1), (1, s2)):  # a record past the end of a file: nothing left of that file from there on
        for v in range(1, v_all):
            if off[which, v] < 0:
                off[which, v] = size
    batches = [(int(off[0, v]), int(off[0, v + 1] - off[0, v]), int(off[1, v]), int(off[1, v + 1] - off[1, v])) for v in range(v_all)]
    return [batches[r * per_rank: (r + 1) * per_rank] for r in range(world)]


def run_files_sharded(dm, path_r1: str, path_r2: str, out_path: str, which: int = L.EMIT_PASSING, append: bool = True,
                      world: int = 1, group=None, target_bytes: Optional[int] = None):
    """One pair of files, N ranks (one process per GPU, torch.distributed initialised): the reference's fan-out over chunk
    files (CLI3:52-64) with GPUs in place of worker processes.  Rank 0 cuts the files into batches of whole records (at
    most SSD_BATCH_BYTES, at least N of them when the files allow), every rank takes a contiguous run of batches through
    the two passes of ssd_b200.distributed.two_pass (read histogram all-reduced, (cell, UMI) keys sent to the owner rank
    batch by batch, so -m is decided globally and exactly), writes its records to <out_path>.part<rank>, and rank 0 strings
    the parts together in rank order -- the bytes of the single-GPU call.  Returns (bytes written by all ranks, stats with
    the per-rank sums added up); call it on every rank."""
    import torch
    import torch.distributed as dist
    from .distributed import two_pass
    if world <= 1:
        return run_files_batched(dm, path_r1, path_r2, out_path, which, append, target_bytes)
    rank = dist.get_rank(group)
    box = [None]
    if rank == 0:
        try:
            for p in (path_r1, path_r2):
                os.path.getsize(p)  # FileNotFoundError like the reference's open()
            _truncate(out_path, append)
        except OSError as e:  # every rank raises what rank 0 met
            box[0] = e
    dist.broadcast_object_list(box, src=0, group=group)
    if isinstance(box[0], Exception):
        raise box[0]
    if os.environ.get("SSD_SHARD_PLAN", "device") == "host":
        # the first version: rank 0 reads both files through and cuts them (ssd_fastq_record_cuts)
        if rank == 0:
            box[0] = plan_shards(path_r1, path_r2, world, target_bytes)
        dist.broadcast_object_list(box, src=0, group=group)
        plan = box[0]
    else:
        def allgather(values):
            parts = [None] * world
            dist.all_gather_object(parts, [int(x) for x in values], group=group)
            return parts
        plan = plan_shards_device(path_r1, path_r2, world, rank, device_tile_counts(dm, 0), device_tile_counts(dm, 1), allgather, target_bytes)
    mine, rounds = plan[rank], max(len(p) for p in plan)
    part = "%s.part%d" % (out_path, rank)
    open(part, "wb").close()
    p1, p2 = os.fsencode(path_r1), os.fsencode(path_r2)
    written = [0]

    def phase1():
        for o1, n1, o2, n2 in mine:
            dm._check(dm.lib.ssd_load_file_range(dm._ctx, 0, p1, o1, n1))
            dm._check(dm.lib.ssd_load_file_range(dm._ctx, 1, p2, o2, n2))
            dm._check(dm.lib.ssd_demux_loaded(dm._ctx))
            yield None

    def emit(_, st):
        written[0] += _emit_file(dm, which, part)

    total = two_pass(dm, phase1, emit, world=world, group=group, rounds=rounds)
    keys = sorted(k for k in total if k not in ("n_cells", "n_passing_cells"))
    if not total:  # a rank without batches: the cell counts come from the tables every rank holds
        keys = []
    names = [None]
    if rank == 0:
        names[0] = keys
    dist.broadcast_object_list(names, src=0, group=group)
    sums = torch.tensor([float(total.get(k, 0)) for k in names[0]] + [float(written[0])], dtype=torch.float64, device="cuda")
    dist.all_reduce(sums, group=group)
    cells = torch.tensor([float(total.get("n_cells", 0)), float(total.get("n_passing_cells", 0))], dtype=torch.float64, device="cuda")
    dist.all_reduce(cells, op=dist.ReduceOp.MAX, group=group)
    stats = {k: int(v) for k, v in zip(names[0], sums.tolist())}
    stats["n_cells"], stats["n_passing_cells"] = int(cells[0]), int(cells[1])
    dist.barrier(group=group)
    if rank == 0:
        with open(out_path, "ab") as out:
            for r in range(world):
                with open("%s.part%d" % (out_path, r), "rb") as fh:
                    while True:
                        buf = fh.read(64 << 20)
                        if not buf:
                            break
                        out.write(buf)
                os.remove("%s.part%d" % (out_path, r))
    dist.barrier(group=group)
    return int(sums[-1]), stats


def run_split_sharded(dm, path_r1: str, path_r2: str, out_dir: str, round_names=None, world: int = 1, group=None):
    """Split mode (one file per passing cell) on N ranks sharing one pair of files: every rank takes its byte
    range of both files (plan_shards_device: cut at the same records), decides -m with the other ranks (global_filter), buckets
    ITS records by cell on the device, and the ranks append their buckets to the per-cell files one after the other, in rank
    order -- the bytes the single-GPU call writes, since the ranks hold consecutive record ranges (flushBuffers appends the
    same way, splitseq_utilities.py:15-21).  One batch per rank (the pair must fit world x SSD_BATCH_BYTES).  Returns (number
    of files this rank wrote to, the rank's stats with the global cell counts)."""
    import torch.distributed as dist
    from .distributed import global_filter
    if world <= 1:
        plan = [[(0, os.path.getsize(path_r1), 0, os.path.getsize(path_r2))]]
        rank = 0
    else:
        rank = dist.get_rank(group)

        def allgather(values):
            parts = [None] * world
            dist.all_gather_object(parts, [int(x) for x in values], group=group)
            return parts
        plan = plan_shards_device(path_r1, path_r2, world, rank, device_tile_counts(dm, 0), device_tile_counts(dm, 1), allgather)
    if any(len(p) != 1 for p in plan):
        raise ValueError("split mode takes one batch per rank: the files are larger than world x SSD_BATCH_BYTES")
    o1, n1, o2, n2 = plan[rank][0]
    dm._check(dm.lib.ssd_load_file_range(dm._ctx, 0, os.fsencode(path_r1), o1, n1))
    dm._check(dm.lib.ssd_load_file_range(dm._ctx, 1, os.fsencode(path_r2), o2, n2))
    dm._check(dm.lib.ssd_demux_loaded(dm._ctx))
    stats = global_filter(dm, world, group=group)
    n_files = 0
    for turn in range(world):
        if turn == rank:
            n_files = len(dm.write_split_files(out_dir, round_names, stats))
        if world > 1:
            dist.barrier(group=group)
    return n_files, stats
