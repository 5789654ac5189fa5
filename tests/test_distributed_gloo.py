"""world_size-2 (and 3) gloo tests, on CPU, of the multi-GPU exchange logic in ssd_b200/distributed.py: the keys
every rank receives are exactly the keys of the cells it owns, and per-cell distinct counts computed from them and
summed over ranks equal the single-process answer."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import conftest  # noqa: F401  (sys.path)

N_CELLS = 96 ** 3
UMI_BITS = 20


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_keys(seed):
    rng = np.random.default_rng(seed)
    cells = rng.choice(N_CELLS, size=300, replace=False)
    c = rng.choice(cells, size=20000)
    u = rng.integers(0, 1 << 12, size=20000)          # few UMIs: plenty of duplicates within and across ranks
    keys = np.unique((c.astype(np.int64) << UMI_BITS) | u)
    ci = rng.choice(cells, size=800)
    hi = (ci.astype(np.int64) << 36) | (rng.integers(0, 11, size=800).astype(np.int64) << 32) | rng.integers(0, 5, size=800)
    lo = rng.integers(0, 3, size=800).astype(np.int64)
    irr = np.unique(np.stack([hi, lo], axis=1), axis=0)
    return keys, irr


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ssd_b200 import distributed as D
    keys, irr = _make_keys(100 + rank)
    rk, ri = D.exchange_keys(torch.from_numpy(keys), torch.from_numpy(irr), N_CELLS, world, UMI_BITS)
    per = (N_CELLS + world - 1) // world
    lo_cell, hi_cell = rank * per, (rank + 1) * per
    assert bool(((rk >> UMI_BITS) >= lo_cell).all()) and bool(((rk >> UMI_BITS) < hi_cell).all())
    assert ri.numel() == 0 or (bool(((ri[:, 0] >> 36) >= lo_cell).all()) and bool(((ri[:, 0] >> 36) < hi_cell).all()))
    # per-cell distinct counts of the owned cells, then the same all-reduce the GPU path does
    distinct = torch.zeros(N_CELLS, dtype=torch.int32)
    uk = torch.unique(rk)
    distinct.index_add_(0, (uk >> UMI_BITS), torch.ones(uk.numel(), dtype=torch.int32))
    if ri.numel():
        ui = torch.unique(ri, dim=0)
        distinct.index_add_(0, (ui[:, 0] >> 36), torch.ones(ui.shape[0], dtype=torch.int32))
    dist.all_reduce(distinct)
    np.save(os.path.join(out_dir, "distinct_%d.npy" % rank), distinct.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_exchange_by_cell_range(tmp_path, world):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    tables = [np.load(os.path.join(str(tmp_path), "distinct_%d.npy" % r)) for r in range(world)]
    for t in tables[1:]:
        assert (t == tables[0]).all()                     # identical on every rank
    all_keys = np.unique(np.concatenate([_make_keys(100 + r)[0] for r in range(world)]))
    all_irr = np.unique(np.concatenate([_make_keys(100 + r)[1] for r in range(world)]), axis=0)
    want = np.zeros(N_CELLS, dtype=np.int64)
    np.add.at(want, all_keys >> UMI_BITS, 1)
    np.add.at(want, all_irr[:, 0] >> 36, 1)
    assert (tables[0] == want).all()


def _worker_raw(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ssd_b200 import distributed as D
    shift, n_buckets = 12, ((N_CELLS - 1) >> 12) + 1   # what ssd_owner_layout returns for 96^3 cells
    rng = np.random.default_rng(500 + rank)
    keys, irr = _make_keys(100 + rank)
    keys = np.concatenate([keys, rng.choice(keys, size=5000)])   # raw keys: duplicates, no order
    rng.shuffle(keys)
    irr = np.concatenate([irr, irr[rng.integers(0, len(irr), size=200)]])
    k, kc = D.group_by_owner(torch.from_numpy(keys), torch.from_numpy(keys) >> UMI_BITS, shift, n_buckets, world)
    # the exchange of the GPU path: irregular keys go to everybody, the receiver keeps its own cells
    rk, everything = D.exchange_owned(k, kc, torch.from_numpy(irr))
    everything = everything[everything[:, 0] != -1]                              # padding rows
    ri = everything[D.owner_of_cells(everything[:, 0] >> 36, shift, n_buckets, world) == rank]
    assert bool((D.owner_of_cells(rk >> UMI_BITS, shift, n_buckets, world) == rank).all())
    distinct = torch.zeros(N_CELLS, dtype=torch.int32)
    uk = torch.unique(rk)
    distinct.index_add_(0, (uk >> UMI_BITS), torch.ones(uk.numel(), dtype=torch.int32))
    if ri.numel():
        ui = torch.unique(ri, dim=0)
        distinct.index_add_(0, (ui[:, 0] >> 36), torch.ones(ui.shape[0], dtype=torch.int32))
    dist.all_reduce(distinct)
    np.save(os.path.join(out_dir, "distinct_%d.npy" % rank), distinct.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_exchange_raw_keys_by_owner(tmp_path, world):
    """The exchange the GPU path uses: raw keys (duplicates kept) grouped by owner rank, owners interleaved by bucket."""
    port = _free_port()
    mp.spawn(_worker_raw, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    tables = [np.load(os.path.join(str(tmp_path), "distinct_%d.npy" % r)) for r in range(world)]
    for t in tables[1:]:
        assert (t == tables[0]).all()
    all_keys = np.unique(np.concatenate([_make_keys(100 + r)[0] for r in range(world)]))
    all_irr = np.unique(np.concatenate([_make_keys(100 + r)[1] for r in range(world)]), axis=0)
    want = np.zeros(N_CELLS, dtype=np.int64)
    np.add.at(want, all_keys >> UMI_BITS, 1)
    np.add.at(want, all_irr[:, 0] >> 36, 1)
    assert (tables[0] == want).all()


def test_shard_record_ranges():
    from ssd_b200.distributed import shard_record_ranges
    assert shard_record_ranges(10, 4) == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert shard_record_ranges(3, 8)[3:] == [(3, 3)] * 5
    r = shard_record_ranges(1_000_000_007, 8)
    assert r[0][0] == 0 and r[-1][1] == 1_000_000_007 and all(a[1] == b[0] for a, b in zip(r, r[1:]))


# ---- two_pass (batches over time and over ranks) with a CPU stand-in for the library -------------------------------
class _StandIn:
    """What two_pass needs from a Demultiplexer, on CPU tensors: per-batch read histogram and (cell, UMI) keys, ownership
    like the library's (interleaved bucket ranges), exact distinct counts of arbitrary key arrays."""
    SHIFT, NB = 12, (N_CELLS + (1 << 12) - 1) >> 12

    def __init__(self, m):
        self.filter_mode, self.m, self._keep = "distinct_umi", m, []
        self.hist = torch.zeros(N_CELLS, dtype=torch.int32)
        self.distinct = torch.zeros(N_CELLS, dtype=torch.int32)
        self.batch_cells = torch.empty(0, dtype=torch.int64)
        self.keys = torch.empty(0, dtype=torch.int64)

    def load(self, cells, umis):  # "phase 1" of a batch
        self.batch_cells = cells
        self.keys = (cells << UMI_BITS) | umis
        self.hist.zero_()
        self.hist.index_add_(0, cells, torch.ones(cells.numel(), dtype=torch.int32))

    def reads_histogram(self):
        return self.hist

    def distinct_histogram(self):
        return self.distinct

    def local_unique_keys(self):
        return torch.unique(self.keys), torch.empty((0, 2), dtype=torch.int64)

    def distinct_local(self):
        self._count(self.keys)

    def pack_decision(self):
        self._packed = self.hist.to(torch.int64) | ((self.distinct >= self.m).to(torch.int64) << 40)
        return self._packed

    def undecided_keys(self, packed, capacity):
        # the library's packed list: row 0 = (-1, number of keys), then (cell << 36 | UMI, 0) rows, unused rows all ones
        cells = self.keys >> UMI_BITS
        und = ((packed >> 40) == 0) & ((packed & ((1 << 40) - 1)) >= self.m)
        k = self.keys[und[cells]]
        rows = torch.full((capacity + 1, 2), -1, dtype=torch.int64)
        rows[0, 1] = k.numel()
        n = min(int(k.numel()), capacity)
        rows[1: 1 + n, 0] = ((k[:n] >> UMI_BITS) << 36) | (k[:n] & ((1 << UMI_BITS) - 1))
        rows[1: 1 + n, 1] = 0
        self.n_undecided_sent = int(k.numel())
        return rows

    def adopt_decision(self, packed):
        self.hist.copy_((packed & ((1 << 40) - 1)).to(torch.int32))
        self.distinct.copy_(torch.where((packed >> 40) != 0, torch.clamp(self.distinct, min=self.m), self.distinct))

    def partition_keys(self, world):
        from ssd_b200 import distributed as D
        rows, counts = D.group_by_owner(self.keys, self.keys >> UMI_BITS, self.SHIFT, self.NB, world)
        offs = [0]
        for c in counts.tolist():
            offs.append(offs[-1] + c)
        return rows, offs, torch.empty((0, 2), dtype=torch.int64)

    def _count(self, keys):
        self.distinct.zero_()
        uk = torch.unique(keys)
        self.distinct.index_add_(0, uk >> UMI_BITS, torch.ones(uk.numel(), dtype=torch.int32))

    def count_distinct(self, keys, irr):
        if irr is not None and irr.numel():  # packed records (padding and headers have hi == -1)
            hi = irr[irr[:, 0] != -1][:, 0]
            keys = torch.cat([keys, ((hi >> 36) << UMI_BITS) | (hi & ((1 << UMI_BITS) - 1))])
        self._count(keys)

    def count_distinct_owned(self, keys, irr, rank, world):
        from ssd_b200 import distributed as D
        assert bool((D.owner_of_cells(keys >> UMI_BITS, self.SHIFT, self.NB, world) == rank).all())  # only this rank's cells arrived
        self._count(keys)

    def build_filter(self):
        passing = self.distinct >= self.m
        return {"n_cells": int((self.hist > 0).sum()), "n_passing_cells": int(passing.sum()),
                "n_retained": int(passing[self.batch_cells].sum()), "n_matched": int(self.batch_cells.numel())}


def _batches_of(rank, counts):
    rng = np.random.default_rng(7)
    cells = rng.choice(N_CELLS, size=400, replace=False)
    out = []
    for r, n_batches in enumerate(counts):
        mine = []
        for b in range(n_batches):
            n = int(rng.integers(1, 4000))
            mine.append((torch.from_numpy(rng.choice(cells, size=n)).to(torch.int64), torch.from_numpy(rng.integers(0, 40, size=n)).to(torch.int64)))
        out.append(mine)
    return out


def _two_pass_worker(rank, world, port, out_dir, counts):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ssd_b200 import distributed as D
    dm = _StandIn(m=5)
    mine = _batches_of(rank, counts)[rank]

    def phase1():
        for cells, umis in mine:
            dm.load(cells, umis)
            yield cells

    kept = []
    total = D.two_pass(dm, phase1, lambda cells, st: kept.append(st["n_retained"]), world=world, rounds=max(counts))
    np.save(os.path.join(out_dir, "tp_%d.npy" % rank), np.array([total.get("n_cells", -1), total.get("n_passing_cells", -1), sum(kept), len(kept)]))
    np.save(os.path.join(out_dir, "tp_distinct_%d.npy" % rank), dm.distinct_histogram().numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("counts", [(3, 1), (2, 2), (1, 0), (0, 2, 4)])
def test_two_pass_over_ranks_with_unequal_batch_counts(tmp_path, counts):
    """Every rank goes through the same number of collective rounds whatever its number of batches (even none); the global
    tables every rank ends up with are those of one process that sees all keys."""
    world = len(counts)
    mp.spawn(_two_pass_worker, args=(world, _free_port(), str(tmp_path), counts), nprocs=world, join=True)
    everything = [b for per_rank in _batches_of(0, counts) for b in per_rank]
    cells = torch.cat([c for c, _ in everything])
    keys = torch.unique(torch.cat([(c << UMI_BITS) | u for c, u in everything]))
    distinct = torch.zeros(N_CELLS, dtype=torch.int32)
    distinct.index_add_(0, keys >> UMI_BITS, torch.ones(keys.numel(), dtype=torch.int32))
    passing = distinct >= 5
    retained = 0
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), "tp_%d.npy" % r))
        assert (np.load(os.path.join(str(tmp_path), "tp_distinct_%d.npy" % r)) == distinct.numpy()).all()
        assert got[3] == counts[r]
        if counts[r]:
            assert got[0] == len(torch.unique(cells)) and got[1] == int(passing.sum())
        retained += got[2]
    assert retained == int(passing[cells].sum())


def _global_filter_worker(rank, world, port, out_dir, exact):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ssd_b200 import distributed as D
    dm = _StandIn(m=5)
    cells, umis = _global_filter_input(rank)
    dm.load(cells, umis)
    st = D.global_filter(dm, world, exact_table=exact == 1, capacity=64 if exact == 2 else 4096)
    np.save(os.path.join(out_dir, "gf_%d.npy" % rank), np.array([st["n_cells"], st["n_passing_cells"], st["n_retained"], getattr(dm, "n_undecided_sent", -1),
                                                                 int(cells.numel())]))
    np.save(os.path.join(out_dir, "gf_distinct_%d.npy" % rank), dm.distinct_histogram().numpy())
    np.save(os.path.join(out_dir, "gf_hist_%d.npy" % rank), dm.reads_histogram().numpy())
    dist.destroy_process_group()


def _global_filter_input(rank):
    """Cells of every kind for m = 5: decided by one rank alone, failing on the read count, and undecided ones (fewer than 5
    distinct UMIs on every rank, 5 or more reads in total -- some of which pass only once the ranks' UMIs are united, some of
    which still fail because the ranks hold the SAME UMIs)."""
    rng = np.random.default_rng(4242)
    pool = rng.choice(N_CELLS, size=600, replace=False)
    rng = np.random.default_rng(900 + rank)
    big = rng.choice(pool[:200], size=6000)                       # plenty of reads per cell: decided locally
    small = rng.choice(pool[200:500], size=700)                   # 2-3 reads per cell and rank: undecided or failing
    same = np.repeat(pool[500:600], 3)                            # the same three UMIs on every rank: 3 distinct globally, >= 5 reads
    cells = np.concatenate([big, small, same])
    umis = np.concatenate([rng.integers(0, 50, size=big.size), rng.integers(0, 1 << 16, size=small.size), np.tile(np.arange(3), 100)])
    return torch.from_numpy(cells).to(torch.int64), torch.from_numpy(umis).to(torch.int64)


@pytest.mark.parametrize("world,exact", [(2, 0), (3, 0), (2, 1), (2, 2)])
def test_global_filter_decides_like_one_process(tmp_path, world, exact):
    """global_filter's exchange of the UNDECIDED cells' keys only (and the exchange by owner rank behind exact_table=True): the
    same passing cells as one process that sees every key; only a fraction of the keys travels."""
    mp.spawn(_global_filter_worker, args=(world, _free_port(), str(tmp_path), exact), nprocs=world, join=True)
    parts = [_global_filter_input(r) for r in range(world)]
    cells = torch.cat([c for c, _ in parts])
    keys = torch.unique(torch.cat([(c << UMI_BITS) | u for c, u in parts]))
    distinct = torch.zeros(N_CELLS, dtype=torch.int32)
    distinct.index_add_(0, keys >> UMI_BITS, torch.ones(keys.numel(), dtype=torch.int32))
    passing = distinct >= 5
    assert 0 < int(passing.sum()) < len(torch.unique(cells))
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), "gf_%d.npy" % r))
        table = torch.from_numpy(np.load(os.path.join(str(tmp_path), "gf_distinct_%d.npy" % r)))
        hist = torch.from_numpy(np.load(os.path.join(str(tmp_path), "gf_hist_%d.npy" % r)))
        assert got[0] == len(torch.unique(cells)) and got[1] == int(passing.sum())
        assert got[2] == int(passing[parts[r][0]].sum())
        assert bool(((table >= 5) == passing).all())               # the table decides like the exact one ...
        assert bool((table <= distinct).all())                     # ... and never overstates a count
        assert int(hist.sum()) == cells.numel()                    # the read histogram ends up global
        if exact:                                                  # 1: by owner rank; 2: the packed list overflowed -> by owner rank
            assert bool((table == distinct).all())
        if exact != 1:
            assert 0 < got[3] < got[4] // 4                        # only the undecided cells' keys were sent
            assert (got[3] > 64) == (exact == 2) or exact == 0


def _plan_worker(rank, world, port, out_dir, paths):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ssd_b200 import filebatch
    filebatch.TILE_LOG2 = 12

    def count(path, offset, length):  # CPU stand-in for filebatch.device_tile_counts (the library counts on the GPU)
        with open(path, "rb") as fh:
            fh.seek(offset)
            buf = np.frombuffer(fh.read(length), dtype=np.uint8)
        t = 1 << filebatch.TILE_LOG2
        return np.array([int((buf[k:k + t] == 10).sum()) for k in range(0, len(buf), t)], dtype=np.uint32)

    def allgather(values):
        parts = [None] * world
        dist.all_gather_object(parts, [int(x) for x in values])
        return parts
    plan = filebatch.plan_shards_device(paths[0], paths[1], world, rank, count, count, allgather, target_bytes=90000)
    np.save(os.path.join(out_dir, "plan_%d.npy" % rank), np.array([b for per_rank in plan for b in per_rank], dtype=np.int64))
    dist.destroy_process_group()


def test_shard_plan_over_two_processes(tmp_path):
    """plan_shards_device with the real collective (all_gather_object over gloo, two processes): every rank counts only its own
    byte ranges, both derive the same plan, and every batch starts at the same record of both files."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import fastq_util as fu
    r1, r2 = fu.load_bundled()
    p1, p2 = tmp_path / "a.fastq", tmp_path / "b.fastq"
    p1.write_bytes(r1)
    p2.write_bytes(r2)
    mp.spawn(_plan_worker, args=(2, _free_port(), str(tmp_path), (str(p1), str(p2))), nprocs=2, join=True)
    plans = [np.load(os.path.join(str(tmp_path), "plan_%d.npy" % r)) for r in range(2)]
    assert (plans[0] == plans[1]).all() and len(plans[0]) >= 8
    flat = plans[0].tolist()
    assert flat[0][0] == 0 and flat[0][2] == 0 and flat[-1][0] + flat[-1][1] == len(r1) and flat[-1][2] + flat[-1][3] == len(r2)
    for (o1, n1, o2, n2), nxt in zip(flat, flat[1:] + [None]):
        assert r1[o1:o1 + n1].count(b"\n") == r2[o2:o2 + n2].count(b"\n") and r1[o1:o1 + n1].count(b"\n") % 4 == 0
        assert nxt is None or (o1 + n1 == nxt[0] and o2 + n2 == nxt[2])
