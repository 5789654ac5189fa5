"""Two ranks, two GPUs (NCCL): reads sharded by record range, global and exact -m decision.  The concatenation of
the ranks' outputs must equal the single-GPU output of the concatenated input, byte for byte."""
import os
import socket

import pytest

import fastq_util as fu

pytestmark = pytest.mark.gpu

WL = fu.load_whitelist()
N_PER_RANK = 60000


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir, filter_mode, own_stream=True, capacity=None):
    import numpy as np
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import ssd_b200
    from ssd_b200.distributed import global_filter
    dm = ssd_b200.Demultiplexer(WL, errors=1, min_count=10, device=rank, filter_mode=filter_mode)
    if own_stream:
        stream = torch.cuda.Stream()
        torch.cuda.set_stream(stream)
        dm.set_stream(stream.cuda_stream)
    # else: a default user -- torch on its default stream, the context on its own; global_filter has to order the two itself
    spec = ssd_b200.synth_spec(WL, N_PER_RANK, seed=77, first_index=1 + rank * N_PER_RANK, cell_pool=3000, p_n=0.01, p_dup_umi=0.3)
    d1, d2 = ssd_b200.synth_device(dm, spec)
    dm.demux_device(d1.data_ptr(), d1.numel(), d2.data_ptr(), d2.numel(), keep=(d1, d2))
    stats = global_filter(dm, world, capacity=capacity)
    out = torch.empty(stats["bytes_passing"] + 16, dtype=torch.uint8, device="cuda")
    n = dm.emit_device(ssd_b200.EMIT_PASSING, out.data_ptr(), out.numel())
    torch.cuda.synchronize()
    np.save(os.path.join(out_dir, "out_%d.npy" % rank), out[:n].cpu().numpy())
    np.save(os.path.join(out_dir, "stats_%d.npy" % rank), np.array([stats["n_cells"], stats["n_passing_cells"], stats["n_retained"]]))
    dist.destroy_process_group()


@pytest.mark.parametrize("filter_mode,own_stream,capacity", [("distinct_umi", True, None), ("reads", True, None), ("distinct_umi", False, None),
                                                             ("distinct_umi", True, 8)])
def test_two_ranks_equal_one(tmp_path, filter_mode, own_stream, capacity):
    """capacity=8: the packed list of undecided keys overflows, global_filter repeats the decision with the exchange by owner."""
    import numpy as np
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import ssd_b200
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), filter_mode, own_stream, capacity), nprocs=world, join=True)
    parts = [np.load(os.path.join(str(tmp_path), "out_%d.npy" % r)).tobytes() for r in range(world)]
    st = [np.load(os.path.join(str(tmp_path), "stats_%d.npy" % r)) for r in range(world)]
    # single GPU on the concatenated input
    r1 = b"".join(ssd_b200.synth_host(ssd_b200.synth_spec(WL, N_PER_RANK, seed=77, first_index=1 + r * N_PER_RANK, cell_pool=3000,
                                                          p_n=0.01, p_dup_umi=0.3))[0] for r in range(world))
    r2 = b"".join(ssd_b200.synth_host(ssd_b200.synth_spec(WL, N_PER_RANK, seed=77, first_index=1 + r * N_PER_RANK, cell_pool=3000,
                                                          p_n=0.01, p_dup_umi=0.3))[1] for r in range(world))
    dm = ssd_b200.Demultiplexer(WL, errors=1, min_count=10, filter_mode=filter_mode)
    out, stats = dm.run_host(r1, r2, ssd_b200.EMIT_PASSING)
    assert b"".join(parts) == out.tobytes()
    assert int(st[0][0]) == int(st[1][0]) == stats["n_cells"]               # global cell counts on every rank
    assert int(st[0][1]) == int(st[1][1]) == stats["n_passing_cells"]
    assert int(st[0][2]) + int(st[1][2]) == stats["n_retained"]             # retained records are per shard
    if filter_mode == "distinct_umi":
        from oracle import ssd_oracle_py as orc
        assert out.tobytes() == orc.demultiplex(r1, r2, WL, 1, 10).passing


N_BATCH = 25000
BATCHES_PER_RANK = 3


def _batch_worker(rank, world, port, out_dir):
    import numpy as np
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import ssd_b200
    from ssd_b200.distributed import demux_batches
    dm = ssd_b200.Demultiplexer(WL, errors=1, min_count=10, device=rank)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    dm.set_stream(stream.cuda_stream)

    def batches():
        for b in range(BATCHES_PER_RANK):
            first = 1 + (rank * BATCHES_PER_RANK + b) * N_BATCH
            yield ssd_b200.synth_device(dm, ssd_b200.synth_spec(WL, N_BATCH, seed=78, first_index=first, cell_pool=2000, p_n=0.01, p_dup_umi=0.3))

    parts = []
    total = demux_batches(dm, batches, world=world, sink=lambda out, st: parts.append(out.cpu().numpy().tobytes()))
    torch.cuda.synchronize()
    np.save(os.path.join(out_dir, "bout_%d.npy" % rank), np.frombuffer(b"".join(parts), dtype=np.uint8))
    np.save(os.path.join(out_dir, "bstats_%d.npy" % rank), np.array([total["n_cells"], total["n_passing_cells"], total["n_retained"]]))
    dist.destroy_process_group()


def test_two_ranks_three_batches_each_equal_one_call(tmp_path):
    """BASELINE config 4 in small: the input sharded over two ranks and cut into three batches per rank (two passes over the
    batches, keys sent to the owner rank batch by batch) gives the bytes of one call on the whole input."""
    import numpy as np
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import ssd_b200
    world = 2
    mp.spawn(_batch_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(os.path.join(str(tmp_path), "bout_%d.npy" % r)).tobytes() for r in range(world)]
    st = [np.load(os.path.join(str(tmp_path), "bstats_%d.npy" % r)) for r in range(world)]
    n = world * BATCHES_PER_RANK * N_BATCH
    r1, r2 = ssd_b200.synth_host(ssd_b200.synth_spec(WL, n, seed=78, cell_pool=2000, p_n=0.01, p_dup_umi=0.3))
    dm = ssd_b200.Demultiplexer(WL, errors=1, min_count=10)
    out, stats = dm.run_host(r1, r2, ssd_b200.EMIT_PASSING)
    assert b"".join(parts) == out.tobytes()
    assert int(st[0][0]) == int(st[1][0]) == stats["n_cells"]
    assert int(st[0][1]) == int(st[1][1]) == stats["n_passing_cells"]
    assert int(st[0][2]) + int(st[1][2]) == stats["n_retained"]


def _cli_worker(rank, world, port, workdir, batch_bytes, extra=()):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                      SSD_BATCH_BYTES=str(batch_bytes))
    os.chdir(workdir)
    import contextlib
    import io
    import runpy
    import sys
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "split-seq_demultiplexing_b200")
    sys.path.insert(0, pkg)
    sys.argv = ["splitseqdemultiplex_0.2.3.py", "-n", "2", "-e", "1", "-m", "10", "-1", "a", "-2", "b", "-3", "c", "-f", "F.fastq", "-r", "R.fastq",
                "-o", "results", "-a", "", "-b", "100000", "-p", "True", "-l", "20000"] + list(extra)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        runpy.run_path(os.path.join(pkg, "splitseqdemultiplex_0.2.3.py"), run_name="__main__")
    open("stdout_%d.txt" % rank, "w").write(buf.getvalue())
    import torch.distributed as dist
    dist.destroy_process_group()


@pytest.mark.parametrize("batch_bytes", [100000, 10**9])
def test_cli_under_two_ranks_matches_the_reference_digest(tmp_path, batch_bytes):
    """The drop-in CLI launched with one process per GPU: the two ranks share the bundled pair of files (8 batches of whole
    records, or one batch and a rank without work), -m is decided globally, rank 0 strings the parts together; the result
    is the reference's golden digest and its printed summary."""
    import json
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    fu.write_whitelist_files(str(tmp_path))
    r1, r2 = fu.load_bundled()
    (tmp_path / "F.fastq").write_bytes(r1)
    (tmp_path / "R.fastq").write_bytes(r2)
    mp.spawn(_cli_worker, args=(2, _free_port(), str(tmp_path), batch_bytes), nprocs=2, join=True)
    gold = json.load(open(os.path.join(fu.GOLDEN_DIR, "bundled_digests.json")))["e1_m10"]
    out = (tmp_path / "results" / "MergedCells_passing.fastq").read_bytes()
    assert fu.canonical_sha(out) == gold["passing_sha"]
    assert not [f for f in os.listdir(str(tmp_path / "results")) if ".part" in f]
    said = (tmp_path / "stdout_0.txt").read_text()
    assert gold["stdout_tail"][0] in said and gold["stdout_tail"][1] in said
    assert (tmp_path / "stdout_1.txt").read_text() == ""
    # and the bytes of the single-GPU call on the same files
    import ssd_b200
    dm = ssd_b200.Demultiplexer(WL, errors=1, min_count=10)
    single = tmp_path / "single.fastq"
    dm.run_files(str(tmp_path / "F.fastq"), str(tmp_path / "R.fastq"), str(single), ssd_b200.EMIT_PASSING, append=False)
    assert out == single.read_bytes()


def test_cli_split_and_cell_table_under_two_ranks(tmp_path):
    """`-v split` and `--cellTable` with one process per GPU: the ranks bucket their record ranges and append to the per-cell
    files in rank order (the files of the single-GPU call, byte for byte); the per-cell table written by rank 0 after the
    merged run is the global one."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import ssd_b200
    from ssd_b200 import hostlogic
    fu.write_whitelist_files(str(tmp_path))
    r1, r2 = fu.load_bundled()
    (tmp_path / "F.fastq").write_bytes(r1)
    (tmp_path / "R.fastq").write_bytes(r2)
    mp.spawn(_cli_worker, args=(2, _free_port(), str(tmp_path), 10**9, ("-v", "split")), nprocs=2, join=True)
    got = {f: (tmp_path / "results" / f).read_bytes() for f in os.listdir(str(tmp_path / "results"))}
    dm = ssd_b200.Demultiplexer(WL, errors=1, min_count=10)
    _, stats = dm.run_host(r1, r2, ssd_b200.EMIT_PASSING)
    single = tmp_path / "single_split"
    dm.write_split_files(str(single), hostlogic.load_round_names(str(tmp_path)), stats)
    want = {f: (single / f).read_bytes() for f in os.listdir(str(single))}
    assert len(want) == 83 and got == want
    assert "# of results files [83]" in (tmp_path / "stdout_0.txt").read_text()
    # merged run with the per-cell table
    import shutil
    shutil.rmtree(str(tmp_path / "results"))
    mp.spawn(_cli_worker, args=(2, _free_port(), str(tmp_path), 10**9, ("--cellTable", "cells.tsv")), nprocs=2, join=True)
    dm.write_cell_table(str(tmp_path / "cells_single.tsv"))
    assert (tmp_path / "cells.tsv").read_bytes() == (tmp_path / "cells_single.tsv").read_bytes()
