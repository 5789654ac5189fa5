"""CPU tests: the C-ABI library loads and exports every symbol include/ssd_b200.h declares, the host-side
logic (position learner, whitelist files, FASTQ splitter, synthetic generator twin) behaves like the reference."""
import ctypes
import json
import os
import re
import sys

import pytest

import fastq_util as fu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "split-seq_demultiplexing_b200")
WL = fu.load_whitelist()


@pytest.fixture(scope="module")
def ssd():
    import ssd_b200
    if ssd_b200.needs_build():
        ssd_b200.build()
    return ssd_b200


def test_library_exports_every_declared_symbol(ssd):
    header = open(os.path.join(ROOT, "include", "ssd_b200.h")).read()
    declared = set(re.findall(r"\b(ssd_[a-z0-9_]+)\s*\(", header))
    declared -= {"ssd_config", "ssd_stats", "ssd_synth_spec", "ssd_ctx"}
    lib = ctypes.CDLL(ssd.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert set(ssd.exported_symbols()) <= declared, set(ssd.exported_symbols()) - declared
    assert len(declared) >= 25
    assert lib.ssd_abi_version() == 1


def test_no_cpu_fallback_without_gpu(ssd):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ssd.Demultiplexer(WL)


def test_bad_config_is_rejected_before_touching_the_gpu(ssd):
    with pytest.raises(ValueError):
        ssd.Demultiplexer(WL, positions={"umi": (5, 2)})
    with pytest.raises(ValueError):
        ssd.Demultiplexer(WL, positions={"bc1": (-2, 6)})


def test_learner_matches_golden_cases(ssd):
    from ssd_b200 import hostlogic
    with open(os.path.join(fu.GOLDEN_DIR, "learner_cases.json")) as fh:
        cases = json.load(fh)
    for c in cases:
        seq = "T" * c["shift"] + "ACGTACGTAC" + WL[3] + fu.LINKER_3_2 + WL[7] + fu.LINKER_2_1 + WL[50] + "ACGT" * 10
        recs = "".join("@l%d %d/2\n%s\n+\n%s\n" % (i, i, seq, "A" * len(seq)) for i in range(300))
        lines = [l + "\n" for l in recs.split("\n")[:1001]]
        pos = hostlogic.learn_positions(lines)
        report = hostlogic.positions_report(pos)
        for printed in c["printed"]:
            assert printed in report


def test_learner_bundled_equals_defaults_and_oracle(ssd):
    from oracle import ssd_oracle_py as orc
    from ssd_b200 import hostlogic
    r2 = fu.load_bundled()[1]
    lines = [l.decode() + "\n" for l in r2.split(b"\n")[:1001]]
    pos = hostlogic.learn_positions(lines)
    assert pos["_found"] == orc.learn_positions("".join(lines).encode()) == (80, 56, 18)
    pos.pop("_found")
    assert pos == hostlogic.DEFAULT_POSITIONS


def test_learner_empty_raises_value_error(ssd):
    from ssd_b200 import hostlogic
    with pytest.raises(ValueError):
        hostlogic.learn_positions(["@a\n", "ACGT\n", "+\n", "EEEE\n"])


def test_whitelist_files(ssd, tmp_path):
    from ssd_b200 import hostlogic
    fu.write_whitelist_files(str(tmp_path))
    assert hostlogic.load_whitelist(str(tmp_path)) == WL
    os.remove(os.path.join(str(tmp_path), hostlogic.ROUND3_FILE))
    with pytest.raises(FileNotFoundError):   # the reference opens all three files (V2:56-68)
        hostlogic.load_whitelist(str(tmp_path))


@pytest.mark.parametrize("n", [1, 3, 4, 8])
def test_split_functions_cut_like_the_reference(ssd, tmp_path, monkeypatch, n):
    from ssd_b200 import hostlogic
    sys.path.insert(0, PKG)
    import Splitseq_fun_lib as lib
    r1, r2 = fu.load_bundled()
    monkeypatch.chdir(tmp_path)
    open("F.fastq", "wb").write(r1)
    open("R.fastq", "wb").write(r2)
    lib.split_fastqF_fun(n, "F.fastq", 20000)
    lib.split_fastqR_fun(n, "R.fastq", 20000)
    size = hostlogic.tuned_split_size(20000, n)
    assert size % 4 == 0 and size * n >= 20000 and size == {1: 20000, 3: 6668, 4: 5000, 8: 2500}[n]
    got = b"".join(open("split_fastq_F_%d" % k, "rb").read() for k in range(n))
    assert got == r1
    assert all(open("split_fastq_R_%d" % k, "rb").read().count(b"\n") == min(size, 20000 - k * size) for k in range(n))
    assert open(hostlogic.LEARNER_FILE, "rb").read() == b"".join(l + b"\n" for l in r2.split(b"\n")[:1001])


def test_synthetic_generator_twin_is_deterministic_and_well_formed(ssd):
    spec = ssd.synth_spec(WL, 3000, seed=11, first_index=998, p_tie=0.02, p_trunc=0.02)
    a1, a2 = ssd.synth_host(spec)
    b1, b2 = ssd.synth_host(ssd.synth_spec(WL, 3000, seed=11, first_index=998, p_tie=0.02, p_trunc=0.02))
    assert (a1, a2) == (b1, b2)
    assert ssd.synth_sizes(spec) == (len(a1), len(a2))
    l1, l2 = a1.split(b"\n"), a2.split(b"\n")
    assert len(l1) == len(l2) == 4 * 3000 + 1
    assert l1[0] == b"@SYN.998 998/1" and l2[0] == b"@SYN.998 998/2" and l1[8] == b"@SYN.1000 1000/1"
    assert all(len(l1[4 * k + 1]) == 150 == len(l1[4 * k + 3]) for k in range(3000))
    assert all(len(l2[4 * k + 1]) == len(l2[4 * k + 3]) for k in range(3000))
    assert any(len(l2[4 * k + 1]) < 94 for k in range(3000))            # truncated read 2 present
    assert sum(l2[4 * k + 1][18:48] == fu.LINKER_3_2.encode() for k in range(3000)) > 2500


def test_oracle_on_synthetic_is_sane(ssd):
    from oracle import ssd_oracle_py as orc
    spec = ssd.synth_spec(WL, 20000, seed=5, cell_pool=500)
    r1, r2 = ssd.synth_host(spec)
    res = orc.demultiplex(r1, r2, WL, 1, 10)
    assert res.n_records == 20000 and 0.8 * 20000 < res.n_matched < 20000
    assert 0 < res.n_passing_cells <= res.n_cells <= 500 + 2000


# ---- record-aligned cut points of FASTQ files (ssd_fastq_record_cuts; host only) ------------------------------------
def _record_starts(data: bytes):
    """Byte offsets at which records start: 0 and the byte after every 4th newline."""
    import numpy as np
    nl = np.flatnonzero(np.frombuffer(data, dtype=np.uint8) == 10)
    starts = np.concatenate([[0], nl[3::4] + 1])
    n_lines = len(nl) + (1 if data and data[-1] != 10 else 0)
    return starts, n_lines // 4


def _check_cuts(ssd, tmp_path, data: bytes, target: int):
    import numpy as np
    from ssd_b200 import filebatch
    p = tmp_path / "x.fastq"
    p.write_bytes(data)
    starts, n_records = _record_starts(data)
    rec, off, total = filebatch.record_cuts(str(p), target_bytes=target)
    assert total == n_records
    assert (rec[0], off[0]) == (0, 0) and (rec[-1], off[-1]) == (n_records, len(data))
    # expected: after a cut at byte b, the next cut is the first record start at or after b + target (and before the end)
    want, prev = [], 0
    while True:
        k = int(np.searchsorted(starts, prev + target))
        if k >= len(starts) or int(starts[k]) >= len(data):
            break
        want.append((k, int(starts[k])))
        prev = int(starts[k])
    inner = list(zip(rec[1:-1], off[1:-1]))
    assert inner == want
    # the same records located again by index (how the mate file is cut), plus indices past the end
    asked = [r for r, _ in inner][:2000] + [n_records + 5]
    rec2, off2, total2 = filebatch.record_cuts(str(p), records=asked)
    assert rec2 == asked and total2 == n_records
    assert off2[:-1] == [o for _, o in inner][:2000] and off2[-1] == len(data)


@pytest.mark.parametrize("target", [1, 100, 333, 4096, 70000, 10**9])
def test_record_cuts_on_the_bundled_file(ssd, tmp_path, target):
    r1, _ = fu.load_bundled()
    _check_cuts(ssd, tmp_path, r1[:200000], target)


@pytest.mark.parametrize("data", [
    b"", b"\n", b"@a\nAC\n+\nEE", b"@a\nAC\n+\nEE\n", b"@a\nAC\n+\nEE\n@b\nGG\n+\n!!", b"@a\r\nAC\r\n+\r\nEE\r\n@b\r\nGG\r\n+\r\n!!\r\n",
    b"\n\n\n\n\n\n\n\n\n", b"@a\nAC\n+\nEE\n" * 50000 + b"@tail\nA"],
    ids=["empty", "newline", "one-open", "one", "two-open", "crlf", "blank-lines", "many"])
@pytest.mark.parametrize("target", [1, 7, 64, 65536, 70001])
def test_record_cuts_edge_cases(ssd, tmp_path, data, target):
    _check_cuts(ssd, tmp_path, data, target)


def test_record_cuts_across_read_buffers(ssd, tmp_path):
    """22 MB: the cutter reads 8 MiB at a time and counts 64 KiB pieces; cuts and record lookups across those edges."""
    data = b"@a\nAC\n+\nEE\n" * 1_000_000 + b"@r 1\n" + b"A" * 70000 + b"\n+\n" + b"E" * 70000 + b"\n" + b"@a\nAC\n+\nEE\n" * 1_000_000
    for target in (65536, (8 << 20) - 3, 5_000_000):
        _check_cuts(ssd, tmp_path, data, target)


def test_pair_batches_cover_both_files_with_the_same_records(ssd, tmp_path):
    from ssd_b200 import filebatch
    r1, r2 = fu.load_bundled()
    p1, p2 = tmp_path / "a.fastq", tmp_path / "b.fastq"
    p1.write_bytes(r1)
    p2.write_bytes(r2 + b"@extra 1/2\nACGT\n+\nEEEE\n")                 # a surplus record ends up in the last batch
    batches = filebatch.pair_batches(str(p1), str(p2), target_bytes=50000)
    assert len(batches) > 10
    assert batches[0][0] == 0 and batches[0][2] == 0
    for (o1, n1, o2, n2), nxt in zip(batches, batches[1:]):
        assert o1 + n1 == nxt[0] and o2 + n2 == nxt[2]
        assert r1[o1:o1 + n1].count(b"\n") == r2[o2:o2 + n2].count(b"\n") and r1[o1:o1 + n1].count(b"\n") % 4 == 0
    o1, n1, o2, n2 = batches[-1]
    assert o1 + n1 == len(r1) and o2 + n2 == os.path.getsize(str(p2))
    with pytest.raises(FileNotFoundError):
        filebatch.pair_batches(str(tmp_path / "missing.fastq"), str(p2))


def test_round_names_and_legacy_split_file_name(ssd, tmp_path):
    from ssd_b200 import hostlogic
    fu.write_whitelist_files(str(tmp_path))
    names = hostlogic.load_round_names(str(tmp_path))
    wl = hostlogic.load_whitelist(str(tmp_path))
    assert [len(n) for n in names] == [len(wl)] * 3
    assert all(a[8:16] == w and b[0:8] == w and c[0:8] == w for a, b, c, w in zip(*names, wl))   # line i of every file is entry i
    # demultiplex_using_barcodes.py:364: round1 + "-" + round2 + "-" + round3 + ".fastq"
    assert hostlogic.split_file_name(names, 0, 1, 2) == names[0][0] + "-" + names[1][1] + "-" + names[2][2] + ".fastq"
    assert (len(names[0][0]), len(names[1][0]), len(names[2][0])) == (16, 13, 14)


def test_whitelist_from_the_given_files(ssd, tmp_path):
    from ssd_b200 import hostlogic
    fu.write_whitelist_files(str(tmp_path))
    paths = [str(tmp_path / n) for n in (hostlogic.ROUND1_FILE, hostlogic.ROUND2_FILE, hostlogic.ROUND3_FILE)]
    assert hostlogic.load_whitelist_from(*paths) == hostlogic.load_whitelist(str(tmp_path)) == WL
    other = tmp_path / "r3_other.txt"
    other.write_text("".join("ACGTACGTGTGGCC\n" for _ in WL))
    with pytest.raises(ValueError, match="do not carry"):
        hostlogic.load_whitelist_from(paths[0], paths[1], str(other))
    w1, w2, w3 = hostlogic.load_whitelists_from(paths[0], paths[1], str(other))   # one list per round
    assert w1 == w2 == WL and w3 == ["ACGTACGT"] * len(WL)


def test_record_cuts_error_codes(ssd, tmp_path):
    import ctypes as C
    from ssd_b200 import _lib as L
    lib = L.load()
    p = tmp_path / "x.fastq"
    p.write_bytes(b"@a\nAC\n+\nEE\n" * 1000)
    rec, off = (C.c_uint64 * 4)(), (C.c_uint64 * 4)()
    n, total = C.c_uint64(), C.c_uint64()
    # more cuts than the arrays hold: SSD_ERANGE (the Python wrapper then retries with larger arrays)
    assert lib.ssd_fastq_record_cuts(str(p).encode(), 100, None, 0, rec, off, 4, C.byref(n), C.byref(total)) == L.ERANGE
    assert lib.ssd_fastq_record_cuts(str(tmp_path / "missing").encode(), 100, None, 0, rec, off, 4, C.byref(n), C.byref(total)) == L.EIO
    assert lib.ssd_fastq_record_cuts(str(p).encode(), 0, None, 0, rec, off, 4, C.byref(n), C.byref(total)) == L.EINVAL
    from ssd_b200 import filebatch
    r, o, t = filebatch.record_cuts(str(p), target_bytes=100)          # grows its arrays until they fit
    assert t == 1000 and len(r) > 4 and o[-1] == 11000


def _split_both_ways(lib, monkeypatch, tmp_path, data: bytes, n: int, length: int):
    """split_fastqF_fun through the byte-range path (when it applies) and line by line: (chunk files, took the fast path)."""
    import glob
    out = []
    for slow in (False, True):
        d = tmp_path / ("slow" if slow else "fast")
        d.mkdir(exist_ok=True)
        monkeypatch.chdir(d)
        for f in glob.glob("split_fastq_F_*"):
            os.remove(f)
        open("in.fastq", "wb").write(data)
        if slow:
            monkeypatch.setenv("SSD_SPLIT_LINE_BY_LINE", "1")
        else:
            monkeypatch.delenv("SSD_SPLIT_LINE_BY_LINE", raising=False)
        lib.split_fastqF_fun(n, "in.fastq", length)
        out.append({f: open(f, "rb").read() for f in sorted(glob.glob("split_fastq_F_*"))})
    return out


@pytest.mark.parametrize("n", [1, 3, 8])
def test_split_by_byte_ranges_equals_line_by_line(ssd, tmp_path, monkeypatch, n):
    import ctypes as C
    import importlib
    if PKG not in sys.path:
        sys.path.insert(0, PKG)
    lib = importlib.import_module("Splitseq_fun_lib")
    L = importlib.import_module("ssd_b200._lib")
    from ssd_b200 import hostlogic
    r1, _ = fu.load_bundled()
    rec = b"@a 1/1\nACGT\n+\nEEEE\n"
    cases = {
        "bundled": (r1, 20000, 1),
        "tiny (fewer chunks than asked for)": (rec * 10, 40, 1),
        "partial last record": (rec * 9 + b"@b\nAC\n", 38, 1),
        "crlf": (rec.replace(b"\n", b"\r\n") * 12, 48, 0),
        "trailing blank": (rec * 5 + b"@c 1/1 \nACGT\n+\nEEEE\n" + rec * 6, 48, 0),
        "leading tab": (rec * 5 + b"@c\n\tACGT\n+\nEEEE\n" + rec * 6, 48, 0),
        "no final newline": (rec * 12 + b"@d\nAC\n+\nEE", 52, 0),
        "non-ascii": (rec * 6 + "@é\nAC\n+\nEE\n".encode() + rec * 5, 48, 0),
        "blank lines": (rec * 3 + b"\n\n\n\n" + rec * 8, 48, 1),
        "empty": (b"", 0, 1),
    }
    for name, (data, length, plain) in cases.items():
        p = tmp_path / "probe.fastq"
        p.write_bytes(data)
        got, nl = C.c_int(-1), C.c_uint64(0)
        assert L.load().ssd_fastq_lines_are_stripped(str(p).encode(), C.byref(got), C.byref(nl)) == 0
        assert got.value == plain, name
        if plain:
            assert nl.value == data.count(b"\n"), name
        if length:
            took = []
            real = lib._split_by_ranges
            monkeypatch.setattr(lib, "_split_by_ranges", lambda *a: took.append(real(*a)) or took[-1])
            fast, slow = _split_both_ways(lib, monkeypatch, tmp_path, data, n, length)
            monkeypatch.setattr(lib, "_split_by_ranges", real)
            assert fast == slow, name
            assert took == [bool(plain)], name                             # the byte-range path ran exactly when the file allows it
            assert len(slow) == min(n, -(-data.count(b"\n") // hostlogic.tuned_split_size(length, n))) or not plain, name


def _splitter_cases():
    with open(os.path.join(fu.GOLDEN_DIR, "splitter_cases.json")) as fh:
        return json.load(fh)


@pytest.mark.parametrize("line_by_line", [False, True], ids=["byte-ranges-when-possible", "line-by-line"])
def test_splitter_matches_the_reference_golden(ssd, tmp_path, monkeypatch, line_by_line):
    """The chunk files (names, sizes, SHA-256) and the learner sample of the UNMODIFIED reference's splitter (LIB:6-110,
    tests/golden/make_golden.py::make_splitter_cases) for 8 inputs x 4 core counts x both functions, through the drop-in
    module with and without its byte-range path."""
    import contextlib
    import hashlib
    import importlib
    import io
    sys.path.insert(0, os.path.join(fu.GOLDEN_DIR))
    inputs = {name: (data, length) for name, data, length in importlib.import_module("make_golden").splitter_inputs()}
    if PKG not in sys.path:
        sys.path.insert(0, PKG)
    lib = importlib.import_module("Splitseq_fun_lib")
    if line_by_line:
        monkeypatch.setenv("SSD_SPLIT_LINE_BY_LINE", "1")
    else:
        monkeypatch.delenv("SSD_SPLIT_LINE_BY_LINE", raising=False)
    cases = _splitter_cases()
    assert len(cases) == 64
    for k, case in enumerate(cases):
        d = tmp_path / ("c%d" % k)
        d.mkdir()
        monkeypatch.chdir(d)
        data, length = inputs[case["input"]]
        assert length == case["length"]
        open("in.fastq", "wb").write(data)
        with contextlib.redirect_stdout(io.StringIO()):
            getattr(lib, case["fn"])(case["n"], "in.fastq", length)
        prefix = "split_fastq_F" if case["fn"] == "split_fastqF_fun" else "split_fastq_R"
        got = {f: [os.path.getsize(f), hashlib.sha256(open(f, "rb").read()).hexdigest()] for f in sorted(os.listdir(".")) if f.startswith(prefix)}
        assert got == case["files"], (case["input"], case["n"], case["fn"])
        if "learner_sha" in case:
            assert hashlib.sha256(open("position_learner_fastqr.fastq", "rb").read()).hexdigest() == case["learner_sha"]


def test_format_cell_table(ssd):
    from ssd_b200 import hostlogic
    a, b = WL[1] + WL[2] + WL[3], WL[0] + WL[0] + WL[95]
    text = hostlogic.format_cell_table({a: (12, 10), b: (9, 9)}, "10")
    rows = text.split("\n")
    assert rows[0] == "cell\treads\tdistinct_umis\tpassing" and rows[-1] == ""
    assert rows[1:3] == sorted(["%s\t12\t10\t1" % a, "%s\t9\t9\t0" % b])


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_plan_shards_deals_contiguous_runs_of_whole_records(ssd, tmp_path, monkeypatch, world):
    from ssd_b200 import filebatch
    r1, r2 = fu.load_bundled()
    p1, p2 = tmp_path / "a.fastq", tmp_path / "b.fastq"
    p1.write_bytes(r1)
    p2.write_bytes(r2)
    for batch_bytes in (10**9, 60000):
        monkeypatch.setenv("SSD_BATCH_BYTES", str(batch_bytes))
        plan = filebatch.plan_shards(str(p1), str(p2), world)
        assert len(plan) == world
        flat = [b for per_rank in plan for b in per_rank]
        assert flat[0][0] == 0 and flat[0][2] == 0
        assert all(a[0] + a[1] == b[0] and a[2] + a[3] == b[2] for a, b in zip(flat, flat[1:]))      # file order, no gaps
        assert flat[-1][0] + flat[-1][1] == len(r1) and flat[-1][2] + flat[-1][3] == len(r2)
        assert all(n1 <= max(batch_bytes, 400) + 400 for _, n1, _, _ in flat)                         # the cut target holds (+ one record)
        for o1, n1, o2, n2 in flat:
            assert r1[o1:o1 + n1].count(b"\n") == r2[o2:o2 + n2].count(b"\n") and r1[o1:o1 + n1].count(b"\n") % 4 == 0
        sizes = [len(per_rank) for per_rank in plan]
        assert max(sizes) - min(s for s in sizes if s or world > len(flat)) <= max(sizes)              # runs, not a deal of single batches
        assert sizes == sorted(sizes, reverse=True) and (len(flat) < world or min(sizes) >= 1 or sum(sizes) == len(flat))
        if batch_bytes == 10**9 and world > 1:
            assert len(flat) >= min(world, 2)                                                          # cut into about one batch per rank


def test_format_cell_umi_table():
    import numpy as np
    from ssd_b200 import hostlogic
    wl = ["AAAAAAAA", "CCCCCCCC", "GGGGGGGG"]
    cells = np.array([0, 5, 26], dtype=np.uint32)            # (0,0,0), (0,1,2), (2,2,2)
    umis = np.array([b"ACGTACGTAC", b"ANNTACGTAC", b"ACGTA"], dtype="S10")
    reads = np.array([1, 12, 1234567], dtype=np.uint32)
    text = hostlogic.format_cell_umi_table(wl, cells, umis, reads)
    assert text == (b"cell\tumi\treads\n"
                    b"AAAAAAAAAAAAAAAAAAAAAAAA\tACGTACGTAC\t1\n"
                    b"AAAAAAAACCCCCCCCGGGGGGGG\tANNTACGTAC\t12\n"
                    b"GGGGGGGGGGGGGGGGGGGGGGGG\tACGTA\t1234567\n")
    assert hostlogic.format_cell_umi_table(wl, cells[:0], umis[:0], reads[:0]) == b"cell\tumi\treads\n"


def _numpy_tile_counts(path, offset, length):
    """CPU stand-in for filebatch.device_tile_counts (the library counts on the GPU): newlines per tile of a byte range."""
    import numpy as np
    from ssd_b200 import filebatch
    with open(path, "rb") as fh:
        fh.seek(offset)
        buf = np.frombuffer(fh.read(length), dtype=np.uint8)
    t = 1 << filebatch.TILE_LOG2
    return np.array([int((buf[k:k + t] == 10).sum()) for k in range(0, len(buf), t)], dtype=np.uint32)


def _plan_on_threads(paths, world, target_bytes=None):
    """plan_shards_device on `world` threads standing in for the ranks (the all-gather is a barrier over a shared list)."""
    import threading
    from ssd_b200 import filebatch
    barrier, slots, plans, errors = threading.Barrier(world), [None] * world, [None] * world, []

    def run(rank):
        def allgather(values):
            slots[rank] = list(values)
            barrier.wait()
            out = [list(x) for x in slots]
            barrier.wait()
            return out
        try:
            plans[rank] = filebatch.plan_shards_device(paths[0], paths[1], world, rank, _numpy_tile_counts, _numpy_tile_counts, allgather, target_bytes)
        except Exception as e:  # noqa: BLE001
            errors.append(e)
            barrier.abort()
    th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors, errors
    assert all(p == plans[0] for p in plans)  # every rank derives the same plan
    return plans[0]


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("ending", ["lf", "crlf", "no_final_newline", "surplus_r2"])
def test_plan_shards_device_cuts_both_files_at_the_same_records(tmp_path, monkeypatch, world, ending):
    """The shard phase without a whole-file pass: every rank counts the newlines of its byte ranges, line indices follow from
    the gathered totals, batches start at the same RECORD in both files (whose records differ in size)."""
    from ssd_b200 import filebatch
    monkeypatch.setattr(filebatch, "TILE_LOG2", 12)  # several tiles per range on these small files
    r1, r2 = fu.load_bundled()
    if ending == "crlf":
        r1, r2 = r1.replace(b"\n", b"\r\n"), r2.replace(b"\n", b"\r\n")
    if ending == "no_final_newline":
        r1, r2 = r1[:-1], r2[:-1]
    if ending == "surplus_r2":
        r1 = b"\n".join(r1.split(b"\n")[:4 * 4000]) + b"\n"
    p1, p2 = tmp_path / "a.fastq", tmp_path / "b.fastq"
    p1.write_bytes(r1)
    p2.write_bytes(r2)
    starts1 = [0] + [i + 1 for i, c in enumerate(r1) if c == 10]
    starts2 = [0] + [i + 1 for i, c in enumerate(r2) if c == 10]
    for target in (None, 70000):
        plan = _plan_on_threads((str(p1), str(p2)), world, target)
        assert len(plan) == world and len({len(x) for x in plan}) == 1      # the same number of batches on every rank
        flat = [b for per_rank in plan for b in per_rank]
        assert flat[0][0] == 0 and flat[0][2] == 0
        assert all(a[0] + a[1] == b[0] and a[2] + a[3] == b[2] for a, b in zip(flat, flat[1:]))
        assert flat[-1][0] + flat[-1][1] == len(r1) and flat[-1][2] + flat[-1][3] == len(r2)
        for o1, _, o2, _ in flat[1:]:
            if o1 < len(r1):
                k = starts1.index(o1)                                            # a line start ...
                assert k % 4 == 0                                                # ... of a record ...
                assert o2 == (starts2[k] if k < len(starts2) else len(r2))       # ... and the same record of the mate file
            else:
                assert o2 == len(r2) or ending == "surplus_r2"
        if target:
            assert len(flat) > world and all(n1 <= 2 * target for _, n1, _, _ in flat)
        if world > 1 and ending == "lf" and target is None:
            sizes = [n1 for _, n1, _, _ in flat]
            assert max(sizes) - min(sizes) < 2000                                # equal byte ranges, moved by less than a record or two


def test_lone_carriage_returns_are_told_from_crlf(ssd):
    """The reference reads with universal newlines (a lone '\\r' ends a line: checked against it in the authoring container, the
    all-'\\r' copy of a FASTQ pair demultiplexes like the original); the scans break lines at '\\n' only and find no record in
    such a file, which the host layer reports instead of an empty result (Demultiplexer._no_records_check)."""
    from ssd_b200 import hostlogic
    from ssd_b200.engine import MalformedFastq
    assert hostlogic.lone_cr_in_head(b"@a 1/1\rACGT\r+\rEEEE\r@b 2/1\r")
    assert hostlogic.lone_cr_in_head(b"@a\r\nAC\rGT\r\n")
    assert not hostlogic.lone_cr_in_head(b"@a 1/1\r\nACGT\r\n+\r\nEEEE\r\n")
    assert not hostlogic.lone_cr_in_head(b"@a 1/1\nACGT\n+\nEEEE\n")
    assert not hostlogic.lone_cr_in_head(b"@a 1/1\nACGT\n+\nEEEE\r")   # the head may end between '\\r' and '\\n'
    assert not hostlogic.lone_cr_in_head(b"")
    hostlogic.reject_lone_cr(b"@a\n", b"")
    with pytest.raises(MalformedFastq) as ei:
        hostlogic.reject_lone_cr(b"@a\nAC\n", b"@a\rAC\r+\rEE\r")
    assert "input 2" in str(ei.value) and isinstance(ei.value, ValueError)

    class St:
        n_records_r1 = 0
        n_records_r2 = 0
    dm = ssd.Demultiplexer.__new__(ssd.Demultiplexer)        # the check itself needs no context
    dm._no_records_check(St, lambda: (b"", b""))
    with pytest.raises(MalformedFastq):
        dm._no_records_check(St, lambda: (b"@a\rAC\r+\rEE\r", b"@a\rAC\r+\rEE\r"))
    St.n_records_r2 = 3                                       # records were found: the heads are not even asked for
    dm._no_records_check(St, lambda: 1 / 0)
