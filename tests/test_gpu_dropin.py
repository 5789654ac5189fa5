"""GPU tests of the drop-in surface: the reference's module / function / CLI names (V2.demultiplex_fun,
V2.calc_demux_results, splitseqdemultiplex_0.2.3.py) backed by the CUDA library, checked against the golden
vectors of the unmodified reference."""
import contextlib
import importlib
import io
import json
import os
import runpy
import sys

import pytest

import fastq_util as fu

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "split-seq_demultiplexing_b200")
WL = fu.load_whitelist()


@pytest.fixture()
def workdir(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    fu.write_whitelist_files(str(tmp_path))
    r1, r2 = fu.load_bundled()
    open("F.fastq", "wb").write(r1)
    open("R.fastq", "wb").write(r2)
    os.makedirs("out")
    if PKG not in sys.path:
        sys.path.insert(0, PKG)
    return tmp_path


@pytest.fixture(scope="module")
def digests():
    with open(os.path.join(fu.GOLDEN_DIR, "bundled_digests.json")) as fh:
        return json.load(fh)


@pytest.mark.parametrize("e,m", [(1, 10), (2, 10), (0, 1), (3, 10)])
def test_reference_function_surface(workdir, digests, e, m):
    V2 = importlib.import_module("DemultiplexUsingBarcodes_New_V2")
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        V2.demultiplex_fun("F.fastq", "R.fastq", "out", 100000, e, True, 10, False, True)
        V2.calc_demux_results("out", True, str(m))        # CLI3 passes -m through as a string
    gold = digests["e%d_m%d" % (e, m)]
    assert fu.canonical_sha(open("out/MergedCells_1.fastq", "rb").read()) == gold["merged1_sha"]
    assert fu.canonical_sha(open("out/MergedCells_passing.fastq", "rb").read()) == gold["passing_sha"]
    assert buf.getvalue().strip().split("\n")[-2:] == gold["stdout_tail"]


def test_position_detection_path(workdir, digests):
    V2 = importlib.import_module("DemultiplexUsingBarcodes_New_V2")
    lib = importlib.import_module("Splitseq_fun_lib")
    with contextlib.redirect_stdout(io.StringIO()) as buf:
        lib.split_fastqR_fun(1, "R.fastq", 20000)          # writes position_learner_fastqr.fastq like CLI3:53
        V2.demultiplex_fun("F.fastq", "R.fastq", "out", 100000, 1, True, 10, True)
    assert "Barcode1 position has been extracted as 86:94" in buf.getvalue()
    assert fu.canonical_sha(open("out/MergedCells_1.fastq", "rb").read()) == digests["e1_m10"]["merged1_sha"]


def test_append_mode_like_the_reference(workdir, digests):
    V2 = importlib.import_module("DemultiplexUsingBarcodes_New_V2")
    with contextlib.redirect_stdout(io.StringIO()):
        V2.demultiplex_fun("F.fastq", "R.fastq", "out", 100000, 1, True, 10, False)
        V2.demultiplex_fun("F.fastq", "R.fastq", "out", 100000, 1, True, 10, False)   # second chunk appends (V2:367)
    recs = fu.records(open("out/MergedCells_1.fastq", "rb").read())
    assert len(recs) == 2 * digests["e1_m10"]["matched"]


def test_cli_drop_in(workdir, digests):
    argv = ["-n", "4", "-e", "1", "-m", "10", "-1", "a", "-2", "b", "-3", "c", "-f", "F.fastq", "-r", "R.fastq", "-o", "results",
            "-a", "", "-b", "100000", "-p", "True", "-l", "20000"]
    buf = io.StringIO()
    old = sys.argv
    sys.argv = ["splitseqdemultiplex_0.2.3.py"] + argv
    try:
        with contextlib.redirect_stdout(buf):
            runpy.run_path(os.path.join(PKG, "splitseqdemultiplex_0.2.3.py"), run_name="__main__")
    finally:
        sys.argv = old
    gold = digests["e1_m10"]
    assert fu.canonical_sha(open("results/MergedCells_passing.fastq", "rb").read()) == gold["passing_sha"]
    out = buf.getvalue()
    assert gold["stdout_tail"][0] in out and gold["stdout_tail"][1] in out
    assert "Barcode1 position has been extracted as 86:94" in out      # detection is on by default (CLI3:38)


RULES = [c for c in fu.load_json_gz("rule_cases.json.gz") if "error" not in c]


@pytest.mark.parametrize("case", RULES, ids=["%s-e%d" % (c["name"], c["e"]) for c in RULES])
def test_calc_demux_results_on_reference_text(case):
    """The annotated-FASTQ path alone: the reference's own MergedCells_1 text in, its MergedCells_passing out."""
    import ssd_b200
    dm = ssd_b200.Demultiplexer(WL, errors=0, min_count=case["m"])
    out, stats = dm.filter_annotated_host(case["merged1"].encode())
    assert fu.canonical(out.tobytes()).decode() == case["passing"]
    assert case["stdout_tail"][0].endswith(" %d" % stats["n_cells"])
    assert case["stdout_tail"][1].endswith(" %d" % stats["n_passing_cells"])


ADV = fu.load_json_gz("adversarial.json.gz")


@pytest.mark.parametrize("g", ADV, ids=["%s-e%d-m%d" % (g["name"], g["e"], g["m"]) for g in ADV])
def test_calc_demux_results_adversarial(g):
    import ssd_b200
    from oracle import ssd_oracle_py as orc
    r1, r2 = fu.make_pairs(g["seed"], g["n"], WL, **g["kwargs"])
    merged1 = orc.demultiplex(r1, r2, WL, g["e"], g["m"]).merged1     # == the reference's text up to record order
    assert fu.canonical_sha(merged1) == g["merged1_sha"]
    dm = ssd_b200.Demultiplexer(WL, errors=0, min_count=g["m"])
    out, stats = dm.filter_annotated_host(merged1)
    assert fu.canonical_sha(out.tobytes()) == g["passing_sha"]
    assert (stats["n_cells"], stats["n_passing_cells"], stats["n_retained"]) == (g["cells"], g["passing_cells"], g["retained"])


def test_calc_demux_results_index_error():
    import ssd_b200
    dm = ssd_b200.Demultiplexer(WL, errors=0, min_count=1)
    with pytest.raises(IndexError):
        dm.filter_annotated_host(b"@noannotation\nACGT\n+\nEEEE\n")


def test_host_pipeline_keeps_order_and_results():
    """HostPipeline: two contexts on two host threads; every batch gives what a single context gives, in order."""
    import ssd_b200
    from oracle import ssd_oracle_py as orc
    wl = fu.load_whitelist()
    batches = []
    for k in range(5):
        spec = ssd_b200.synth_spec(wl, 3000 + 500 * k, seed=900 + k, cell_pool=200)
        batches.append(ssd_b200.synth_host(spec))
    pipe = ssd_b200.HostPipeline(wl, depth=2, errors=1, min_count=2)
    got = pipe.map(batches)
    pipe.close()
    for (r1, r2), (out, st) in zip(batches, got):
        ref = orc.demultiplex(r1, r2, wl, 1, 2)
        assert out.tobytes() == ref.passing and st["n_matched"] == ref.n_matched


def test_run_files_matches_run_host(tmp_path):
    """ssd_run_files / ssd_annotated_file: file to file, same bytes as the buffer path and the oracle; append semantics
    (V2:367, 415); a missing input raises FileNotFoundError."""
    import ssd_b200
    from oracle import ssd_oracle_py as orc
    wl = fu.load_whitelist()
    spec = ssd_b200.synth_spec(wl, 150_000, seed=31, cell_pool=2000)      # ~49 MB per file: several ring chunks
    r1, r2 = ssd_b200.synth_host(spec)
    p1, p2 = tmp_path / "f.fastq", tmp_path / "r.fastq"
    p1.write_bytes(r1)
    p2.write_bytes(r2)
    ref = orc.demultiplex(r1, r2, wl, 1, 10)
    dm = ssd_b200.Demultiplexer(wl, errors=1, min_count=10)
    merged = tmp_path / "MergedCells_1.fastq"
    n, st = dm.run_files(str(p1), str(p2), str(merged), ssd_b200.EMIT_MATCHED, append=True)
    assert n == len(ref.merged1) and merged.read_bytes() == ref.merged1 and st["n_matched"] == ref.n_matched
    n, _ = dm.run_files(str(p1), str(p2), str(merged), ssd_b200.EMIT_MATCHED, append=True)       # appends like the reference
    assert merged.read_bytes() == ref.merged1 * 2
    n, _ = dm.run_files(str(p1), str(p2), str(merged), ssd_b200.EMIT_MATCHED, append=False)      # truncates on request
    assert merged.read_bytes() == ref.merged1
    passing = tmp_path / "MergedCells_passing.fastq"
    dm2 = ssd_b200.Demultiplexer(wl, errors=0, min_count=10)
    n, st2 = dm2.filter_annotated_file(str(merged), str(passing), append=True)
    assert passing.read_bytes() == ref.passing and st2["n_passing_cells"] == ref.n_passing_cells
    with pytest.raises(FileNotFoundError):
        dm.run_files(str(tmp_path / "missing.fastq"), str(p2), str(merged))
    empty = tmp_path / "empty.fastq"
    empty.write_bytes(b"")
    out = tmp_path / "out.fastq"
    n, st3 = dm.run_files(str(empty), str(empty), str(out), ssd_b200.EMIT_PASSING, append=False)
    assert n == 0 and out.read_bytes() == b"" and st3["n_records_r1"] == 0


@pytest.mark.parametrize("batch_bytes", [40000, 300000])
@pytest.mark.parametrize("e,m", [(1, 10), (2, 10)])
def test_reference_function_surface_in_batches(workdir, digests, monkeypatch, e, m, batch_bytes):
    """Files above SSD_BATCH_BYTES go through the GPU in batches of whole records (ssd_b200.filebatch: one pass for
    demultiplex_fun, two passes for calc_demux_results).  Forced here with tiny batches on the bundled files: the
    reference's golden digests and printed summary must come out unchanged, and so must the bytes of the single call."""
    V2 = importlib.import_module("DemultiplexUsingBarcodes_New_V2")
    with contextlib.redirect_stdout(io.StringIO()):
        V2.demultiplex_fun("F.fastq", "R.fastq", "out", 100000, e, True, 10, False, True)
        V2.calc_demux_results("out", True, str(m))
    whole = [open("out/MergedCells_1.fastq", "rb").read(), open("out/MergedCells_passing.fastq", "rb").read()]
    os.remove("out/MergedCells_1.fastq")
    os.remove("out/MergedCells_passing.fastq")
    monkeypatch.setenv("SSD_BATCH_BYTES", str(batch_bytes))
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        V2.demultiplex_fun("F.fastq", "R.fastq", "out", 100000, e, True, 10, False, True)
        V2.calc_demux_results("out", True, str(m))
    gold = digests["e%d_m%d" % (e, m)]
    got = [open("out/MergedCells_1.fastq", "rb").read(), open("out/MergedCells_passing.fastq", "rb").read()]
    assert fu.canonical_sha(got[0]) == gold["merged1_sha"] and fu.canonical_sha(got[1]) == gold["passing_sha"]
    assert got == whole
    assert buf.getvalue().strip().split("\n")[-2:] == gold["stdout_tail"]
    assert "The linesInInputFastq value is set to 20000" in buf.getvalue()


def test_run_files_in_batches_matches_the_oracle(tmp_path, monkeypatch):
    """run_files / filter_annotated_file with SSD_BATCH_BYTES far below the file sizes, -m decided over all batches (both
    filter modes), a surplus read-2 record in the last batch, append semantics."""
    import ssd_b200
    from oracle import ssd_oracle_py as orc
    wl = fu.load_whitelist()
    r1, r2 = ssd_b200.synth_host(ssd_b200.synth_spec(wl, 60_000, seed=32, cell_pool=800, p_dup_umi=0.4, p_n=0.01))
    p1, p2 = tmp_path / "f.fastq", tmp_path / "r.fastq"
    p1.write_bytes(r1)
    p2.write_bytes(r2)
    ref = orc.demultiplex(r1, r2, wl, 1, 10)
    monkeypatch.setenv("SSD_BATCH_BYTES", str(3_000_000))                 # 7 batches
    dm = ssd_b200.Demultiplexer(wl, errors=1, min_count=10)
    passing = tmp_path / "passing.fastq"
    n, st = dm.run_files(str(p1), str(p2), str(passing), ssd_b200.EMIT_PASSING, append=False)
    assert passing.read_bytes() == ref.passing and n == len(ref.passing)
    assert (st["n_matched"], st["n_cells"], st["n_passing_cells"], st["n_retained"], st["n_records_r1"]) == (
        ref.n_matched, ref.n_cells, ref.n_passing_cells, ref.n_retained, 60_000)
    n, _ = dm.run_files(str(p1), str(p2), str(passing), ssd_b200.EMIT_PASSING, append=True)
    assert passing.read_bytes() == ref.passing * 2
    merged = tmp_path / "merged.fastq"
    n, st = dm.run_files(str(p1), str(p2), str(merged), ssd_b200.EMIT_MATCHED, append=False)
    assert merged.read_bytes() == ref.merged1 and st["n_matched"] == ref.n_matched
    dm2 = ssd_b200.Demultiplexer(wl, errors=0, min_count=10)
    out2 = tmp_path / "passing2.fastq"
    n, st2 = dm2.filter_annotated_file(str(merged), str(out2), append=True)
    assert out2.read_bytes() == ref.passing and (st2["n_cells"], st2["n_passing_cells"]) == (ref.n_cells, ref.n_passing_cells)
    # reads mode over batches
    dm3 = ssd_b200.Demultiplexer(wl, errors=1, min_count=10, filter_mode="reads")
    out3 = tmp_path / "passing3.fastq"
    dm3.run_files(str(p1), str(p2), str(out3), ssd_b200.EMIT_PASSING, append=False)
    monkeypatch.delenv("SSD_BATCH_BYTES")
    whole3 = tmp_path / "whole3.fastq"
    dm3.run_files(str(p1), str(p2), str(whole3), ssd_b200.EMIT_PASSING, append=False)
    assert out3.read_bytes() == whole3.read_bytes()
    # a mate missing in the middle of the files is still the reference's KeyError, from whichever batch it falls in
    monkeypatch.setenv("SSD_BATCH_BYTES", str(3_000_000))
    bad = tmp_path / "bad.fastq"
    assert r2.count(b"@SYN.30001 ") == 1
    bad.write_bytes(r2.replace(b"@SYN.30001 ", b"@SYX.30001 ", 1))
    with pytest.raises(KeyError):
        dm.run_files(str(p1), str(bad), str(tmp_path / "x.fastq"), ssd_b200.EMIT_MATCHED, append=False)


def test_cli_split_version_writes_one_file_per_cell(workdir, digests):
    """-v split: the fast path's passing records as one file per cell, named like the legacy pipeline's results
    (round1-round2-round3.fastq with the full-length barcodes, demultiplex_using_barcodes.py:364), appended to like
    flushBuffers does.  The union of the files is the merged output of -v fast (the reference's golden digest)."""
    from ssd_b200 import hostlogic
    argv = ["-n", "4", "-e", "1", "-m", "10", "-1", "a", "-2", "b", "-3", "c", "-f", "F.fastq", "-r", "R.fastq", "-o", "cells",
            "-a", "", "-b", "100000", "-p", "True", "-l", "20000", "-v", "split"]
    old = sys.argv
    sys.argv = ["splitseqdemultiplex_0.2.3.py"] + argv
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            runpy.run_path(os.path.join(PKG, "splitseqdemultiplex_0.2.3.py"), run_name="__main__")
    finally:
        sys.argv = old
    gold = digests["e1_m10"]
    files = sorted(os.listdir("cells"))
    assert len(files) == gold["passing_cells"]
    assert "# of results files [%d]" % len(files) in buf.getvalue()
    names = hostlogic.load_round_names(".")
    r1 = {n[8:16]: n for n in names[0]}
    r2 = {n[0:8]: n for n in names[1]}
    r3 = {n[0:8]: n for n in names[2]}
    everything = []
    for f in files:
        recs = fu.records(open(os.path.join("cells", f), "rb").read())
        cells = {r.split(b"\n", 1)[0].split(b"_")[1].decode() for r in recs}
        assert len(cells) == 1                                             # one cell per file
        c = cells.pop()
        assert f == r1[c[0:8]] + "-" + r2[c[8:16]] + "-" + r3[c[16:24]] + ".fastq"
        everything.append(open(os.path.join("cells", f), "rb").read())
    assert fu.canonical_sha(b"".join(everything)) == gold["passing_sha"]


@pytest.mark.parametrize("variant", ["bundled", "shift3", "shift11", "crlf", "short", "no_final_newline", "leading_blanks"])
def test_device_learner_equals_host_learner(tmp_path, variant):
    """ssd_anchor_positions (find() of the three static sequences per sequence line, on the device) + the mode on the host
    give what the reference's learner gives (V2:164-199, golden learner_cases.json go through the same host half)."""
    from ssd_b200 import hostlogic
    r2 = fu.load_bundled()[1]
    if variant.startswith("shift"):
        k = int(variant[5:])
        seq = "T" * k + "ACGTACGTAC" + WL[3] + fu.LINKER_3_2 + WL[7] + fu.LINKER_2_1 + WL[50] + "ACGT" * 10
        r2 = "".join("@l%d %d/2\n%s\n+\n%s\n" % (i, i, seq, "A" * len(seq)) for i in range(300)).encode()
    if variant == "crlf":
        r2 = r2.replace(b"\n", b"\r\n")
    if variant == "short":
        r2 = b"\n".join(r2.split(b"\n")[:4 * 30]) + b"\n"       # fewer lines than the sample asks for
    if variant == "no_final_newline":
        r2 = b"\n".join(r2.split(b"\n")[:4 * 7 + 2])            # ends inside a record, on a sequence line without newline
    if variant == "leading_blanks":
        lines = r2.split(b"\n")
        lines[1::4] = [b"  " + x + b" " for x in lines[1::4]]     # strip() removes them before find()
        r2 = b"\n".join(lines)
    p = tmp_path / "r2.fastq"
    p.write_bytes(r2)
    want = hostlogic.learn_positions(hostlogic.learner_sample(str(p)))
    assert hostlogic.learn_positions_device(str(p)) == want
    if variant == "bundled":
        assert want["_found"] == (80, 56, 18)


def test_device_learner_empty_raises_value_error(tmp_path):
    from ssd_b200 import hostlogic
    p = tmp_path / "r2.fastq"
    p.write_bytes(b"@a\nACGT\n+\nEEEE\n")
    with pytest.raises(ValueError):
        hostlogic.learn_positions_device(str(p))
    p.write_bytes(b"")
    with pytest.raises(ValueError):
        hostlogic.learn_positions_device(str(p))


def test_count_newlines_on_the_device():
    import torch
    import ssd_b200
    r1 = fu.load_bundled()[0]
    dm = ssd_b200.Demultiplexer(WL)
    for data in (r1, r1[:12345], b"", b"\n" * 100, b"no newline at all"):
        t = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda() if data else torch.empty(0, dtype=torch.uint8, device="cuda")
        assert dm.count_newlines(t) == data.count(b"\n")
