"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the CPU oracle and the
golden vectors of the reference.  Bit-exact on the canonical (sorted-record) form, on the per-cell
counts and on the per-record cell ids."""
import hashlib
import json
import os

import numpy as np
import pytest

import fastq_util as fu

pytestmark = pytest.mark.gpu

WL = fu.load_whitelist()


@pytest.fixture(scope="module")
def ssd():
    import ssd_b200
    return ssd_b200


@pytest.fixture(scope="module")
def orc():
    from oracle import ssd_oracle_py
    return ssd_oracle_py


@pytest.fixture(scope="module")
def bundled():
    return fu.load_bundled()


@pytest.fixture(scope="module")
def digests():
    with open(os.path.join(fu.GOLDEN_DIR, "bundled_digests.json")) as fh:
        return json.load(fh)


def run_both(ssd, r1, r2, e, m, **kw):
    dm = ssd.Demultiplexer(WL, errors=e, min_count=m, **kw)
    passing, stats = dm.run_host(r1, r2, ssd.EMIT_PASSING)
    merged = np.empty(stats["bytes_matched"], dtype=np.uint8)
    import ctypes as C
    n = C.c_size_t()
    if merged.size:
        dm._check(dm.lib.ssd_emit_host(dm._ctx, ssd.EMIT_MATCHED, merged.ctypes.data_as(C.c_void_p), merged.size, C.byref(n)))
    return dm, merged[: n.value].tobytes(), passing.tobytes(), dm.stats()


def table_bytes(dm):
    return b"".join(b"%s\t%d\t%d\n" % (c.encode(), v[0], v[1]) for c, v in sorted(dm.cell_table().items()))


@pytest.mark.parametrize("e", [0, 1, 2, 3])
@pytest.mark.parametrize("m", [1, 10])
def test_bundled_digests(ssd, bundled, digests, e, m):
    gold = digests["e%d_m%d" % (e, m)]
    dm, merged, passing, stats = run_both(ssd, bundled[0], bundled[1], e, m)
    assert stats["n_records_r1"] == stats["n_records_r2"] == 5000
    assert (stats["n_matched"], stats["n_cells"], stats["n_passing_cells"], stats["n_retained"]) == (
        gold["matched"], gold["cells"], gold["passing_cells"], gold["retained"])
    assert len(merged) == stats["bytes_matched"] and len(passing) == stats["bytes_passing"]
    assert fu.canonical_sha(merged) == gold["merged1_sha"]
    assert fu.canonical_sha(passing) == gold["passing_sha"]
    assert hashlib.sha256(table_bytes(dm)).hexdigest() == gold["counts_sha"]


def test_bundled_order_is_input_order(ssd, orc, bundled):
    # the reference's order is arbitrary (set iteration); ours is the input order, like the oracle's
    _, merged, passing, _ = run_both(ssd, bundled[0], bundled[1], 1, 10)
    ref = orc.demultiplex(bundled[0], bundled[1], WL, 1, 10)
    assert merged == ref.merged1
    assert passing == ref.passing


def test_record_cells(ssd, orc, bundled):
    dm, _, _, _ = run_both(ssd, bundled[0], bundled[1], 2, 1)
    ref = orc.demultiplex(bundled[0], bundled[1], WL, 2, 1)
    assert dm.record_cells().tolist() == ref.rec_cell


RULES = fu.load_json_gz("rule_cases.json.gz")


@pytest.mark.parametrize("case", RULES, ids=["%s-e%d" % (c["name"], c["e"]) for c in RULES])
def test_rule_cases(ssd, case):
    r1, r2 = case["r1"].encode(), case["r2"].encode()
    if "error" in case:
        with pytest.raises(KeyError) as ei:
            run_both(ssd, r1, r2, case["e"], case["m"])
        assert ei.value.args[0] == case["error_arg"].strip("'")
        return
    dm, merged, passing, stats = run_both(ssd, r1, r2, case["e"], case["m"])
    assert fu.canonical(merged).decode() == case["merged1"]
    assert fu.canonical(passing).decode() == case["passing"]
    assert case["stdout_tail"][0].endswith(" %d" % stats["n_cells"])
    assert case["stdout_tail"][1].endswith(" %d" % stats["n_passing_cells"])


ADV = fu.load_json_gz("adversarial.json.gz")


@pytest.mark.parametrize("g", ADV, ids=["%s-e%d-m%d" % (g["name"], g["e"], g["m"]) for g in ADV])
def test_adversarial(ssd, g):
    r1, r2 = fu.make_pairs(g["seed"], g["n"], WL, **g["kwargs"])
    dm, merged, passing, stats = run_both(ssd, r1, r2, g["e"], g["m"])
    assert (stats["n_matched"], stats["n_cells"], stats["n_passing_cells"], stats["n_retained"]) == (
        g["matched"], g["cells"], g["passing_cells"], g["retained"])
    assert fu.canonical_sha(merged) == g["merged1_sha"]
    assert fu.canonical_sha(passing) == g["passing_sha"]
    assert table_bytes(dm).decode() == g["counts_table"]


@pytest.mark.parametrize("e,m,kw", [
    (1, 10, {}),
    (0, 3, {"p_n": 0.02}),
    (2, 5, {"p_tie": 0.05, "p_trunc": 0.02, "p_n": 0.01}),
    (3, 2, {"p_sub": 0.05, "p_junk": 0.3, "cell_pool": 300}),
])
def test_synthetic_vs_oracle(ssd, orc, e, m, kw):
    spec = ssd.synth_spec(WL, 30000, seed=0xABC0 + e, cell_pool=kw.pop("cell_pool", 2000), **kw)
    r1, r2 = ssd.synth_host(spec)
    dm, merged, passing, stats = run_both(ssd, r1, r2, e, m)
    ref = orc.demultiplex(r1, r2, WL, e, m)
    assert dm.record_cells().tolist() == ref.rec_cell
    assert merged == ref.merged1
    assert passing == ref.passing
    assert (stats["n_matched"], stats["n_cells"], stats["n_passing_cells"], stats["n_retained"]) == (
        ref.n_matched, ref.n_cells, ref.n_passing_cells, ref.n_retained)
    want = {c.decode(): (v[0], v[1]) for c, v in ref.cells.items()}
    assert dm.cell_table() == want


# ---- per-round whitelists (extension) -----------------------------------------------------------------------
def _first_hit(window: bytes, entries, e: int) -> int:
    """V2:72-73 + 302-304 restated in three lines (a second opinion next to the C oracle): zip-truncated Hamming distance,
    first entry in list order within e."""
    for i, w in enumerate(entries):
        if sum(a != b for a, b in zip(w.encode(), window)) <= e:
            return i
    return -1


PER_ROUND = (WL, WL[::-1][:50], WL[10:] + ["ACGTACGT", "TTTTTTTT"])  # three different lists of three different sizes


@pytest.mark.parametrize("e", [0, 1, 2])
def test_per_round_lists_vs_oracle(ssd, orc, e):
    """Every window against the list of its own round: record-for-record cell ids, both outputs, per-cell table and the
    cell x UMI table equal the oracle's; the cell ids also equal a three-line Python restatement of the first-hit rule."""
    l1, l2, l3 = PER_ROUND
    r1, r2 = ssd.synth_host(ssd.synth_spec(WL, 20000, seed=0x3117 + e, cell_pool=1500, p_n=0.01, p_trunc=0.01, p_sub=0.03))
    dm = ssd.Demultiplexer(l1, errors=e, min_count=3, whitelist2=l2, whitelist3=l3)
    assert dm.n_cells_total == len(l1) * len(l2) * len(l3)
    passing, stats = dm.run_host(r1, r2, ssd.EMIT_PASSING)
    ref = orc.demultiplex(r1, r2, l1, e, 3, whitelist2=l2, whitelist3=l3)
    got_cells = dm.record_cells().tolist()
    assert got_cells == ref.rec_cell
    assert passing.tobytes() == ref.passing
    assert (stats["n_matched"], stats["n_cells"], stats["n_passing_cells"], stats["n_retained"]) == (
        ref.n_matched, ref.n_cells, ref.n_passing_cells, ref.n_retained)
    assert 0 < ref.n_matched < 20000 and ref.n_passing_cells > 0
    assert dm.cell_table() == {c.decode(): (v[0], v[1]) for c, v in ref.cells.items()}
    cells, umis, reads = dm.cell_umi_table()
    got = {(dm.cell_string(int(c)).encode(), u): int(n) for c, u, n in zip(cells.tolist(), umis.tolist(), reads.tolist())}
    assert got == dict(_cell_umi_rows_from_text(ref.merged1))
    # second opinion on the first 3000 records
    seqs = r2.split(b"\n")[1::4]
    for k in range(3000):
        s = seqs[k].rstrip()
        i1, i2, i3 = _first_hit(s[86:94], l1, e), _first_hit(s[48:56], l2, e), _first_hit(s[10:18], l3, e)
        want = -1 if min(i1, i2, i3) < 0 else (i1 * len(l2) + i2) * len(l3) + i3
        assert got_cells[k] == want, k


def test_three_equal_lists_are_the_one_list(ssd, bundled, digests):
    """The reference's configuration spelled as three lists gives the reference's bytes (golden digest of V2 on the bundled
    files), through the device tables shared between the rounds."""
    dm = ssd.Demultiplexer(WL, errors=1, min_count=10, whitelist2=list(WL), whitelist3=list(WL))
    passing, stats = dm.run_host(bundled[0], bundled[1], ssd.EMIT_PASSING)
    assert fu.canonical_sha(passing.tobytes()) == digests["e1_m10"]["passing_sha"]
    assert (stats["n_matched"], stats["n_cells"], stats["n_passing_cells"], stats["n_retained"]) == (3541, 430, 83, 2780)


def test_per_round_lists_annotated_split_and_collapse(ssd, orc):
    """The other consumers of the cell id under per-round lists: the annotated-text filter (V2:389-454) parses cells back with
    the list of each round, split mode buckets by cell, the collapse map folds Round-1 index 48 + k into k."""
    l1, l2, l3 = PER_ROUND
    r1, r2 = ssd.synth_host(ssd.synth_spec(WL, 15000, seed=0x3120, cell_pool=600, p_n=0.01))
    ref = orc.demultiplex(r1, r2, l1, 1, 4, whitelist2=l2, whitelist3=l3)
    dm = ssd.Demultiplexer(l1, errors=1, min_count=4, whitelist2=l2, whitelist3=l3)
    out, st = dm.filter_annotated_host(ref.merged1)
    assert out.tobytes() == ref.passing and st["n_passing_cells"] == ref.n_passing_cells
    # split mode: the buckets partition the merged output, one cell each
    import torch
    dm.run_host(r1, r2, ssd.EMIT_PASSING)
    buf = torch.empty(len(ref.passing) + 64, dtype=torch.uint8, device="cuda")
    n, table, _ = dm.emit_split_device(buf.data_ptr(), buf.numel())
    text = buf[:n].cpu().numpy().tobytes()
    assert fu.canonical_sha(text) == fu.canonical_sha(ref.passing)
    for (_, off, _, nbytes), cid in zip(table, dm._split_ids):
        heads = text[off:off + nbytes].split(b"\n")[0::4]
        assert all(h.split(b"_")[1] == dm.cell_string(cid).encode() for h in heads if h)
    # collapse on: cells (48 + k, i2, i3) fold into (k, i2, i3) when both pass
    dmc = ssd.Demultiplexer(l1, errors=1, min_count=4, whitelist2=l2, whitelist3=l3, collapse_pairs=48)
    outc, stc = dmc.run_host(r1, r2, ssd.EMIT_PASSING)
    passing = {c for c, v in ref.cells.items() if v[1] >= 4}
    want = []
    recs = ref.passing.split(b"\n")
    for k in range(0, len(recs) - 1, 4):
        idp, cell, umi = recs[k].split(b"_")
        i1 = l1.index(cell[:8].decode())
        if i1 >= 48 and (l1[i1 - 48].encode() + cell[8:]) in passing:
            cell = l1[i1 - 48].encode() + cell[8:]
        want += [idp + b"_" + cell + b"_" + umi] + recs[k + 1:k + 4]
    assert outc.tobytes() == b"\n".join(want) + b"\n"


def _cell_umi_rows_from_text(merged: bytes, passing_cells=None):
    """(cell, UMI) -> reads, from the headers `@id_CELL_UMI` of an annotated FASTQ text (what V2:395-398 parses back)."""
    from collections import Counter
    rows = Counter()
    for h in merged.split(b"\n")[0::4]:
        if not h:
            continue
        parts = h.split(b"_")
        if passing_cells is None or parts[1] in passing_cells:
            rows[(parts[1], parts[2])] += 1
    return rows


@pytest.mark.parametrize("e,m,kw", [(1, 3, {"p_n": 0.02, "p_dup_umi": 0.3}), (2, 5, {"p_trunc": 0.02, "p_n": 0.01})])
def test_cell_umi_table_vs_oracle(ssd, orc, tmp_path, e, m, kw):
    """the per-cell x UMI read counts straight from the device (sort + run lengths) equal what the oracle's
    MergedCells_1 text says header by header -- regular UMIs, UMIs with N and truncated ones alike."""
    r1, r2 = ssd.synth_host(ssd.synth_spec(WL, 40000, seed=0xC0DE + e, cell_pool=1500, **kw))
    dm = ssd.Demultiplexer(WL, errors=e, min_count=m)
    dm.run_host(r1, r2, ssd.EMIT_PASSING)
    ref = orc.demultiplex(r1, r2, WL, e, m)
    cells, umis, reads = dm.cell_umi_table()
    got = {(dm.cell_string(int(c)).encode(), u): int(n) for c, u, n in zip(cells.tolist(), umis.tolist(), reads.tolist())}
    assert len(got) == cells.size                      # one row per pair
    assert got == dict(_cell_umi_rows_from_text(ref.merged1))
    assert int(reads.sum()) == ref.n_matched
    assert (np.diff(cells.astype(np.int64)) >= 0).all()  # sorted by cell
    # the file: passing cells only, same rows
    path = str(tmp_path / "cell_umi.tsv")
    n_rows = dm.write_cell_umi_table(path)
    lines = open(path, "rb").read().split(b"\n")
    assert lines[0] == b"cell\tumi\treads" and lines[-1] == b"" and len(lines) == n_rows + 2
    passing = {c for c, v in ref.cells.items() if v[1] >= m}
    want = _cell_umi_rows_from_text(ref.merged1, passing)
    assert {tuple(ln.split(b"\t")[:2]): int(ln.split(b"\t")[2]) for ln in lines[1:-1]} == dict(want)


def test_device_generator_matches_host_twin(ssd):
    import torch
    spec = ssd.synth_spec(WL, 5000, seed=77, first_index=999990, p_tie=0.02, p_trunc=0.02)
    h1, h2 = ssd.synth_host(spec)
    dm = ssd.Demultiplexer(WL)
    d1, d2 = ssd.synth_device(dm, spec)
    torch.cuda.synchronize()
    assert d1.cpu().numpy().tobytes() == h1
    assert d2.cpu().numpy().tobytes() == h2


def test_device_resident_path_1m(ssd, orc):
    """1 M synthetic pairs generated on the device, demultiplexed device-resident, checked record for record."""
    import torch
    spec = ssd.synth_spec(WL, 1_000_000, seed=0x5EED0001, cell_pool=20000)
    dm = ssd.Demultiplexer(WL, errors=1, min_count=10)
    d1, d2 = ssd.synth_device(dm, spec)
    out, stats = dm.run_device(d1, d2, ssd.EMIT_PASSING)
    torch.cuda.synchronize()
    r1, r2 = d1.cpu().numpy().tobytes(), d2.cpu().numpy().tobytes()
    ref = orc.demultiplex(r1, r2, WL, 1, 10)
    assert (stats["n_matched"], stats["n_cells"], stats["n_passing_cells"], stats["n_retained"]) == (
        ref.n_matched, ref.n_cells, ref.n_passing_cells, ref.n_retained)
    assert hashlib.sha256(out.cpu().numpy().tobytes()).hexdigest() == hashlib.sha256(ref.passing).hexdigest()
    assert dm.record_cells().tolist() == ref.rec_cell


@pytest.mark.parametrize("cell_pool,p_dup", [(3, 0.5), (40, 0.05)])
def test_large_cells_use_umi_bitmaps(ssd, orc, cell_pool, p_dup):
    """Cells with more than 4^10 / 64 keys count their distinct UMIs in a bitmap, smaller ones in a hash set (k_umi_sets);
    both next to each other, with many duplicates, against the oracle's len(set(...))."""
    spec = ssd.synth_spec(WL, 150000, seed=0xB17, cell_pool=cell_pool, p_dup_umi=p_dup, p_junk=0.02)
    r1, r2 = ssd.synth_host(spec)
    ref = orc.demultiplex(r1, r2, WL, 1, 10)
    dm, merged, passing, stats = run_both(ssd, r1, r2, 1, 10)
    assert max(v[0] for v in ref.cells.values()) > 20000  # at least one cell is past the bitmap threshold
    assert dm.cell_table() == {c.decode(): (v[0], v[1]) for c, v in ref.cells.items()}
    assert passing == ref.passing


def test_reads_filter_mode(ssd, orc):
    # SSD_FILTER_READS keeps cells with >= m READS (the north_star histogram filter); V2 counts distinct UMIs
    spec = ssd.synth_spec(WL, 20000, seed=5, cell_pool=300, p_dup_umi=0.6)
    r1, r2 = ssd.synth_host(spec)
    ref = orc.demultiplex(r1, r2, WL, 1, 8)
    dm, merged, passing, stats = run_both(ssd, r1, r2, 1, 8, filter_mode="reads")
    keep = {c for c, v in ref.cells.items() if v[0] >= 8}
    want = [r for r in fu.records(ref.merged1) if r.split(b"\n", 1)[0].split(b"_")[1] in keep]
    assert fu.records(passing) == want
    assert stats["n_passing_cells"] == len(keep)
    assert len(keep) != ref.n_passing_cells  # the two semantics differ on this input


def test_collapse(ssd, orc):
    # Collapse_RanHex_Odt.py:44-71: RanHex (idx 48+k) folds into OligoDT (idx k) when both cells pass
    spec = ssd.synth_spec(WL, 40000, seed=9, cell_pool=40)
    r1, r2 = ssd.synth_host(spec)
    ref = orc.demultiplex(r1, r2, WL, 1, 5)
    n = len(WL)
    idx = {w: i for i, w in enumerate(WL)}
    present = bytearray(n * n * n)
    for c, v in ref.cells.items():
        if v[2]:
            s = c.decode()
            present[(idx[s[0:8]] * n + idx[s[8:16]]) * n + idx[s[16:24]]] = 1
    remap = orc.collapse_map(bytes(present), n, 48)
    want = []
    for rec in fu.records(ref.passing):
        hdr, rest = rec.split(b"\n", 1)
        a, cell, umi = hdr.split(b"_")
        s = cell.decode()
        cid = remap[(idx[s[0:8]] * n + idx[s[8:16]]) * n + idx[s[16:24]]]
        cell2 = (WL[cid // (n * n)] + WL[(cid // n) % n] + WL[cid % n]).encode()
        want.append(a + b"_" + cell2 + b"_" + umi + b"\n" + rest)
    dm, merged, passing, stats = run_both(ssd, r1, r2, 1, 5, collapse_pairs=48)
    assert fu.records(passing) == want
    assert merged == ref.merged1  # the pre-filter output is never collapsed


@pytest.mark.parametrize("bad", [b"@a 1/1\nACGT\n+\nEEEE\nXa 2/1\nACGT\n+\nEEEE\n"])
def test_malformed_raises(ssd, bad):
    good = b"@a 1/2\nACGT\n+\nEEEE\n@b 2/2\nACGT\n+\nEEEE\n"
    with pytest.raises(ssd.MalformedFastq):
        run_both(ssd, bad, good, 1, 1)
    with pytest.raises(ssd.MalformedFastq):
        run_both(ssd, good.replace(b"/2", b"/1"), bad, 1, 1)


def test_empty_and_tiny_inputs(ssd):
    dm, merged, passing, stats = run_both(ssd, b"", b"", 1, 1)
    assert merged == b"" and passing == b"" and stats["n_records_r1"] == 0
    dm, merged, passing, stats = run_both(ssd, b"@a\nAC\n+\nEE", b"@a\nAC\n+\nEE", 1, 1)
    assert stats["n_records_r2"] == 1 and merged == b"@a_AACGTGATAACGTGATAACGTGAT_AC\nAC\n+\nEE\n"


def test_long_reads_take_the_slow_path(ssd, orc):
    # records longer than the 2 KiB halo / the emit staging buffers go through the global-memory path
    import random
    rng = random.Random(3)
    f, r = [], []
    for i in range(200):
        L = rng.choice([100, 3000, 9000, 30000])
        s1 = "".join(rng.choice("ACGT") for _ in range(L))
        s2 = "ACGTACGTAC" + WL[i % 96] + fu.LINKER_3_2 + WL[(i * 7) % 96] + fu.LINKER_2_1 + WL[(i * 3) % 96] + "".join(rng.choice("ACGT") for _ in range(L))
        f.append("@L.%d %d/1\n%s\n+\n%s\n" % (i, i, s1, "E" * L))
        r.append("@L.%d %d/2\n%s\n+\n%s\n" % (i, i, s2, "A" * len(s2)))
    r1, r2 = "".join(f).encode(), "".join(r).encode()
    ref = orc.demultiplex(r1, r2, WL, 1, 1)
    dm, merged, passing, stats = run_both(ssd, r1, r2, 1, 1)
    assert merged == ref.merged1 and passing == ref.passing and stats["n_matched"] == 200


@pytest.mark.parametrize("collapse", [0, 48])
def test_split_mode_buckets(ssd, orc, collapse):
    """Per-cell buckets (BASELINE config 5): union == merged output, every bucket holds one cell, input order inside."""
    import torch
    spec = ssd.synth_spec(WL, 50000, seed=21, cell_pool=400, p_n=0.01)
    dm = ssd.Demultiplexer(WL, errors=1, min_count=10, collapse_pairs=collapse)
    d1, d2 = ssd.synth_device(dm, spec)
    merged, stats = dm.run_device(d1, d2, ssd.EMIT_PASSING)
    merged = merged.cpu().numpy().tobytes()
    out = torch.empty(stats["bytes_passing"] + 16, dtype=torch.uint8, device="cuda")
    n, table, _ = dm.emit_split_device(out.data_ptr(), out.numel())
    split = out[:n].cpu().numpy().tobytes()
    assert n == len(merged) == stats["bytes_passing"]
    by_cell = {}
    for rec in fu.records(merged):
        by_cell.setdefault(rec.split(b"\n", 1)[0].split(b"_")[1], []).append(rec)
    assert [t[0].encode() for t in table] == [c for c in sorted(by_cell, key=lambda s: [WL.index(s[i:i + 8].decode()) for i in (0, 8, 16)])]
    assert len(table) == stats["n_passing_cells"] or collapse
    for cell, off, _, nbytes in table:
        assert fu.records(split[off:off + nbytes]) == by_cell[cell.encode()]
    if not collapse:
        r1, r2 = d1.cpu().numpy().tobytes(), d2.cpu().numpy().tobytes()
        assert merged == orc.demultiplex(r1, r2, WL, 1, 10).passing


def test_wide_offsets_path(ssd, orc, monkeypatch):
    """Files of 4 GiB and more use 64-bit record offsets; SSD_FORCE_WIDE exercises that instantiation on small inputs."""
    monkeypatch.setenv("SSD_FORCE_WIDE", "1")
    spec = ssd.synth_spec(WL, 20000, seed=31, cell_pool=300, p_n=0.01, p_trunc=0.01)
    r1, r2 = ssd.synth_host(spec)
    dm, merged, passing, stats = run_both(ssd, r1, r2, 1, 5)
    ref = orc.demultiplex(r1, r2, WL, 1, 5)
    assert merged == ref.merged1 and passing == ref.passing


def test_config3_e2_ties_n_collapse(ssd, orc):
    """BASELINE config 3 in small: e=2, N bases, planted 2/2 ties, truncated read 2, RanHex/OligoDT collapse on."""
    spec = ssd.synth_spec(WL, 200000, seed=0x5EED0002, cell_pool=3000, p_n=0.01, p_tie=0.01, p_trunc=0.005)
    r1, r2 = ssd.synth_host(spec)
    ref = orc.demultiplex(r1, r2, WL, 2, 10)
    dm, merged, passing, stats = run_both(ssd, r1, r2, 2, 10)
    assert merged == ref.merged1 and passing == ref.passing
    assert dm.record_cells().tolist() == ref.rec_cell
    # collapse on: same records, RanHex cells renamed when their OligoDT partner also passes
    dm2, merged2, passing2, stats2 = run_both(ssd, r1, r2, 2, 10, collapse_pairs=48)
    assert merged2 == ref.merged1 and len(passing2) == len(ref.passing)
    assert stats2["n_retained"] == ref.n_retained
    n = len(WL)
    idx = {w: i for i, w in enumerate(WL)}
    present = bytearray(n * n * n)
    for c, v in ref.cells.items():
        if v[2]:
            s = c.decode()
            present[(idx[s[0:8]] * n + idx[s[8:16]]) * n + idx[s[16:24]]] = 1
    remap = orc.collapse_map(bytes(present), n, 48)
    renamed = sum(1 for i, r in enumerate(remap) if r != i)
    cells_after = {rec.split(b"\n", 1)[0].split(b"_")[1] for rec in fu.records(passing2)}
    assert len(cells_after) == ref.n_passing_cells - renamed


@pytest.mark.parametrize("skew", [(16, 32, 5), (0, 16, 9), (48, 0, 15)])
def test_unaligned_output_buffer(ssd, orc, skew):
    """Output at device addresses that are not 16-byte aligned (inputs must be 16-byte aligned, include/ssd_b200.h): the
    writer moves its slices with bulk copies between aligned addresses and copies the ragged ends bytewise."""
    import torch
    spec = ssd.synth_spec(WL, 30000, seed=77, cell_pool=400)
    h1, h2 = ssd.synth_host(spec)
    ref = orc.demultiplex(h1, h2, WL, 1, 3)
    dm = ssd.Demultiplexer(WL, errors=1, min_count=3)
    s1, s2, so = skew
    b1 = torch.zeros(len(h1) + 64, dtype=torch.uint8, device="cuda")
    b2 = torch.zeros(len(h2) + 64, dtype=torch.uint8, device="cuda")
    d1, d2 = b1[s1:s1 + len(h1)], b2[s2:s2 + len(h2)]
    d1.copy_(torch.frombuffer(bytearray(h1), dtype=torch.uint8))
    d2.copy_(torch.frombuffer(bytearray(h2), dtype=torch.uint8))
    torch.cuda.synchronize()
    handle = torch.cuda.current_stream().cuda_stream
    dm.set_stream(handle)
    dm.demux_device(d1.data_ptr(), d1.numel(), d2.data_ptr(), d2.numel(), keep=(b1, b2))
    stats = dm.filter_single()
    for which, want in ((ssd.EMIT_MATCHED, ref.merged1), (ssd.EMIT_PASSING, ref.passing)):
        guard = torch.full((len(want) + 96,), 0xA5, dtype=torch.uint8, device="cuda")
        n = dm.emit_device(which, guard[so:].data_ptr(), len(want) + 16)
        torch.cuda.synchronize()
        got = guard.cpu().numpy()
        assert n == len(want) and got[so:so + n].tobytes() == want
        assert (got[:so] == 0xA5).all() and (got[so + n:] == 0xA5).all()  # nothing outside the slice is touched
    assert stats["n_matched"] == ref.n_matched


def test_group_writer_matches_tiled_writer(ssd, orc, monkeypatch):
    """SSD_EMIT_GROUPS=1 selects the eight-lanes-per-record writer (the one split mode and annotated input use)."""
    spec = ssd.synth_spec(WL, 20000, seed=11, cell_pool=300)
    h1, h2 = ssd.synth_host(spec)
    ref = orc.demultiplex(h1, h2, WL, 1, 2)
    monkeypatch.setenv("SSD_EMIT_GROUPS", "1")
    dm, merged, passing, stats = run_both(ssd, h1, h2, 1, 2)
    assert merged == ref.merged1 and passing == ref.passing


def test_scan_without_guessing_matches(ssd, orc, monkeypatch):
    """SSD_SCAN_NODEFER=1 runs the scan variant that waits for each tile's line index instead of guessing the record
    phase and parking the results (the variant the library falls back to when a guess is wrong)."""
    spec = ssd.synth_spec(WL, 50000, seed=21, cell_pool=500)
    h1, h2 = ssd.synth_host(spec)
    ref = orc.demultiplex(h1, h2, WL, 1, 2)
    monkeypatch.setenv("SSD_SCAN_NODEFER", "1")
    dm, merged, passing, stats = run_both(ssd, h1, h2, 1, 2)
    assert merged == ref.merged1 and passing == ref.passing
    assert dm.record_cells().tolist() == ref.rec_cell


def test_misleading_lines_repeat_the_scan(ssd, orc):
    """Quality lines that start with '@' followed two lines later by a read that starts with '+' cannot be told from a
    record start inside one tile: the read-2 scan's guess fails, the library repeats that scan without guessing, results
    stay exact (the '+' is part of the UMI, which is copied verbatim: V2:295)."""
    recs1, recs2 = [], []
    for i in range(6000):
        s2 = "+CGTACGTAC" + WL[i % 96] + fu.LINKER_3_2 + WL[(i * 5) % 96] + fu.LINKER_2_1 + WL[(i * 11) % 96] + "ACGT" * 10
        q2 = "@" + "F" * (len(s2) - 1)
        recs2.append("@M.%d\n%s\n+\n%s\n" % (i, s2, q2))
        recs1.append("@M.%d\n%s\n+\n%s\n" % (i, "+" + "ACGT" * 20, "@" + "E" * 80))
    r1, r2 = "".join(recs1).encode(), "".join(recs2).encode()
    ref = orc.demultiplex(r1, r2, WL, 1, 1)
    dm, merged, passing, stats = run_both(ssd, r1, r2, 1, 1)
    assert merged == ref.merged1 and passing == ref.passing and stats["n_matched"] == ref.n_matched
    assert dm.scan_repeats() >= 1


def test_short_records_repeat_once_per_context(ssd, orc):
    """Short reads (the bundled ones are 66 / 94 bp) put more records into a tile than stage B parks at once: the first
    batch repeats the read-1 scan with the variant that waits, and the context remembers the choice for later batches."""
    import random
    rng = random.Random(5)
    f, r = [], []
    for i in range(30000):
        s1 = "".join(rng.choice("ACGT") for _ in range(40))
        s2 = "".join(rng.choice("ACGT") for _ in range(10)) + WL[i % 96] + fu.LINKER_3_2 + WL[(i * 7) % 96] + fu.LINKER_2_1 + WL[(i // 5) % 96]
        f.append("@S.%d %d/1\n%s\n+\n%s\n" % (i, i, s1, "E" * 40))
        r.append("@S.%d %d/2\n%s\n+\n%s\n" % (i, i, s2, "A" * len(s2)))
    r1, r2 = "".join(f).encode(), "".join(r).encode()
    ref = orc.demultiplex(r1, r2, WL, 1, 3)
    dm, merged, passing, stats = run_both(ssd, r1, r2, 1, 3)
    assert merged == ref.merged1 and passing == ref.passing and stats["n_matched"] == ref.n_matched == 30000
    assert dm.scan_repeats() == 1
    again, _ = dm.run_host(r1, r2, ssd.EMIT_PASSING)
    assert again.tobytes() == ref.passing and dm.scan_repeats() == 1


FULL_SIZE = [
    # BASELINE.json configs[1]: 10 M pairs, e=1, merged output
    pytest.param(10_000_000, 1, 10, dict(seed=0x5EED0001), dict(), False, id="config2-10M-e1"),
    # configs[2]: 100 M pairs, e=2, N bases, planted 2/2 ties, truncated read 2, RanHex/OligoDT collapse on
    pytest.param(100_000_000, 2, 10, dict(seed=0x5EED0002, cell_pool=1_000_000, p_n=0.01, p_tie=0.01, p_trunc=0.005),
                 dict(collapse_pairs=48), False, id="config3-100M-e2-collapse"),
    # configs[4]: split-mode bucketing of 100 M pairs over the 96^3 cells
    pytest.param(100_000_000, 1, 10, dict(seed=0x5EED0004, cell_pool=1_000_000), dict(), True, id="config5-100M-split"),
]


@pytest.mark.parametrize("n,e,m,spec_kw,dm_kw,split", FULL_SIZE)
def test_full_size_properties(ssd, n, e, m, spec_kw, dm_kw, split):
    """BASELINE.json's full-size configurations are far beyond what the CPU oracle finishes in seconds, so they are checked
    through size-independent properties: the per-cell tables are additive over a split of the input into two halves (reads
    exactly, distinct UMIs as an upper bound that is tight for cells seen in only one half), the statistics follow from the
    tables, the output has exactly four lines per retained record and every record starts with '@'; in split mode the
    buckets partition the merged output (same bytes in total, one passing cell per bucket, sizes = the cells' bytes).
    Inputs of 100 M pairs are 32.7 GB per file: they also exercise the 64-bit record offsets for real."""
    import gc
    import torch
    gc.collect()
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info()
    need = (1_480 if split else 1_150) * n  # both files, the output (twice in split mode), per-record tables, slack
    if free < need:
        pytest.skip("needs %d GB of free device memory, %d GB are free" % (need >> 30, free >> 30))
    dm = ssd.Demultiplexer(WL, errors=e, min_count=m, record_bytes_hint=160, **dm_kw)
    d1, d2 = ssd.synth_device(dm, ssd.synth_spec(WL, n, **spec_kw))
    out, stats = dm.run_device(d1, d2, ssd.EMIT_PASSING)
    torch.cuda.synchronize()
    hist = dm.reads_histogram().clone().to(torch.int64)
    distinct = dm.distinct_histogram().clone().to(torch.int64)
    passing = distinct >= m
    assert stats["n_records_r1"] == n and stats["n_records_r2"] == n
    assert int(hist.sum()) == stats["n_matched"] == stats["n_regular_keys"] + stats["n_irregular_keys"]
    assert int((hist > 0).sum()) == stats["n_cells"] and int(passing.sum()) == stats["n_passing_cells"]
    assert int(hist[passing].sum()) == stats["n_retained"]
    assert bool((distinct <= hist).all()) and bool(((distinct > 0) == (hist > 0)).all())
    assert out.numel() == stats["bytes_passing"]
    # four lines per retained record, each record starting with '@' (counted in slices: the masks are as large as the output)
    n_nl, chunk = 0, 1 << 30
    for lo in range(0, out.numel(), chunk):
        nl = torch.nonzero(out[lo:lo + chunk] == 10).flatten() + lo
        first_idx = n_nl                                   # global index of this slice's first newline
        n_nl += nl.numel()
        k0 = (3 - first_idx) % 4                           # newlines number 3, 7, 11, ... end a record
        ends = nl[k0::4]
        ends = ends[ends + 1 < out.numel()]
        assert bool((out[ends + 1] == ord("@")).all())
        del nl, ends
    assert n_nl == 4 * stats["n_retained"] and (out.numel() == 0 or int(out[0]) == ord("@"))
    if split:
        sp_out = torch.empty(stats["bytes_passing"] + 16, dtype=torch.uint8, device="cuda")
        nbytes, table, _ = dm.emit_split_device(sp_out.data_ptr(), sp_out.numel())
        assert nbytes == stats["bytes_passing"] and len(table) == stats["n_passing_cells"]
        cells = [t[0] for t in table]
        assert len(set(cells)) == len(cells)
        offs = [t[1] for t in table]
        assert offs == sorted(offs) and offs[0] == 0 and offs[-1] + table[-1][3] == nbytes
        assert all(offs[k] + table[k][3] == offs[k + 1] for k in range(len(table) - 1))
        # the same multiset of bytes as the merged output, and a sample of buckets holds only its own cell
        hm, hs = torch.zeros(256, dtype=torch.int64, device="cuda"), torch.zeros(256, dtype=torch.int64, device="cuda")
        for lo in range(0, nbytes, chunk):
            hm += torch.bincount(out[lo:lo + chunk].to(torch.int32), minlength=256)
            hs += torch.bincount(sp_out[lo:min(lo + chunk, nbytes)].to(torch.int32), minlength=256)
        assert bool((hm == hs).all())
        step = max(1, len(table) // 50)
        for cell, off, _, nb in table[::step]:
            recs = fu.records(sp_out[off:off + nb].cpu().numpy().tobytes())
            assert recs and all(r.split(b"\n", 1)[0].split(b"_")[1] == cell.encode() for r in recs)
        del sp_out
    del out
    # additivity over two halves of the same input (generated separately: the generator is counter based)
    halves_hist = torch.zeros_like(hist)
    halves_distinct = torch.zeros_like(distinct)
    del d1, d2
    torch.cuda.empty_cache()
    for first in (1, n // 2 + 1):
        h1, h2 = ssd.synth_device(dm, ssd.synth_spec(WL, n // 2, first_index=first, **spec_kw))
        dm.run_device(h1, h2, ssd.EMIT_PASSING)
        torch.cuda.synchronize()
        hh = dm.reads_histogram().to(torch.int64)
        dd = dm.distinct_histogram().to(torch.int64)
        only_here = (hh > 0) & (hist == hh)                                        # cells whose reads all sit in this half
        assert bool((dd[only_here] == distinct[only_here]).all())
        halves_hist += hh
        halves_distinct += dd
        del h1, h2
    assert bool((halves_hist == hist).all())
    assert bool((halves_distinct >= distinct).all())
    dm.close()
    torch.cuda.empty_cache()


FULL_SIZE_ORACLE = [
    # BASELINE.json configs[1], the bench workload itself: 10 M pairs, seed 0x5EED0001, e=1, -m 10
    pytest.param(10_000_000, 1, 1, 10, dict(seed=0x5EED0001), id="config2-10M-e1"),
    # configs[2]'s shape (e=2, N bases, planted 2/2 ties, truncated read 2): a 10 M slice from the middle of the 100 M set
    pytest.param(10_000_000, 45_000_001, 2, 10, dict(seed=0x5EED0002, cell_pool=1_000_000, p_n=0.01, p_tie=0.01, p_trunc=0.005),
                 id="config3-slice-10M-e2"),
]


@pytest.mark.parametrize("n,first,e,m,spec_kw", FULL_SIZE_ORACLE)
def test_full_size_against_oracle(ssd, orc, n, first, e, m, spec_kw):
    """SURVEY 8c "full at 10 M": the bench workload (and a 10 M slice of config 3) bit for bit against the oracle -- CLI3's
    structure (split, one demultiplex_fun per chunk on every host core, calc_demux_results) on the bytes the device
    generated: both outputs, the statistics and the whole per-cell (reads, distinct UMIs) table."""
    import gc
    import torch
    gc.collect()
    torch.cuda.empty_cache()
    dm = ssd.Demultiplexer(WL, errors=e, min_count=m, record_bytes_hint=160)
    d1, d2 = ssd.synth_device(dm, ssd.synth_spec(WL, n, first_index=first, **spec_kw))
    out, stats = dm.run_device(d1, d2, ssd.EMIT_PASSING)
    torch.cuda.synchronize()
    r1, r2 = d1.cpu().numpy(), d2.cpu().numpy()
    g1, g2 = orc.synth_pairs(WL, n, first_index=first, **spec_kw)  # the oracle-side generator makes the same bytes
    assert r1.size == g1.size and r2.size == g2.size and bool((r1 == g1).all()) and bool((r2 == g2).all())
    del g1, g2
    ref = orc.run_cli(r1.tobytes(), r2.tobytes(), WL, e, m, n_workers=os.cpu_count() or 1, length_fastq=4 * n)
    assert (stats["n_records_r1"], stats["n_matched"], stats["n_cells"], stats["n_passing_cells"], stats["n_retained"]) == (
        n, ref.n_matched, ref.n_cells, ref.n_passing_cells, ref.n_retained)
    got = out.cpu().numpy().tobytes()
    assert len(got) == len(ref.passing) == stats["bytes_passing"]
    assert hashlib.sha256(got).hexdigest() == hashlib.sha256(ref.passing).hexdigest()
    del got
    merged = torch.empty(stats["bytes_matched"] + 16, dtype=torch.uint8, device="cuda")
    k = dm.emit_device(ssd.EMIT_MATCHED, merged.data_ptr(), merged.numel())
    assert k == len(ref.merged1)
    assert hashlib.sha256(merged[:k].cpu().numpy().tobytes()).hexdigest() == hashlib.sha256(ref.merged1).hexdigest()
    hist, distinct = dm.reads_histogram().cpu().numpy(), dm.distinct_histogram().cpu().numpy()
    idx = {w: i for i, w in enumerate(WL)}
    nw = len(WL)
    want_h, want_d = np.zeros_like(hist), np.zeros_like(distinct)
    for c, v in ref.cells.items():
        s = c.decode()
        cid = (idx[s[0:8]] * nw + idx[s[8:16]]) * nw + idx[s[16:24]]
        want_h[cid], want_d[cid] = v[0], v[1]
    assert bool((hist == want_h).all()) and bool((distinct == want_d).all())
    dm.close()
    del d1, d2, out, merged
    torch.cuda.empty_cache()


@pytest.mark.parametrize("filter_mode", ["distinct_umi", "reads"])
def test_batches_equal_one_call(ssd, orc, filter_mode):
    """Inputs cut into batches (two passes: global tables first, then the writer) give the bytes of one call on the whole
    input, which in turn are the oracle's."""
    from ssd_b200.distributed import demux_batches
    sizes = [20000, 1, 35000, 4999]
    kw = dict(seed=0xBA7C, cell_pool=1500, p_n=0.01, p_dup_umi=0.4)
    dm = ssd.Demultiplexer(WL, errors=1, min_count=10, filter_mode=filter_mode)

    def batches():
        first = 1
        for n in sizes:
            yield ssd.synth_device(dm, ssd.synth_spec(WL, n, first_index=first, **kw))
            first += n

    parts = []
    total = demux_batches(dm, batches, sink=lambda out, st: parts.append(out.cpu().numpy().tobytes()))
    r1, r2 = ssd.synth_host(ssd.synth_spec(WL, sum(sizes), **kw))
    whole, stats = ssd.Demultiplexer(WL, errors=1, min_count=10, filter_mode=filter_mode).run_host(r1, r2, ssd.EMIT_PASSING)
    assert b"".join(parts) == whole.tobytes()
    assert [total[k] for k in ("n_cells", "n_passing_cells", "n_retained", "n_matched")] == [stats[k] for k in ("n_cells", "n_passing_cells", "n_retained", "n_matched")]
    if filter_mode == "distinct_umi":
        assert whole.tobytes() == orc.demultiplex(r1, r2, WL, 1, 10).passing


@pytest.mark.parametrize("seed", range(24))
def test_fuzz_against_oracle(ssd, orc, seed):
    """Seeded random configurations of the adversarial generator (sizes, line endings, header forms, N / tie / junk /
    truncation rates, e, m), CUDA path against the CPU oracle: both outputs byte for byte, per-record cells, counts."""
    _fuzz(ssd, orc, seed)


def _fuzz(ssd, orc, seed):
    import random
    rng = random.Random(1000 + seed)
    n = rng.choice([1, 7, 300, 2500, 12000])
    kw = dict(ncells=rng.choice([1, 5, 40, 400]), crlf=rng.random() < 0.3, r1_len=rng.choice([(1, 30), (20, 80), (150, 150), (400, 900)]),
              weird_headers=rng.random() < 0.7, final_newline=rng.random() < 0.7, p_short=rng.choice([0.0, 0.05, 0.3]),
              p_n=rng.choice([0.0, 0.02, 0.15]), p_junk=rng.choice([0.0, 0.1, 0.5]), p_tie=rng.choice([0.0, 0.03, 0.2]),
              p_lower=rng.choice([0.0, 0.01]), p_atqual=rng.choice([0.0, 0.1, 0.6]), p_dup_umi=rng.choice([0.0, 0.2, 0.8]))
    e, m = rng.choice([0, 1, 1, 2, 3]), rng.choice([1, 2, 5, 10])
    r1, r2 = fu.make_pairs(5000 + seed, n, WL, **kw)
    ref = orc.demultiplex(r1, r2, WL, e, m)
    dm, merged, passing, stats = run_both(ssd, r1, r2, e, m)
    assert merged == ref.merged1 and passing == ref.passing
    assert (stats["n_matched"], stats["n_cells"], stats["n_passing_cells"], stats["n_retained"]) == (
        ref.n_matched, ref.n_cells, ref.n_passing_cells, ref.n_retained)
    assert dm.record_cells().tolist() == ref.rec_cell


def test_extract_style_headers_equal_the_reference_script(ssd, orc):
    """SSD_HEADER_EXTRACT against the UNMODIFIED Extract_BC_UMI.py (tests/golden/make_golden_extract.py ran it on these very
    reads): `id_cell_umi rest-of-name`, then read, "+", quality.  The script names the cell after its input file and finds the
    UMI by its own rule, so the golden record is compared with the cell field swapped for the fast path's cell, and the UMI
    wherever the script's position finder lands on the default offset; the outputs, the filter and the tables otherwise equal
    the default header style's."""
    import gzip
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_extract", os.path.join(fu.GOLDEN_DIR, "make_golden_extract.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    with gzip.open(os.path.join(fu.GOLDEN_DIR, "extract_bc_umi_case.json.gz")) as fh:
        case = json.loads(fh.read())
    r1, r2 = gen.inputs()
    gold = case["output"].encode().split(b"\n")
    gold = {gold[k].split(b"_")[0]: gold[k:k + 4] for k in range(0, len(gold) - 1, 4)}
    assert len(gold) == case["n"]
    dm = ssd.Demultiplexer(WL, errors=1, min_count=1, header_style="extract")
    out, stats = dm.run_host(r1, r2, ssd.EMIT_MATCHED)
    ref = orc.demultiplex(r1, r2, WL, 1, 1)  # the default style: same records, same cells, same UMIs
    lines, fast = out.tobytes().split(b"\n"), ref.merged1.split(b"\n")
    assert stats["n_matched"] == ref.n_matched > 150 and len(lines) == len(fast)
    same_umi = 0
    for k in range(0, len(lines) - 1, 4):
        head, rest = lines[k].split(b" ", 1)
        ident, cell, umi = head.split(b"_")
        g = gold[ident]
        g_ident, g_cell, g_umi_rest = g[0].split(b"_")
        g_umi, g_rest = g_umi_rest.split(b" ", 1)
        assert g_cell == case["cell"].encode() and rest == g_rest and lines[k + 1:k + 4] == g[1:4]
        assert (cell, umi) == tuple(fast[k].split(b"_")[1:3]) and lines[k + 1:k + 4] == fast[k + 1:k + 4]
        same_umi += umi == g_umi
    assert same_umi >= 0.9 * stats["n_matched"]
    # the filter does not look at headers: -m decides as with the default style
    dm10 = ssd.Demultiplexer(WL, errors=1, min_count=3, header_style="extract")
    out10, st10 = dm10.run_host(r1, r2, ssd.EMIT_PASSING)
    ref10 = orc.demultiplex(r1, r2, WL, 1, 3)
    assert (st10["n_passing_cells"], st10["n_retained"]) == (ref10.n_passing_cells, ref10.n_retained)
    assert out10.tobytes().split(b"\n")[1::4] == ref10.passing.split(b"\n")[1::4]  # the same records, in the same order
    # a name of one token: ReadName_parts[1] is an IndexError in the script
    with pytest.raises(IndexError):
        ssd.Demultiplexer(WL, errors=1, min_count=1, header_style="extract").run_host(r1.replace(b"@SRR6750041.5 5/1", b"@SRR6750041.5"), r2, ssd.EMIT_MATCHED)


# ---- the read-2 scan variant for short sequences (k_scan<.., SHORT>) ------------------------------------------------
# A context starts with the plain read-2 scan and takes the variant that keeps short sequences and windows with several non-ACGT
# bytes on the fast path once a batch sent more than 1 record in 1 000 down the exact path; SSD_SCAN_SHORT=1 / 0 pins the choice.
@pytest.mark.parametrize("case", RULES, ids=["%s-e%d" % (c["name"], c["e"]) for c in RULES])
def test_rule_cases_short_variant(ssd, case, monkeypatch):
    monkeypatch.setenv("SSD_SCAN_SHORT", "1")
    test_rule_cases(ssd, case)


@pytest.mark.parametrize("g", ADV, ids=["%s-e%d-m%d" % (g["name"], g["e"], g["m"]) for g in ADV])
def test_adversarial_short_variant(ssd, g, monkeypatch):
    monkeypatch.setenv("SSD_SCAN_SHORT", "1")
    test_adversarial(ssd, g)


@pytest.mark.parametrize("seed", range(24))
def test_fuzz_against_oracle_short_variant(ssd, orc, seed, monkeypatch):
    monkeypatch.setenv("SSD_SCAN_SHORT", "1")
    _fuzz(ssd, orc, seed)


@pytest.mark.parametrize("e,m,kw", [
    (2, 5, {"p_tie": 0.05, "p_trunc": 0.02, "p_n": 0.03}),
    (1, 3, {"p_trunc": 0.3, "p_n": 0.01}),
    (0, 1, {"p_trunc": 0.05, "p_n": 0.02}),
    (3, 2, {"p_sub": 0.05, "p_n": 0.05, "p_trunc": 0.01, "cell_pool": 300}),
])
def test_short_variant_vs_oracle(ssd, orc, monkeypatch, e, m, kw):
    """Truncated read 2 (V2:296-301 slices past the end; hamming() zips to the shorter operand, V2:72-73), windows with one, two and
    more N: the variant's warp-wide whitelist search against the oracle, record for record; and it really keeps them off the
    exact path (the plain scan sends every truncated read there)."""
    kw = dict(kw)
    spec = ssd.synth_spec(WL, 30000, seed=0x5407 + e, cell_pool=kw.pop("cell_pool", 2000), **kw)
    r1, r2 = ssd.synth_host(spec)
    ref = orc.demultiplex(r1, r2, WL, e, m)
    exact = {}
    for pin in ("0", "1"):
        monkeypatch.setenv("SSD_SCAN_SHORT", pin)
        dm, merged, passing, stats = run_both(ssd, r1, r2, e, m)
        assert dm.record_cells().tolist() == ref.rec_cell
        assert merged == ref.merged1 and passing == ref.passing
        assert (stats["n_matched"], stats["n_cells"], stats["n_passing_cells"], stats["n_retained"]) == (
            ref.n_matched, ref.n_cells, ref.n_passing_cells, ref.n_retained)
        assert dm.cell_table() == {c.decode(): (v[0], v[1]) for c, v in ref.cells.items()}
        short, exact[pin] = dm.scan_variant()
        assert short == (pin == "1")
    assert exact["0"] >= 0.005 * 30000 and exact["1"] * 20 <= exact["0"]


def test_scan_variant_follows_the_data(ssd, orc, monkeypatch):
    """No pin: the first batch with truncated reads runs the plain scan and turns the variant on, the next batch of the context
    runs it -- the same bytes; a context that only sees whole reads never switches."""
    monkeypatch.delenv("SSD_SCAN_SHORT", raising=False)
    r1, r2 = ssd.synth_host(ssd.synth_spec(WL, 40000, seed=0x5408, cell_pool=1500, p_trunc=0.02, p_n=0.01))
    ref = orc.demultiplex(r1, r2, WL, 2, 4)
    dm = ssd.Demultiplexer(WL, errors=2, min_count=4)
    assert dm.scan_variant() == (False, 0)
    out1, st1 = dm.run_host(r1, r2, ssd.EMIT_PASSING)
    out1 = out1.tobytes()
    short, exact1 = dm.scan_variant()
    assert short and exact1 >= 0.01 * 40000
    out2, st2 = dm.run_host(r1, r2, ssd.EMIT_PASSING)
    short, exact2 = dm.scan_variant()
    assert short and exact2 * 20 <= exact1
    assert out1 == out2.tobytes() == ref.passing and st1 == st2
    assert dm.record_cells().tolist() == ref.rec_cell
    dm.close()
    c1, c2 = ssd.synth_host(ssd.synth_spec(WL, 40000, seed=0x5409, cell_pool=1500))
    dm = ssd.Demultiplexer(WL, errors=1, min_count=4)
    dm.run_host(c1, c2, ssd.EMIT_PASSING)
    dm.run_host(c1, c2, ssd.EMIT_PASSING)
    assert dm.scan_variant()[0] is False
    dm.close()


# ---- the reference CLI itself on the head of the bench workloads (tests/golden/make_reference_on_synthetic.py) -----------------------
with open(os.path.join(fu.GOLDEN_DIR, "reference_on_synthetic.json")) as _fh:
    REF_SYNTH = json.load(_fh)["cases"]


@pytest.mark.parametrize("case", REF_SYNTH, ids=[c["name"] for c in REF_SYNTH])
def test_reference_cli_on_the_bench_workloads(ssd, orc, case):
    """The CUDA path against what the UNMODIFIED reference CLI wrote for the first 200 k pairs of config 2 and the first 100 k of
    config 3's shape (e=2, N bases, 2/2 ties, truncated read 2): same retained records (canonical SHA-256), same counts -- no
    oracle in between."""
    spec = {k: (int(v, 16) if k == "seed" else v) for k, v in case["spec"].items()}
    a, b = orc.synth_pairs(WL, case["pairs"], read_len=150, **spec)
    r1, r2 = a.tobytes(), b.tobytes()
    assert [hashlib.sha256(r1).hexdigest(), hashlib.sha256(r2).hexdigest()] == case["input_sha256"]
    dm, merged, passing, stats = run_both(ssd, r1, r2, case["e"], case["m"])
    assert fu.canonical_sha(passing) == case["passing_canonical_sha256"]
    assert (stats["n_cells"], stats["n_passing_cells"], stats["n_retained"], len(passing)) == (
        case["cells"], case["passing_cells"], case["retained_records"], case["passing_bytes"])
