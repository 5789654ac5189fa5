"""Pins the CPU oracle (oracle/ssd_oracle.c) against golden vectors produced by the unmodified
reference (tests/golden/make_golden.py): SURVEY.md Appendix A digests on the bundled FASTQs, the
Appendix B rule cases, seeded adversarial sets and the position learner."""
import hashlib
import json
import os

import pytest

import fastq_util as fu
from oracle import ssd_oracle_py as orc

WL = fu.load_whitelist()


def test_whitelist_shape():
    assert len(WL) == 96 and len(set(WL)) == 96 and all(len(w) == 8 and set(w) <= set("ACGT") for w in WL)
    assert WL[0] == "AACGTGAT" and WL[14] == "AACGCTTA"


@pytest.fixture(scope="module")
def bundled():
    return fu.load_bundled()


@pytest.fixture(scope="module")
def digests():
    with open(os.path.join(fu.GOLDEN_DIR, "bundled_digests.json")) as fh:
        return json.load(fh)


@pytest.mark.parametrize("e", [0, 1, 2, 3])
@pytest.mark.parametrize("m", [1, 10])
def test_bundled_digests(bundled, digests, e, m):
    gold = digests["e%d_m%d" % (e, m)]
    res = orc.demultiplex(bundled[0], bundled[1], WL, e, m)
    assert res.n_matched == gold["matched"]
    assert res.n_cells == gold["cells"]
    assert res.n_passing_cells == gold["passing_cells"]
    assert res.n_retained == gold["retained"]
    assert fu.canonical_sha(res.merged1) == gold["merged1_sha"]
    assert fu.canonical_sha(res.passing) == gold["passing_sha"]
    assert hashlib.sha256(fu.counts_table(res.merged1)).hexdigest() == gold["counts_sha"]
    table = b"".join(b"%s\t%d\t%d\n" % (c, v[0], v[1]) for c, v in sorted(res.cells.items()))
    assert hashlib.sha256(table).hexdigest() == gold["counts_sha"]


def test_bundled_headline_digest(bundled):
    # BASELINE.md config 1 / SURVEY Appendix A bold row, hard-coded so a regenerated fixture cannot drift
    res = orc.demultiplex(bundled[0], bundled[1], WL, 1, 10)
    assert (res.n_matched, res.n_cells, res.n_passing_cells, res.n_retained) == (3541, 430, 83, 2780)
    assert fu.canonical_sha(res.passing) == "56027fce91e05c9542e61c9dd2cfe1f0d01793a406ac39846c94fc82951dc751"


@pytest.mark.parametrize("bin_reads", [777, 1, 100000])
def test_bundled_bins(bundled, digests, bin_reads):
    res = orc.demultiplex(bundled[0], bundled[1], WL, 1, 10, bin_reads=bin_reads)
    assert fu.canonical_sha(res.passing) == digests["e1_m10"]["passing_sha"]


@pytest.mark.parametrize("workers", [1, 3, 4, 8])
def test_bundled_cli_workers(bundled, digests, workers):
    # CLI3 at -n 1,3,4 gives the same canonical output (SURVEY Appendix A note)
    res = orc.run_cli(bundled[0], bundled[1], WL, 1, 10, n_workers=workers, length_fastq=20000)
    assert fu.canonical_sha(res.passing) == digests["e1_m10"]["passing_sha"]
    assert res.n_retained == 2780


RULES = fu.load_json_gz("rule_cases.json.gz")


@pytest.mark.parametrize("case", RULES, ids=["%s-e%d" % (c["name"], c["e"]) for c in RULES])
def test_rule_cases(case):
    r1, r2 = case["r1"].encode(), case["r2"].encode()
    if "error" in case:
        with pytest.raises(orc.OracleError) as ei:
            orc.demultiplex(r1, r2, WL, case["e"], case["m"], bin_reads=case.get("bin", 100000))
        assert case["error"] == "KeyError" and ei.value.code == orc.EKEY
        assert ei.value.msg == case["error_arg"].strip("'")
        return
    res = orc.demultiplex(r1, r2, WL, case["e"], case["m"], bin_reads=case.get("bin", 100000))
    assert fu.canonical(res.merged1).decode() == case["merged1"]
    assert fu.canonical(res.passing).decode() == case["passing"]
    assert case["stdout_tail"][0].endswith(" %d" % res.n_cells)
    assert case["stdout_tail"][1].endswith(" %d" % res.n_passing_cells)


ADV = fu.load_json_gz("adversarial.json.gz")


@pytest.mark.parametrize("g", ADV, ids=["%s-e%d-m%d" % (g["name"], g["e"], g["m"]) for g in ADV])
def test_adversarial(g):
    r1, r2 = fu.make_pairs(g["seed"], g["n"], WL, **g["kwargs"])
    assert hashlib.sha256(r1 + b"|" + r2).hexdigest() == g["input_sha"], "generator drifted; rerun make_golden.py"
    res = orc.demultiplex(r1, r2, WL, g["e"], g["m"])
    assert (res.n_matched, res.n_cells, res.n_passing_cells, res.n_retained) == (g["matched"], g["cells"], g["passing_cells"], g["retained"])
    assert fu.canonical_sha(res.merged1) == g["merged1_sha"]
    assert fu.canonical_sha(res.passing) == g["passing_sha"]
    assert fu.counts_table(res.merged1).decode() == g["counts_table"]


def test_learner_cases():
    with open(os.path.join(fu.GOLDEN_DIR, "learner_cases.json")) as fh:
        cases = json.load(fh)
    for c in cases:
        shift = c["shift"]
        seq = "T" * shift + "ACGTACGTAC" + WL[3] + fu.LINKER_3_2 + WL[7] + fu.LINKER_2_1 + WL[50] + "ACGT" * 10
        recs = "".join("@l%d %d/2\n%s\n+\n%s\n" % (i, i, seq, "A" * len(seq)) for i in range(300))
        sample = "".join(l + "\n" for l in recs.split("\n")[:1001]).encode()
        p1, p2, p3 = orc.learn_positions(sample)
        printed = "\n".join(c["printed"])
        assert "UMI position has been extracted as %d:%d" % (p3 - 18, p3 - 8) in printed
        assert "Barcode3 position has been extracted as %d:%d" % (p3 - 8, p3) in printed
        assert "Barcode2 position has been extracted as %d:%d" % (p2 - 8, p2) in printed
        assert "Barcode1 position has been extracted as %d:%d" % (p1 + 6, p1 + 14) in printed


def test_learner_bundled_equals_defaults(bundled):
    sample = b"".join(l + b"\n" for l in bundled[1].split(b"\n")[:1001])
    assert orc.learn_positions(sample) == (80, 56, 18)


def test_learner_empty_raises():
    with pytest.raises(orc.OracleError):
        orc.learn_positions(b"@a\nACGT\n+\nEEEE\n")


def test_hammingtest_vectors():
    # InDevOptimizations/HammingTest.py:6-20: distances 5,3,6,1; the single <=1 hit is the 4th entry
    green = ["GGCGCATG", "ACTGAAAT", "ATGCCCGT", "ACTGAGTG"]
    seq = "A" * 10 + "ACTGACTG" + "C" * 30 + "ACTGACTG" + "C" * 30 + "ACTGACTG"
    r1 = b"@h 1/1\nACGT\n+\nEEEE\n"
    r2 = ("@h 1/2\n%s\n+\n%s\n" % (seq, "E" * len(seq))).encode()
    for e, expect in ((0, -1), (1, 3), (2, 3), (3, 1), (4, 1), (5, 0)):
        res = orc.demultiplex(r1, r2, green, e, 1)
        want = -1 if expect < 0 else (expect * 4 + expect) * 4 + expect
        assert res.rec_cell == [want]


def test_collapse_rule():
    # Collapse_RanHex_Odt.py:44-71 — RanHex[k] (idx 48+k) folds into OligoDT[k] (idx k) only when both exist
    n, pairs = 96, 48
    present = bytearray(n * n * n)
    cid = lambda a, b, c: (a * n + b) * n + c  # noqa: E731
    for cell in (cid(0, 5, 6), cid(48, 5, 6), cid(48, 5, 7), cid(49, 5, 6), cid(2, 1, 1)):
        present[cell] = 1
    remap = orc.collapse_map(bytes(present), n, pairs)
    assert remap[cid(48, 5, 6)] == cid(0, 5, 6)
    assert remap[cid(48, 5, 7)] == cid(48, 5, 7)
    assert remap[cid(49, 5, 6)] == cid(49, 5, 6)
    assert remap[cid(0, 5, 6)] == cid(0, 5, 6)
    assert sum(1 for i, r in enumerate(remap) if r != i) == 1


def test_collapse_rule_matches_the_reference_script():
    """Collapse_RanHex_Odt.py run unmodified on per-cell files of seeded random cell sets (tests/golden/make_golden.py::
    make_collapse_cases): the oracle's remap table folds exactly the cells the script folded, into the same partners."""
    import json
    import os
    with open(os.path.join(fu.GOLDEN_DIR, "collapse_cases.json")) as fh:
        cases = json.load(fh)
    n, pairs = 96, 48
    cid = lambda t: (t[0] * n + t[1]) * n + t[2]  # noqa: E731
    assert sum(len(c["merges"]) for c in cases) > 200
    for c in cases:
        present = bytearray(n * n * n)
        for cell in c["cells"]:
            present[cid(cell)] = 1
        remap = orc.collapse_map(bytes(present), n, pairs)
        want = {cid(src): cid(dst) for src, dst in c["merges"]}
        got = {i: int(r) for i, r in enumerate(remap) if r != i and present[i]}
        assert got == want, c["seed"]
        assert c["n_files_after"] == len(c["cells"]) - len(want)


@pytest.mark.parametrize("kw", [
    dict(seed=0x5EED0001),
    dict(seed=0x5EED0002, cell_pool=3000, p_n=0.01, p_tie=0.01, p_trunc=0.005, first_index=999_990),
    dict(seed=3, read_len=94, p_atqual=0.5, p_dup_umi=0.5, p_junk=0.3),
], ids=["config2", "config3-shape", "short-reads"])
def test_oracle_generator_equals_the_products_host_twin(kw):
    """oracle/ssd_synth_ref.c (what bench.py's reference arm generates its workload with, so that it never loads the product
    library) against ssd_synth_generate_host: the same bytes for the same spec."""
    import ssd_b200
    from oracle import ssd_oracle_py as orc
    wl = fu.load_whitelist()
    h1, h2 = ssd_b200.synth_host(ssd_b200.synth_spec(wl, 6000, **kw))
    a, b = orc.synth_pairs(wl, 6000, **kw)
    assert a.tobytes() == h1 and b.tobytes() == h2
    a1, b1 = orc.synth_pairs(wl, 6000, n_threads=1, **kw)
    assert a1.tobytes() == h1 and b1.tobytes() == h2


def _ref_synth_cases():
    with open(os.path.join(fu.GOLDEN_DIR, "reference_on_synthetic.json")) as fh:
        return json.load(fh)["cases"]


@pytest.mark.parametrize("case", _ref_synth_cases(), ids=lambda c: c["name"])
def test_oracle_equals_the_reference_cli_on_the_bench_workloads(case):
    """tests/golden/make_reference_on_synthetic.py ran the UNMODIFIED reference CLI (splitseqdemultiplex_0.2.3.py, 8 worker
    processes, position detection on) on the head of the synthetic bench workloads (config 2: e=1; config 3: e=2, N bases, 2/2
    ties, truncated read 2): the oracle, on the same regenerated bytes, gives the same retained records and counts."""
    import hashlib
    from oracle import ssd_oracle_py as orc
    wl = fu.load_whitelist()
    spec = {k: (int(v, 16) if k == "seed" else v) for k, v in case["spec"].items()}
    a, b = orc.synth_pairs(wl, case["pairs"], read_len=150, **spec)
    r1, r2 = a.tobytes(), b.tobytes()
    assert [hashlib.sha256(r1).hexdigest(), hashlib.sha256(r2).hexdigest()] == case["input_sha256"]
    for workers in (1, 5):  # the chunking of CLI3:52-64 does not show in the result
        ref = orc.run_cli(r1, r2, wl, case["e"], case["m"], n_workers=workers, length_fastq=r2.count(b"\n"))
        assert fu.canonical_sha(ref.passing) == case["passing_canonical_sha256"]
        assert (ref.n_cells, ref.n_passing_cells, ref.n_retained, len(ref.passing)) == (
            case["cells"], case["passing_cells"], case["retained_records"], case["passing_bytes"])
