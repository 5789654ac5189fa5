#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference
(/root/reference/python_only_tool/DemultiplexUsingBarcodes_New_V2.py, imported, never copied).

Run only in the authoring container (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Outputs (all committed):
  barcodes_8mer.txt        the 96 8-mers the reference derives as Round1 line[8:16] (V2:56-60)
  bundled_R{1,2}.fastq.gz  the reference's bundled 5 000-pair test FASTQs (data, gzip -9)
  bundled_digests.json     SURVEY.md Appendix A re-derived: counts + canonical SHA-256 for e in 0..3, m in {1,10}
  rule_cases.json.gz       SURVEY.md Appendix B: small hand-built inputs with the reference's outputs / exception
  adversarial.json.gz      seeded adversarial sets (tests/fastq_util.make_pairs) with the reference's canonical outputs
  learner_cases.json       position-learner known answers (V2:164-199)
  collapse_cases.json      which per-cell files Collapse_RanHex_Odt.py folds into which, for seeded random cell sets
  splitter_cases.json      chunk files of the reference's splitter (LIB:6-110): names, sizes, SHA-256
"""
import contextlib
import gzip
import hashlib
import importlib
import io
import json
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import fastq_util as fu  # noqa: E402

REF = "/root/reference"
REF_TOOL = os.path.join(REF, "python_only_tool")


def ref_module():
    if REF_TOOL not in sys.path:
        sys.path.insert(0, REF_TOOL)
    return importlib.import_module("DemultiplexUsingBarcodes_New_V2")


def run_reference(r1: bytes, r2: bytes, e: int, m, *, bin_size=100000, position_detection=False,
                  learner_text: bytes | None = None):
    """Run demultiplex_fun + calc_demux_results in a scratch CWD. Returns dict with raw outputs or error."""
    V2 = ref_module()
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix="ssdgold_")
    try:
        os.chdir(tmp)
        for name in ("Round1_barcodes_new5.txt", "Round2_barcodes_new4.txt", "Round3_barcodes_new4.txt"):
            shutil.copy(os.path.join(REF_TOOL, name), name)
        with open("F.fastq", "wb") as fh:
            fh.write(r1)
        with open("R.fastq", "wb") as fh:
            fh.write(r2)
        if position_detection:
            with open("position_learner_fastqr.fastq", "wb") as fh:
                fh.write(learner_text if learner_text is not None else r2)
        os.makedirs("out")
        buf = io.StringIO()
        res = {}
        try:
            with contextlib.redirect_stdout(buf):
                V2.demultiplex_fun("F.fastq", "R.fastq", "out", bin_size, e, True, 10, position_detection, False)
                V2.calc_demux_results("out", True, m)
        except Exception as exc:  # noqa: BLE001 - the exception type IS the golden answer
            res["error"] = type(exc).__name__
            res["error_arg"] = str(exc)
            return res
        with open("out/MergedCells_1.fastq", "rb") as fh:
            res["merged1"] = fh.read()
        with open("out/MergedCells_passing.fastq", "rb") as fh:
            res["passing"] = fh.read()
        res["stdout_tail"] = buf.getvalue().strip().split("\n")[-2:]
        return res
    finally:
        os.chdir(cwd)
        shutil.rmtree(tmp, ignore_errors=True)


def summarise(res):
    if "error" in res:
        return {"error": res["error"], "error_arg": res["error_arg"]}
    m1, ps = res["merged1"], res["passing"]
    table = fu.counts_table(m1)
    return {
        "matched": len(fu.records(m1)),
        "cells": table.count(b"\n"),
        "passing_cells": len({r.split(b"\n", 1)[0].split(b"_")[1] for r in fu.records(ps)}),
        "retained": len(fu.records(ps)),
        "merged1_sha": fu.canonical_sha(m1),
        "passing_sha": fu.canonical_sha(ps),
        "counts_sha": hashlib.sha256(table).hexdigest(),
        "stdout_tail": res["stdout_tail"],
    }


def make_whitelist():
    wl = [l.rstrip()[8:16] for l in open(os.path.join(REF_TOOL, "Round1_barcodes_new5.txt"))]
    assert all(l.startswith(fu.ROUND1_PREFIX) for l in open(os.path.join(REF_TOOL, "Round1_barcodes_new5.txt")))
    with open(os.path.join(HERE, "barcodes_8mer.txt"), "w") as fh:
        fh.write("\n".join(wl) + "\n")
    return wl


def make_bundled():
    out = {}
    r1 = open(os.path.join(REF_TOOL, "SRR6750041_1_smalltest.fastq"), "rb").read()
    r2 = open(os.path.join(REF_TOOL, "SRR6750041_2_smalltest.fastq"), "rb").read()
    for name, data in (("bundled_R1.fastq.gz", r1), ("bundled_R2.fastq.gz", r2)):
        with open(os.path.join(HERE, name), "wb") as raw:
            with gzip.GzipFile(filename="", mode="wb", fileobj=raw, compresslevel=9, mtime=0) as fh:
                fh.write(data)
    for e in (0, 1, 2, 3):
        for m in (1, 10):
            res = run_reference(r1, r2, e, m)
            out["e%d_m%d" % (e, m)] = summarise(res)
            print("bundled e=%d m=%d" % (e, m), out["e%d_m%d" % (e, m)])
    # learned positions must equal the defaults on this fixture
    learner = b"".join(r2.split(b"\n")[i] + b"\n" for i in range(1001))
    res = run_reference(r1, r2, 1, 10, position_detection=True, learner_text=learner)
    out["e1_m10_learned"] = summarise(res)
    # multi-bin path: 777 reads per bin
    res = run_reference(r1, r2, 1, 10, bin_size=777)
    out["e1_m10_bin777"] = summarise(res)
    with open(os.path.join(HERE, "bundled_digests.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)


def _r2(umi, bc3, bc2, bc1, tail=""):
    return umi + bc3 + fu.LINKER_3_2 + bc2 + fu.LINKER_2_1 + bc1 + tail


def _pair(idx, seq2, hdr1=None, hdr2=None, seq1="ACGTACGTACGTACGTACGT"):
    hdr1 = hdr1 or "@id%d %d/1" % (idx, idx)
    hdr2 = hdr2 or "@id%d %d/2" % (idx, idx)
    f = "%s\n%s\n+\n%s\n" % (hdr1, seq1, "E" * len(seq1))
    r = "%s\n%s\n+\n%s\n" % (hdr2, seq2, "A" * len(seq2))
    return f, r


def make_rule_cases(wl):
    """SURVEY.md Appendix B. One read pair per case unless stated; run at every e in 0..3."""
    umi = "ACGTACGTAC"
    b3, b2, b1 = wl[3], wl[7], wl[50]
    full = _r2(umi, b3, b2, b1)
    singles = {
        "01_exact": full,
        "02_tie_2_2": _r2(umi, b3, b2, "AACGCTAT"),
        "03_first_not_nearest": _r2(umi, b3, b2, "AACGTTTA"),
        "04_bc2_N": _r2(umi, b3, "N" + b2[1:], b1),
        "05_bc2_NN": _r2(umi, b3, "NN" + b2[2:], b1),
        "06_umi_N": _r2("ANNTACGTAC", b3, b2, b1),
        "07_trunc90": full[:90],
        "08_trunc86": full[:86],
        "09_trunc50": full[:50],
        "10_trunc12": full[:12],
        "11_five_bases": "ACGTA",
        "12_bc2_lower": _r2(umi, b3, b2.lower(), b1),
        "13_150bp": _r2(umi, b3, b2, b1, "ACGT" * 14),
        "20_empty_seq": "",
    }
    cases = []
    for name, seq2 in singles.items():
        f, r = _pair(0, seq2)
        cases.append({"name": name, "r1": f, "r2": r, "m": 1})
    f, r = _pair(0, full, hdr1="@id0:exact x y/1", hdr2="@id0:exact x y/2")
    cases.append({"name": "14_header_spaces", "r1": f, "r2": r, "m": 1})
    f, r = _pair(0, full, hdr1="@a 1/1", hdr2="@b 1/2")
    cases.append({"name": "16_mate_missing", "r1": f, "r2": r, "m": 1})
    # 17-19: distinct-UMI semantics of -m (V2:411); bin size 2 exercises the multi-bin path
    cellA, cellB, cellC = (wl[1], wl[2], wl[3]), (wl[4], wl[5], wl[6]), (wl[7], wl[8], wl[9])
    plan = [(cellA, "A" * 10), (cellB, "A" * 10), (cellC, "T" * 10),
            (cellA, "C" * 10), (cellB, "A" * 10), (cellC, "T" * 10),
            (cellA, "G" * 10), (cellB, "G" * 10), (cellC, "T" * 10)]
    fs, rs = [], []
    for i, (cell, u) in enumerate(plan):
        f, r = _pair(i, _r2(u, cell[2], cell[1], cell[0]))
        fs.append(f)
        rs.append(r)
    for m in (3, 2, 1):
        cases.append({"name": "17_19_distinct_umi_m%d" % m, "r1": "".join(fs), "r2": "".join(rs), "m": m, "bin": 2})
    # CRLF input, no final newline, '@' quality, '+' line repeating the id
    f = "@c1 1/1\r\nACGT\r\n+c1 1/1\r\n@EEE\r\n@c2 2/1\r\nGGCC\r\n+\r\n@@@@"
    r = "@c1 1/2\r\n%s\r\n+\r\n@%s\r\n@c2 2/2\r\n%s\r\n+\r\n@%s" % (full, "A" * (len(full) - 1), full, "A" * (len(full) - 1))
    cases.append({"name": "21_crlf_nofinalnl_atqual", "r1": f, "r2": r, "m": 1})
    # trailing incomplete record in both files (dropped), and R1 longer than R2
    f1, r1_ = _pair(1, full)
    f2, r2_ = _pair(2, full)
    cases.append({"name": "22_trailing_partial", "r1": f1 + "@id9 9/1\nACGT\n+\n", "r2": r1_ + "@id9 9/2\n" + full + "\n", "m": 1})
    cases.append({"name": "23_r1_longer", "r1": f1 + f2, "r2": r1_, "m": 1})
    cases.append({"name": "24_r2_longer_keyerror", "r1": f1, "r2": r1_ + r2_, "m": 1})
    # header forms
    for k, (h1, h2) in enumerate([("@x\t7 /1", "@x\t7/2"), ("@x 7/1  ", "@x 7/2\t"), ("@x", "@x"),
                                  ("@x 1/1/2 3/4", "@x q"), ("@x:a b \t/1", "@x:a")]):
        f, r = _pair(0, full, hdr1=h1, hdr2=h2)
        cases.append({"name": "25_header_form_%d" % k, "r1": f, "r2": r, "m": 1})
    out = []
    for c in cases:
        for e in (0, 1, 2, 3):
            res = run_reference(c["r1"].encode(), c["r2"].encode(), e, c["m"], bin_size=c.get("bin", 100000))
            entry = dict(c)
            entry["e"] = e
            if "error" in res:
                entry["error"] = res["error"]
                entry["error_arg"] = res["error_arg"]
            else:
                entry["merged1"] = fu.canonical(res["merged1"]).decode()
                entry["passing"] = fu.canonical(res["passing"]).decode()
                entry["stdout_tail"] = res["stdout_tail"]
            out.append(entry)
    with gzip.GzipFile(os.path.join(HERE, "rule_cases.json.gz"), "wb", compresslevel=9, mtime=0) as fh:
        fh.write(json.dumps(out, indent=0, sort_keys=True).encode())
    print("rule cases:", len(out))


ADVERSARIAL_SETS = [
    # name, seed, n, kwargs for fastq_util.make_pairs, list of (e, m, bin)
    ("adv_a", 101, 1500, {}, [(0, 1, 100000), (1, 2, 100000), (2, 3, 64), (3, 10, 100000)]),
    ("adv_crlf", 202, 600, {"crlf": True}, [(1, 1, 100000), (2, 2, 100000)]),
    ("adv_nofinalnl", 303, 601, {"final_newline": False, "p_short": 0.0}, [(1, 2, 100000)]),
    ("adv_dense", 404, 3000, {"ncells": 5, "p_dup_umi": 0.6, "p_junk": 0.02}, [(1, 10, 1000), (2, 25, 100000)]),
    ("adv_noisy", 505, 1200, {"p_n": 0.08, "p_short": 0.15, "p_tie": 0.2, "p_lower": 0.05}, [(0, 1, 100000), (1, 1, 100000), (2, 2, 100000), (3, 2, 100000)]),
]


def make_adversarial(wl):
    out = []
    for name, seed, n, kw, runs in ADVERSARIAL_SETS:
        r1, r2 = fu.make_pairs(seed, n, wl, **kw)
        for e, m, b in runs:
            res = run_reference(r1, r2, e, m, bin_size=b)
            assert "error" not in res, (name, res)
            s = summarise(res)
            s.update({"name": name, "seed": seed, "n": n, "kwargs": kw, "e": e, "m": m,
                      "input_sha": hashlib.sha256(r1 + b"|" + r2).hexdigest(),
                      "counts_table": fu.counts_table(res["merged1"]).decode()})
            out.append(s)
            print(name, e, m, s["matched"], s["cells"], s["passing_cells"], s["retained"])
    with gzip.GzipFile(os.path.join(HERE, "adversarial.json.gz"), "wb", compresslevel=9, mtime=0) as fh:
        fh.write(json.dumps(out, indent=0, sort_keys=True).encode())


def make_learner_cases(wl):
    """Drive the reference's learner (V2:164-199) with shifted layouts; record the offsets it prints."""
    V2 = ref_module()
    cases = []
    full = _r2("ACGTACGTAC", wl[3], wl[7], wl[50], "ACGT" * 10)
    for shift in (0, 2, 5):
        seq = "T" * shift + full
        recs = "".join("@l%d %d/2\n%s\n+\n%s\n" % (i, i, seq, "A" * len(seq)) for i in range(300))
        learner = "".join(l + "\n" for l in recs.split("\n")[:1001])
        f = "".join("@l%d %d/1\nACGTAC\n+\nEEEEEE\n" % (i, i) for i in range(300))
        cwd = os.getcwd()
        tmp = tempfile.mkdtemp(prefix="ssdlearn_")
        try:
            os.chdir(tmp)
            for name in ("Round1_barcodes_new5.txt", "Round2_barcodes_new4.txt", "Round3_barcodes_new4.txt"):
                shutil.copy(os.path.join(REF_TOOL, name), name)
            open("F.fastq", "w").write(f)
            open("R.fastq", "w").write(recs)
            open("position_learner_fastqr.fastq", "w").write(learner)
            os.makedirs("out")
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                V2.demultiplex_fun("F.fastq", "R.fastq", "out", 100000, 1, True, 10, True, False)
            lines = [l for l in buf.getvalue().split("\n") if "position" in l and "extracted as" in l]
            merged = open("out/MergedCells_1.fastq", "rb").read()
        finally:
            os.chdir(cwd)
            shutil.rmtree(tmp, ignore_errors=True)
        cases.append({"shift": shift, "printed": lines, "matched": len(fu.records(merged)),
                      "merged1_sha": fu.canonical_sha(merged)})
    with open(os.path.join(HERE, "learner_cases.json"), "w") as fh:
        json.dump(cases, fh, indent=1, sort_keys=True)
    print("learner:", cases)


def splitter_inputs():
    """Inputs of the splitter cases: (name, bytes, lengthFastq).  Shared with the tests."""
    r1, _ = fu.load_bundled()
    rec = b"@a 1/1\nACGT\n+\nEEEE\n"
    return [
        ("bundled", r1, 20000),
        ("tiny", rec * 10, 40),
        ("partial_last_record", rec * 9 + b"@b\nAC\n", 38),
        ("crlf", rec.replace(b"\n", b"\r\n") * 12, 48),
        ("trailing_blank", rec * 5 + b"@c 1/1 \nACGT\n+\nEEEE\n" + rec * 6, 48),
        ("leading_tab", rec * 5 + b"@c\n\tACGT\n+\nEEEE\n" + rec * 6, 48),
        ("no_final_newline", rec * 12 + b"@d\nAC\n+\nEE", 52),
        ("blank_lines", rec * 3 + b"\n\n\n\n" + rec * 8, 48),
    ]


def make_splitter_cases():
    """The reference's Splitseq_fun_lib.split_fastqF_fun / split_fastqR_fun (LIB:6-110), run unmodified: names, sizes and
    SHA-256 of the chunk files (and of the position-learner sample) for several inputs and core counts."""
    if REF_TOOL not in sys.path:
        sys.path.insert(0, REF_TOOL)
    lib = importlib.import_module("Splitseq_fun_lib")
    cases = []
    cwd = os.getcwd()
    for name, data, length in splitter_inputs():
        for n in (1, 3, 4, 8):
            for fn, prefix in (("split_fastqF_fun", "split_fastq_F"), ("split_fastqR_fun", "split_fastq_R")):
                tmp = tempfile.mkdtemp(prefix="ssdgold_")
                try:
                    os.chdir(tmp)
                    with open("in.fastq", "wb") as fh:
                        fh.write(data)
                    with contextlib.redirect_stdout(io.StringIO()):
                        getattr(lib, fn)(n, "in.fastq", length)
                    files = sorted(f for f in os.listdir(".") if f.startswith(prefix))
                    rec = {"input": name, "n": n, "fn": fn, "length": length,
                           "files": {f: [os.path.getsize(f), hashlib.sha256(open(f, "rb").read()).hexdigest()] for f in files}}
                    if os.path.exists("position_learner_fastqr.fastq"):
                        rec["learner_sha"] = hashlib.sha256(open("position_learner_fastqr.fastq", "rb").read()).hexdigest()
                    cases.append(rec)
                finally:
                    os.chdir(cwd)
                    shutil.rmtree(tmp, ignore_errors=True)
    with open(os.path.join(HERE, "splitter_cases.json"), "w") as fh:
        json.dump(cases, fh, indent=1, sort_keys=True)
    print("splitter cases:", len(cases))


def make_collapse_cases():
    """Collapse_RanHex_Odt.py (the reference's script, run unmodified in a scratch directory next to copies of its RanHex.txt /
    OligoDT.txt): per-cell files named round1-round2-round3.fastq (demultiplex_using_barcodes.py:364) for seeded random
    cell sets, each file holding its own name; afterwards the surviving files and their contents say which RanHex cell was
    folded into which OligoDT cell.  Cells are stored as index triples into the Round files (line i = whitelist entry i)."""
    import random
    import subprocess
    rounds = []
    for name in ("Round1_barcodes_new5.txt", "Round2_barcodes_new4.txt", "Round3_barcodes_new4.txt"):
        with open(os.path.join(REF, name)) as fh:
            rounds.append([l.strip() for l in fh if l.strip()])
    index = [{b: i for i, b in enumerate(r)} for r in rounds]
    cases = []
    for seed, n_cells, p_pair in ((1, 40, 0.5), (2, 400, 0.3), (3, 1500, 0.1), (4, 5, 1.0), (5, 300, 0.0)):
        rng = random.Random(seed)
        cells = set()
        while len(cells) < n_cells:
            i1, i2, i3 = rng.randrange(96), rng.randrange(96), rng.randrange(96)
            cells.add((i1, i2, i3))
            if rng.random() < p_pair:                       # plant the partner well: RanHex k + 48 <-> OligoDT k
                cells.add(((i1 + 48) % 96, i2, i3))
        tmp = tempfile.mkdtemp(prefix="ssdgold_")
        try:
            os.makedirs(os.path.join(tmp, "results"))
            for f in ("RanHex.txt", "OligoDT.txt"):
                shutil.copy(os.path.join(REF, f), os.path.join(tmp, f))
            for i1, i2, i3 in cells:
                name = rounds[0][i1] + "-" + rounds[1][i2] + "-" + rounds[2][i3] + ".fastq"
                with open(os.path.join(tmp, "results", name), "w") as fh:
                    fh.write(name + "\n")
            subprocess.run([sys.executable, os.path.join(REF, "Collapse_RanHex_Odt.py")], cwd=tmp, check=True,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            merges = []
            survivors = sorted(os.listdir(os.path.join(tmp, "results")))
            for f in survivors:
                lines = open(os.path.join(tmp, "results", f)).read().split()
                assert lines[0] == f
                for other in lines[1:]:
                    a, b = other[:-6].split("-"), f[:-6].split("-")
                    merges.append([[index[k][a[k]] for k in range(3)], [index[k][b[k]] for k in range(3)]])
            cases.append({"seed": seed, "cells": sorted(list(c) for c in cells), "merges": sorted(merges), "n_files_after": len(survivors)})
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    with open(os.path.join(HERE, "collapse_cases.json"), "w") as fh:
        json.dump(cases, fh, sort_keys=True)
    print("collapse cases:", [(c["seed"], len(c["cells"]), len(c["merges"])) for c in cases])


if __name__ == "__main__":
    if not os.path.isdir(REF_TOOL):
        sys.exit("the reference is not mounted; golden vectors can only be regenerated in the authoring container")
    wl = make_whitelist()
    make_bundled()
    make_rule_cases(wl)
    make_adversarial(wl)
    make_learner_cases(wl)
    make_splitter_cases()
    make_collapse_cases()
