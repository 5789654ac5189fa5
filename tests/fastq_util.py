"""Shared test helpers: canonical FASTQ form, fixture loading, and a
seeded adversarial read-pair generator used both to make the golden vectors and to drive the
oracle-vs-CUDA parity tests.

Canonical form of a FASTQ output: split on "\\n", drop the final empty element, group 4 lines per
record, sort the records bytewise, join with "\\n", add a trailing "\\n".  The reference writes records
in `set` iteration order (DemultiplexUsingBarcodes_New_V2.py:368), so only this form is comparable.
"""
from __future__ import annotations

import gzip
import hashlib
import json
import os
import random
from typing import Dict, List, Tuple

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Constant 8-base prefix of every line of the reference's Round1 whitelist file; the reference keeps
# characters [8:16] (DemultiplexUsingBarcodes_New_V2.py:60).
ROUND1_PREFIX = "AGCATTCG"
LINKER_3_2 = "GTGGCCGCTGTTTCGCATCGGCGTACGACT"   # between BC3 and BC2 (modal linker of the fixture)
LINKER_2_1 = "ATCCACGTGCTTGAGAGGCCAGAGCATTCG"   # between BC2 and BC1


def records(data: bytes) -> List[bytes]:
    lines = data.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    return [b"\n".join(lines[i:i + 4]) for i in range(0, len(lines), 4)]


def canonical(data: bytes) -> bytes:
    recs = sorted(records(data))
    if not recs:
        return b""
    return b"\n".join(recs) + b"\n"


def canonical_sha(data: bytes) -> str:
    return hashlib.sha256(canonical(data)).hexdigest()


def counts_table(merged1: bytes) -> bytes:
    """`<cell>\\t<reads>\\t<distinct UMIs>\\n` per cell of a pre-filter output, sorted by cell."""
    cells: Dict[bytes, List[bytes]] = {}
    for rec in records(merged1):
        hdr = rec.split(b"\n", 1)[0]
        parts = hdr.split(b"_")
        cells.setdefault(parts[1], []).append(parts[2])
    out = []
    for cell in sorted(cells):
        out.append(b"%s\t%d\t%d\n" % (cell, len(cells[cell]), len(set(cells[cell]))))
    return b"".join(out)


def load_whitelist() -> List[str]:
    with open(os.path.join(GOLDEN_DIR, "barcodes_8mer.txt")) as fh:
        return [l.strip() for l in fh if l.strip()]


def write_whitelist_files(dirpath: str, wl: List[str] | None = None) -> None:
    """Materialise the three files the reference opens relative to CWD (V2:56-66)."""
    wl = wl or load_whitelist()
    with open(os.path.join(dirpath, "Round1_barcodes_new5.txt"), "w") as fh:
        for w in wl:
            fh.write(ROUND1_PREFIX + w + "\n")
    with open(os.path.join(dirpath, "Round2_barcodes_new4.txt"), "w") as fh:
        for w in wl:
            fh.write(w + "ATCCA\n")
    with open(os.path.join(dirpath, "Round3_barcodes_new4.txt"), "w") as fh:
        for w in wl:
            fh.write(w + "GTGGCC\n")


def load_bundled() -> Tuple[bytes, bytes]:
    with gzip.open(os.path.join(GOLDEN_DIR, "bundled_R1.fastq.gz"), "rb") as fh:
        r1 = fh.read()
    with gzip.open(os.path.join(GOLDEN_DIR, "bundled_R2.fastq.gz"), "rb") as fh:
        r2 = fh.read()
    return r1, r2


def load_json_gz(name: str):
    with gzip.open(os.path.join(GOLDEN_DIR, name), "rt") as fh:
        return json.load(fh)


# ----------------------------------------------------------------------------------------------
# adversarial generator
# ----------------------------------------------------------------------------------------------
QUAL_ALPHABET = "EA6/<"


def _mutate(rng: random.Random, s: str, nsub: int) -> str:
    s = list(s)
    for pos in rng.sample(range(len(s)), min(nsub, len(s))):
        s[pos] = rng.choice([c for c in "ACGT" if c != s[pos]])
    return "".join(s)


def _tie_barcode(rng: random.Random, wl: List[str]) -> str:
    """An 8-mer at distance 2 from two whitelist entries that are 4 apart (exact 2/2 tie)."""
    for _ in range(200):
        a, b = rng.sample(wl, 2)
        diff = [i for i in range(8) if a[i] != b[i]]
        if len(diff) == 4:
            s = list(a)
            for i in rng.sample(diff, 2):
                s[i] = b[i]
            return "".join(s)
    return rng.choice(wl)


def make_pairs(seed: int, n: int, wl: List[str], *, ncells: int = 12, crlf: bool = False,
               id_prefix: str = "T", r1_len: Tuple[int, int] = (20, 80),
               weird_headers: bool = True, final_newline: bool = True,
               p_short: float = 0.05, p_n: float = 0.02, p_junk: float = 0.1,
               p_tie: float = 0.03, p_lower: float = 0.01, p_atqual: float = 0.1,
               p_dup_umi: float = 0.2) -> Tuple[bytes, bytes]:
    """Return (R1, R2) FASTQ bytes with planted barcodes, substitutions, N bases, 2/2 ties, truncated
    read 2, '@'-leading quality lines and awkward (but in-contract: unique, no '_') headers."""
    rng = random.Random(seed)
    cells = [(rng.randrange(len(wl)), rng.randrange(len(wl)), rng.randrange(len(wl))) for _ in range(ncells)]
    cell_umis: Dict[int, List[str]] = {}
    nl = "\r\n" if crlf else "\n"
    f_out: List[str] = []
    r_out: List[str] = []
    for i in range(1, n + 1):
        ci = min(int(rng.paretovariate(1.0)) - 1, ncells - 1) if rng.random() < 0.8 else rng.randrange(ncells)
        i1, i2, i3 = cells[ci]
        bc1, bc2, bc3 = wl[i1], wl[i2], wl[i3]
        if rng.random() < p_junk:
            bc1 = "".join(rng.choice("ACGT") for _ in range(8))
        for which in range(3):
            r = rng.random()
            if r < 0.25:
                nsub = rng.choice([1, 1, 1, 2, 2, 3])
                if which == 0:
                    bc1 = _mutate(rng, bc1, nsub)
                elif which == 1:
                    bc2 = _mutate(rng, bc2, nsub)
                else:
                    bc3 = _mutate(rng, bc3, nsub)
        if rng.random() < p_tie:
            bc2 = _tie_barcode(rng, wl)
        umis = cell_umis.setdefault(ci, [])
        if umis and rng.random() < p_dup_umi:
            umi = rng.choice(umis)
        else:
            umi = "".join(rng.choice("ACGT") for _ in range(10))
            umis.append(umi)
        tail = "".join(rng.choice("ACGT") for _ in range(rng.choice([0, 0, 6, 56])))
        seq2 = umi + bc3 + LINKER_3_2 + bc2 + LINKER_2_1 + bc1 + tail
        seq2 = "".join(("N" if rng.random() < p_n else c) for c in seq2)
        if rng.random() < p_lower:
            a = rng.choice([10, 48, 86])
            seq2 = seq2[:a] + seq2[a:a + 8].lower() + seq2[a + 8:]
        if rng.random() < p_short:
            seq2 = seq2[:rng.randrange(0, 94)]
        q2 = "".join(rng.choice(QUAL_ALPHABET) for _ in seq2)
        if q2 and rng.random() < p_atqual:
            q2 = "@" + q2[1:]
        l1 = rng.randint(*r1_len)
        seq1 = "".join(rng.choice("ACGTN" if rng.random() < 0.1 else "ACGT") for _ in range(l1))
        q1 = "".join(rng.choice(QUAL_ALPHABET) for _ in range(l1))
        if q1 and rng.random() < p_atqual:
            q1 = "@" + q1[1:]
        tok = "@%s.%d" % (id_prefix, i)
        style = rng.randrange(8) if weird_headers else 0
        if style == 1:
            h1, h2 = "%s  %d extra words/1" % (tok, i), "%s  %d extra words/2" % (tok, i)
        elif style == 2:
            h1, h2 = "%s" % tok, "%s" % tok                         # no '/', no description
        elif style == 3:
            h1, h2 = "%s %d/1/x y/z" % (tok, i), "%s %d/2" % (tok, i)  # several '/'
        elif style == 4:
            h1, h2 = "%s\t%d /1" % (tok, i), "%s\t%d/2" % (tok, i)   # tab-separated token, blank before '/'
        elif style == 5:
            h1, h2 = "%s %d/1   " % (tok, i), "%s %d/2 \t" % (tok, i)  # trailing blanks
        elif style == 6:
            h1, h2 = "%s:a b:c \t/1" % tok, "%s:a 2" % tok          # token differs after first blank only
        else:
            h1, h2 = "%s %d/1" % (tok, i), "%s %d/2" % (tok, i)
        plus1 = "+" if rng.random() < 0.9 else "+" + h1[1:]
        f_out.append(h1 + nl + seq1 + nl + plus1 + nl + q1 + nl)
        r_out.append(h2 + nl + seq2 + nl + "+" + nl + q2 + nl)
    f = "".join(f_out)
    r = "".join(r_out)
    if not final_newline:
        f = f[:-len(nl)]
        r = r[:-len(nl)]
    return f.encode("ascii"), r.encode("ascii")
