"""CPU test of csrc/ssd_match.cuh, the host/device header with the packed first-hit rule the read-2 scan applies to windows its
lookup tables do not answer (short read 2: lineRead[s:e] gives what is there and hamming() zips to the shorter operand,
V2:72-73, 296-304; several bytes outside ACGT).  The header is compiled with g++ into a small harness that runs the rule the way
the warp does (32 lanes voting on 32 whitelist entries at a time) and is compared with the reference's formula in Python."""
import ctypes as C
import os
import random
import subprocess

import pytest

import fastq_util as fu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "split-seq_demultiplexing_b200", "csrc")
WL = fu.load_whitelist()


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("match") / "libmatch_logic.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", "-Wno-unknown-pragmas", "-I", CSRC, "-o", so,
                    os.path.join(ROOT, "tests", "match_logic_harness.cpp")], check=True)
    lib = C.CDLL(so)
    lib.ml_first_hit_packed.argtypes = lib.ml_first_hit_bytes.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_int]
    lib.ml_pack4.argtypes = [C.c_char_p]
    lib.ml_bulk_check.argtypes = [C.c_uint64, C.c_long, C.c_char_p, C.c_int]
    lib.ml_bulk_check.restype = C.c_long
    return lib


def hamming(s1, s2):
    return sum(c1 != c2 for c1, c2 in zip(s1, s2))           # V2:72-73


def first_hit(window, e, whitelist):
    hits = [i for i, w in enumerate(whitelist) if hamming(window, w) <= e]   # V2:302-304: the filter keeps file order
    return hits[0] if hits else -1                            # V2:327-329: [0] of the list


def test_pack4_codes_and_base_test(harness):
    for a in range(256):
        four = bytes([a, ord("A"), ord("G"), (a * 7 + 3) & 0xFF])
        got = harness.ml_pack4(four)
        code, bad = got & 0xFF, got >> 8
        assert code == sum((((c >> 1) & 3) << (2 * j)) for j, c in enumerate(four))
        assert bad == sum((0 if c in b"ACGT" else 1) << j for j, c in enumerate(four))
    # the 2-bit codes tell the four bases apart
    assert len({(c >> 1) & 3 for c in b"ACGT"}) == 4


def test_packed_rule_equals_the_reference_formula(harness):
    """Windows of every length 0..8 (the rest of the 12 loaded bytes is whatever follows in the record), clean, with N, with
    lower case, at 0..3 errors: the packed rule, voted on by 32 lanes, gives the index the reference's list comprehension gives."""
    rng = random.Random(20261020)
    wl_raw = "".join(WL).encode()
    alphabet = "ACGT" * 6 + "NNnacgt\n+E6/<#@ "
    for case in range(6000):
        if case % 3:
            base = list(rng.choice(WL))
            for _ in range(rng.choice([0, 1, 1, 2, 2, 3, 4])):
                base[rng.randrange(8)] = rng.choice(alphabet)
            window = "".join(base)
        else:
            window = "".join(rng.choice(alphabet) for _ in range(8))
        tail = "".join(rng.choice(alphabet) for _ in range(4))
        n, e = rng.choice([8, 8, 7, 6, 5, 4, 3, 2, 1, 0]), rng.randrange(4)
        want = first_hit(window[:n], e, WL)
        got = harness.ml_first_hit_packed((window + tail).encode("latin-1"), n, e, wl_raw, len(WL))
        assert got == want, (window, n, e, got, want)
    # an empty window is at distance 0 from everything: the first entry
    assert harness.ml_first_hit_packed(b"NNNNNNNNNNNN", 0, 0, wl_raw, len(WL)) == 0
    # the bytes behind a short window do not count, whatever they are
    w = WL[17]
    assert harness.ml_first_hit_packed((w[:5] + "\n+\nAAAA").encode(), 5, 0, wl_raw, len(WL)) == first_hit(w[:5], 0, WL)


def test_packed_rule_in_bulk_and_on_other_lists(harness):
    wl_raw = "".join(WL).encode()
    assert harness.ml_bulk_check(0x5EED, 200_000, wl_raw, len(WL)) == 0
    rng = random.Random(7)
    for n_wl in (1, 31, 32, 33, 64, 97, 128):                 # list sizes around the 32-entry rounds of votes
        other = "".join(rng.choice("ACGT") for _ in range(8 * n_wl)).encode()
        assert harness.ml_bulk_check(n_wl, 20_000, other, n_wl) == 0
