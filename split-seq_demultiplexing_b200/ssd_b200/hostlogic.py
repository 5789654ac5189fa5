"""Host-side logic of the reference's fast path that is not byte crunching: whitelist files, the position
learner and argument conventions.  Mirrors python_only_tool/DemultiplexUsingBarcodes_New_V2.py (V2) and
python_only_tool/Splitseq_fun_lib.py (LIB) of the reference; file:line citations are relative to its root."""
from __future__ import annotations

import itertools
import os
from collections import Counter
from typing import Sequence, Optional, Dict, Iterable, List, Tuple

ROUND1_FILE = "Round1_barcodes_new5.txt"   # V2:56  (opened relative to the CWD, like the reference)
ROUND2_FILE = "Round2_barcodes_new4.txt"   # V2:62
ROUND3_FILE = "Round3_barcodes_new4.txt"   # V2:66
LEARNER_FILE = "position_learner_fastqr.fastq"  # V2:171, written by LIB:99-110
ANCHORS = ("CATTCG", "ATCCAC", "GTGGCC")   # Round1/2/3 static sequences, V2:36-38
DEFAULT_POSITIONS = {"umi": (0, 10), "bc3": (10, 18), "bc2": (48, 56), "bc1": (86, 94)}  # V2:155-162


def load_whitelist(directory: str = ".") -> List[str]:
    """V2:51-68: one list serves all three rounds — characters [8:16] of every line of the Round-1 file.  The
    Round-2/3 files are opened (and must exist) but their content is never used."""
    with open(os.path.join(directory, ROUND1_FILE), "r") as fh:
        wl = [line.rstrip()[8:16] for line in fh]
    for name in (ROUND2_FILE, ROUND3_FILE):
        with open(os.path.join(directory, name), "r") as fh:
            for _ in fh:
                pass
    return wl


def load_whitelists_from(round1: str, round2: str, round3: str) -> Tuple[List[str], List[str], List[str]]:
    """Opt-in extension: take the whitelists from the files given with -1/-2/-3 instead of the fixed names in
    the CWD, one list per round.  Round 1 carries its 8-mer at [8:16] (V2:58), Rounds 2 and 3 at [0:8] (the bundled 13- and
    14-mers start with it).  The reference's fast path matches all three rounds against the Round-1 list (V2:60, 302-304);
    with the bundled files the three lists are equal, so this gives the reference's result -- and with files that differ,
    every window is matched against the list of its own round (Demultiplexer(whitelist, whitelist2=, whitelist3=))."""
    out = []
    for path, lo, hi in ((round1, 8, 16), (round2, 0, 8), (round3, 0, 8)):
        with open(path, "r") as fh:
            out.append([line.rstrip()[lo:hi] for line in fh if line.strip()])
    return out[0], out[1], out[2]


def load_whitelist_from(round1: str, round2: str, round3: str) -> List[str]:
    """The one list of the reference's fast path from the -1/-2/-3 files; ValueError when the files disagree (use
    load_whitelists_from and per-round lists then)."""
    wl, w2, w3 = load_whitelists_from(round1, round2, round3)
    if w2 != wl or w3 != wl:
        raise ValueError("the Round 2 / Round 3 files do not carry the 8-mers of %s" % round1)
    return wl


def load_round_names(directory: str = ".") -> Tuple[List[str], List[str], List[str]]:
    """The full-length barcodes of the three rounds (16-, 13- and 14-mers in the bundled files), line i of every file being
    entry i of the whitelist (Round 1 carries its 8-mer at [8:16], Rounds 2 and 3 at [0:8]).  The legacy pipeline names
    its per-cell files after them: round1 + "-" + round2 + "-" + round3 + ".fastq" (demultiplex_using_barcodes.py:364)."""
    out = []
    for name in (ROUND1_FILE, ROUND2_FILE, ROUND3_FILE):
        with open(os.path.join(directory, name), "r") as fh:
            out.append([line.rstrip() for line in fh])
    return out[0], out[1], out[2]


def split_file_name(names: Tuple[List[str], List[str], List[str]], i1: int, i2: int, i3: int) -> str:
    """demultiplex_using_barcodes.py:364: bufferKey = round1 + HYPHEN + round2 + HYPHEN + round3 + FASTQ_EXTENSION."""
    return names[0][i1] + "-" + names[1][i2] + "-" + names[2][i3] + ".fastq"


def learner_sample(fastq_r: str, n_lines: int = 1001) -> List[str]:
    """LIB:99-110: the first 1001 lines of read 2, each strip()ped."""
    with open(fastq_r, "r") as fh:
        return [line.strip() + "\n" for line in itertools.islice(fh, n_lines)]


def learn_positions(lines: Iterable[str]) -> Dict[str, Tuple[int, int]]:
    """V2:164-199.  Over sequence lines (index % 4 == 1) keep find() positions >= 17 of the three static
    sequences, take the mode of each (first maximum in set iteration order, exactly what
    max(set(l), key=l.count) returns), derive the four slices.  ValueError when a list is empty."""
    found: List[List[int]] = [[], [], []]
    for n, line in enumerate(lines):
        if n % 4 == 1:
            for k, anchor in enumerate(ANCHORS):
                p = line.find(anchor)
                if p >= 17:
                    found[k].append(p)
    return positions_from_found(found)


def positions_from_found(found: Sequence[Sequence[int]]) -> Dict[str, Tuple[int, int]]:
    """The second half of V2:164-199: found[k] = the find() positions >= 17 of static sequence k, in line order; the mode of
    each list (first maximum in set iteration order, exactly what max(set(l), key=l.count) returns) gives the four slices."""
    modes = []
    for lst in found:
        if not lst:
            raise ValueError("max() arg is an empty sequence")  # what V2:181-183 raises
        counts = Counter(lst)
        best = None
        for v in set(lst):
            if best is None or counts[v] > counts[best]:
                best = v
        modes.append(best)
    p1, p2, p3 = modes
    return {"umi": (p3 - 18, p3 - 8), "bc3": (p3 - 8, p3), "bc2": (p2 - 8, p2), "bc1": (p1 + 6, p1 + 14),
            "_found": (p1, p2, p3)}


def learn_positions_device(fastq_r: str, n_lines: int = 1001, device: Optional[int] = None) -> Dict[str, Tuple[int, int]]:
    """learn_positions with the searching done on the GPU (ssd_anchor_positions): the head of the file goes to the device, one
    thread per sequence line runs the three find()s, the host takes the modes.  Same result as learner_sample +
    learn_positions on files whose lines end in "\\n" or "\\r\\n" (the library's input contract)."""
    import ctypes as C
    import torch
    from . import _lib as L
    lib = L.load()
    size = os.path.getsize(fastq_r)
    want = min(size, 1 << 20)
    while True:  # enough of the file to hold n_lines lines
        with open(fastq_r, "rb") as fh:
            head = fh.read(want)
        if head.count(b"\n") >= n_lines or want >= size:
            break
        want = min(size, want * 4)
    if device is not None:
        torch.cuda.set_device(device)
    d = torch.frombuffer(bytearray(head), dtype=torch.uint8).cuda() if head else torch.empty(0, dtype=torch.uint8, device="cuda")
    cap = n_lines // 4 + 1
    found = (C.c_int32 * (3 * cap))()
    n_seq = C.c_uint32()
    anchors = (C.c_char_p * 3)(*[a.encode() for a in ANCHORS])
    rc = lib.ssd_anchor_positions(C.c_void_p(d.data_ptr() if d.numel() else 0), d.numel(), n_lines, anchors, found, cap, C.byref(n_seq))
    if rc != L.OK:
        raise RuntimeError("ssd_anchor_positions failed (%d)" % rc)
    return positions_from_found([[p for p in found[k * cap: k * cap + n_seq.value] if p >= 17] for k in range(3)])


def positions_report(pos: Dict[str, Tuple[int, int]]) -> List[str]:
    """The lines V2:184-199 prints."""
    p1, p2, p3 = pos["_found"]
    return ["Extracted position1 = %d" % p1, "Extracted position2 = %d" % p2, "Extracted position3 = %d" % p3,
            "UMI position has been extracted as %d:%d" % pos["umi"],
            "Barcode3 position has been extracted as %d:%d" % pos["bc3"],
            "Barcode2 position has been extracted as %d:%d" % pos["bc2"],
            "Barcode1 position has been extracted as %d:%d" % pos["bc1"]]


def tuned_split_size(length_fastq: int, split_num: int, tries: int = 100) -> int:
    """LIB:11-19 / 58-66: lengthFastq / N bumped until it is a multiple of 4."""
    size = length_fastq / int(split_num)
    for _ in range(tries):
        if size % 4 != 0:
            size = int(size) + 1
    return int(size)


def lone_cr_in_head(head: bytes) -> bool:
    """True when the first bytes of a FASTQ file hold a carriage return that is not followed by a line feed (classic Mac line
    ends).  The reference reads its files with Python's universal newlines, which make such a '\r' a line break; the library's
    scans break lines at '\n' only, so a file of that kind holds no records for them.  The last byte of the head is
    not judged (its '\n' may be the next byte)."""
    i = head.find(b"\r")
    while 0 <= i < len(head) - 1:
        if head[i + 1:i + 2] != b"\n":
            return True
        i = head.find(b"\r", i + 1)
    return False


def reject_lone_cr(*heads: bytes):
    """Raise instead of returning an empty result for inputs whose lines end in a lone '\r' (see lone_cr_in_head)."""
    from .engine import MalformedFastq
    for k, head in enumerate(heads):
        if lone_cr_in_head(head):
            raise MalformedFastq("input %d: lines end in a lone carriage return (classic Mac line ends); convert the file to "
                                 "\\n or \\r\\n line ends first (the reference's universal-newline reading is not reproduced)" % (k + 1))


def summary_lines(n_cells: int, threshold, n_passing: int) -> List[str]:
    """V2:458-459, verbatim."""
    return ["The total number of unique barcodes detected was " + str(n_cells),
            "The number of barcodes passing the minimum UMI threshold of " + str(threshold) + " was " + str(n_passing)]


def format_cell_table(table: Dict[str, Tuple[int, int]], threshold: int) -> str:
    """Per-cell summary as tab-separated text (the counts without a FASTQ round trip): one line per detected
    cell, in whitelist-independent (lexicographic) order of the 24-base cell id -- cell, reads, distinct UMIs, and whether
    the cell passes the -m threshold on distinct UMIs (V2:411)."""
    lines = ["cell\treads\tdistinct_umis\tpassing\n"]
    for cell in sorted(table):
        reads, distinct = table[cell]
        lines.append("%s\t%d\t%d\t%d\n" % (cell, reads, distinct, 1 if distinct >= int(threshold) else 0))
    return "".join(lines)


def format_cell_umi_table(whitelist, cells, umis, reads) -> bytes:
    """The per-cell x UMI table as tab-separated text: `cell <tab> UMI <tab> reads` with the 24-base cell id spelled from the
    whitelist (cell = (i1 * n2 + i2) * n3 + i3, printed bc1 bc2 bc3 like V2:139).  whitelist: one list of strings, or the
    three lists of the rounds.  numpy arrays in, bytes out; vectorised, so that tens of millions of rows take seconds."""
    import numpy as np
    rounds = [whitelist] * 3 if isinstance(whitelist[0], str) else list(whitelist)
    wl = [np.array([w.encode() for w in r], dtype="S8") for r in rounds]
    n2, n3 = len(rounds[1]), len(rounds[2])
    cells = np.asarray(cells, dtype=np.int64)
    ids = np.char.add(np.char.add(wl[0][cells // (n2 * n3)], wl[1][(cells // n3) % n2]), wl[2][cells % n3])
    rows = np.char.add(np.char.add(np.char.add(np.char.add(ids, b"\t"), np.asarray(umis)), b"\t"), np.asarray(reads).astype("S10"))
    return b"cell\tumi\treads\n" + b"".join(r + b"\n" for r in rows.tolist())
