#!/usr/bin/env python
"""Golden vector for the Extract_BC_UMI-style headers (SSD_HEADER_EXTRACT): the UNMODIFIED reference script
/root/reference/Extract_BC_UMI.py (run as a subprocess, never copied) on a per-cell pair of files made of the first 300
read pairs of the bundled test FASTQs, plus three hand-made names (a tab, several blanks and a third token, trailing blanks).

Run only in the authoring container (the GPU box has no /root/reference):

    python tests/golden/make_golden_extract.py

Output (committed): extract_bc_umi_case.json.gz = {"cell": the barcode the script takes from the file name, "n": records,
"extra_names": the hand-made read-1 names, "output": the text of <cell>_1.fastq}
"""
import gzip
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import fastq_util as fu  # noqa: E402

SCRIPT = "/root/reference/Extract_BC_UMI.py"
N = 300
EXTRA_NAMES = ["@x.1\tleft/1", "@x.2   two  three four/1", "@x.3 7/1  "]


def inputs():
    """The two per-cell files the script reads (first N bundled pairs + the hand-made names on copies of the first pairs)."""
    r1, r2 = fu.load_bundled()
    l1, l2 = r1.split(b"\n"), r2.split(b"\n")
    f, r = l1[: 4 * N], l2[: 4 * N]
    for k, name in enumerate(EXTRA_NAMES):
        f += [name.encode()] + l1[4 * k + 1: 4 * k + 4]
        r += [name.split()[0].encode() + b" mate/2"] + l2[4 * k + 1: 4 * k + 4]  # the same id: the fast path checks mates (V2:350)
    return b"\n".join(f) + b"\n", b"\n".join(r) + b"\n"


def main():
    fwd, rev = inputs()
    with tempfile.TemporaryDirectory() as tmp:
        name = "AAAAAAAA-CCCCCCCC-GGGGGGGG.fastq"
        open(os.path.join(tmp, name), "wb").write(fwd)
        open(os.path.join(tmp, name + "-MATEPAIR"), "wb").write(rev)
        subprocess.run([sys.executable, SCRIPT, "-F", name, "-R", name + "-MATEPAIR"], cwd=tmp, check=True)
        out = open(os.path.join(tmp, "AAAAAAAA-CCCCCCCC-GGGGGGGG_1.fastq"), "rb").read()
    case = {"cell": "AAAAAAAACCCCCCCCGGGGGGGG", "n": N + len(EXTRA_NAMES), "extra_names": EXTRA_NAMES, "output": out.decode()}
    with gzip.GzipFile(os.path.join(HERE, "extract_bc_umi_case.json.gz"), "wb", mtime=0) as fh:
        fh.write(json.dumps(case, sort_keys=True).encode())
    print("extract_bc_umi_case.json.gz:", case["n"], "records,", len(out), "bytes")


if __name__ == "__main__":
    main()
