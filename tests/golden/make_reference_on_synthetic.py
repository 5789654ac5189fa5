#!/usr/bin/env python
"""The UNMODIFIED reference CLI (/root/reference/python_only_tool/splitseqdemultiplex_0.2.3.py, run as a subprocess from a
scratch copy of its directory, never copied into this repo) on the head of the synthetic bench workloads (BASELINE configs 2
and 3 as oracle/ssd_synth_ref.c generates them): its wall-clock time on this container's cores -- the CPU figure BASELINE.md 3
asks for, measured on the reference itself -- and the canonical digest and counts of its output, which tests/test_oracle_golden.py
holds the C oracle to (so the oracle is pinned to the reference on the bench workload's own shape, not only on the bundled and
hand-built sets).

Run only in the authoring container (the GPU box has no /root/reference):

    python tests/golden/make_reference_on_synthetic.py

Output (committed): reference_on_synthetic.json
"""
import hashlib
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import fastq_util as fu  # noqa: E402
from oracle import ssd_oracle_py as orc  # noqa: E402

REF_TOOL = "/root/reference/python_only_tool"
CASES = [
    # name, pairs, e, m, generator keywords (bench.py CONFIGS[2] / CONFIGS[3]; a smaller cell pool for the config-3 head so that
    # some cells reach m)
    dict(name="config2-head", pairs=200_000, e=1, m=10, spec=dict(seed=0x5EED0001)),
    dict(name="config3-head", pairs=100_000, e=2, m=10, spec=dict(seed=0x5EED0002, cell_pool=20_000, p_n=0.01, p_tie=0.01, p_trunc=0.005)),
]


def run_cli3(r1: bytes, r2: bytes, e: int, m: int, cores: int):
    """CLI3 the way its README runs it; returns (seconds, MergedCells_passing bytes, the two summary lines)."""
    tmp = tempfile.mkdtemp(prefix="ssdref_")
    try:
        tool = os.path.join(tmp, "tool")
        shutil.copytree(REF_TOOL, tool, ignore=shutil.ignore_patterns("*.saf", "*.fastq", "__pycache__"))
        with open(os.path.join(tool, "F.fastq"), "wb") as fh:
            fh.write(r1)
        with open(os.path.join(tool, "R.fastq"), "wb") as fh:
            fh.write(r2)
        os.makedirs(os.path.join(tool, "__pycache__"), exist_ok=True)  # CLI3:100 removes it and raises when it is absent
        n_lines = r2.count(b"\n")
        cmd = [sys.executable, "splitseqdemultiplex_0.2.3.py", "-n", str(cores), "-e", str(e), "-m", str(m), "-1", "a", "-2", "b", "-3", "c",
               "-f", "F.fastq", "-r", "R.fastq", "-o", "results", "-a", "", "-b", "100000", "-p", "True", "-l", str(n_lines)]
        t0 = time.perf_counter()
        proc = subprocess.run(cmd, cwd=tool, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        dt = time.perf_counter() - t0
        if proc.returncode != 0:
            raise RuntimeError("reference CLI failed:\n" + proc.stderr[-2000:])
        with open(os.path.join(tool, "results", "MergedCells_passing.fastq"), "rb") as fh:
            passing = fh.read()
        tail = [l for l in proc.stdout.splitlines() if l.startswith("The total number of unique barcodes") or l.startswith("The number of barcodes passing")]
        return dt, passing, tail[-2:]
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def main():
    wl = fu.load_whitelist()
    cores = os.cpu_count() or 1
    out = {"reference": "python_only_tool/splitseqdemultiplex_0.2.3.py (unmodified), -n %d -b 100000 -p True, position detection on (its default)" % cores,
           "host": {"cores": cores, "cpu": next((l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")), "?"),
                    "python": sys.version.split()[0]},
           "cases": []}
    for c in CASES:
        a, b = orc.synth_pairs(wl, c["pairs"], read_len=150, **c["spec"])
        r1, r2 = a.tobytes(), b.tobytes()
        dt, passing, tail = run_cli3(r1, r2, c["e"], c["m"], cores)
        t0 = time.perf_counter()
        ref = orc.run_cli(r1, r2, wl, c["e"], c["m"], n_workers=cores, length_fastq=r2.count(b"\n"))
        dt_oracle = time.perf_counter() - t0
        cells, passing_cells = (int(re.search(r"(\d+)\s*$", l).group(1)) for l in tail)
        case = {"name": c["name"], "pairs": c["pairs"], "e": c["e"], "m": c["m"], "spec": {k: (hex(v) if k == "seed" else v) for k, v in c["spec"].items()},
                "input_sha256": [hashlib.sha256(r1).hexdigest(), hashlib.sha256(r2).hexdigest()],
                "seconds": round(dt, 2), "pairs_per_s": round(c["pairs"] / dt, 1), "pairs_per_s_per_core": round(c["pairs"] / dt / cores, 1),
                "retained_records": passing.count(b"\n") // 4, "cells": cells, "passing_cells": passing_cells,
                "passing_bytes": len(passing), "passing_canonical_sha256": fu.canonical_sha(passing),
                "summary_lines": tail,
                "oracle": {"seconds": round(dt_oracle, 3), "equal": fu.canonical_sha(ref.passing) == fu.canonical_sha(passing)
                           and (ref.n_cells, ref.n_passing_cells, ref.n_retained) == (cells, passing_cells, passing.count(b"\n") // 4)}}
        print(json.dumps(case))
        out["cases"].append(case)
    with open(os.path.join(HERE, "reference_on_synthetic.json"), "w") as fh:
        json.dump(out, fh, indent=1)
        fh.write("\n")


if __name__ == "__main__":
    main()
