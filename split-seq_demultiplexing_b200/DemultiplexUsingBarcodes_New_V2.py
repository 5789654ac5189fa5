#!/usr/bin/env python
"""Drop-in for the reference's python_only_tool/DemultiplexUsingBarcodes_New_V2.py (V2): same function names,
arguments, side effects and printed summary, with the work done on a B200 through libssd_b200.so.

    demultiplex_fun(...)       V2:32    appends the barcode-matched, re-headed read pairs to <outputDir>/MergedCells_1.fastq
    calc_demux_results(...)    V2:384   keeps the cells with >= -m distinct UMIs, appends them to MergedCells_passing.fastq

There is no CPU fallback: without a usable GPU both functions raise.  Differences from the reference that a caller
can observe are listed in DESIGN.md ("contract")."""
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)

import ssd_b200  # noqa: E402
from ssd_b200 import hostlogic  # noqa: E402


def _read_file(path):
    return np.fromfile(path, dtype=np.uint8)


def demultiplex_fun(inputFastqF, inputFastqR, outputDir, numReadsBin, errorThreshold, performanceMetrics,
                    readsPerCellThreshold, positionDetection, verbose=False):
    """V2:32-381.  numReadsBin only shapes the reference's memory use; readsPerCellThreshold is unused there too
    (CLI3:62 passes a constant 10).  Records come out in input order (the reference's order is arbitrary)."""
    whitelist = hostlogic.load_whitelist(".")                                   # V2:51-68
    print("binIterator is set to " + str(int(numReadsBin * 4)))                # V2:80-81
    print("Using default barcode positions...")
    print("Barcode position setting is " + str(positionDetection))
    positions = dict(hostlogic.DEFAULT_POSITIONS)                               # V2:155-162
    if bool(positionDetection) is True:
        print("Learning barcode positions to replace defaults...")
        with open(hostlogic.LEARNER_FILE, "r") as fh:                           # V2:171
            learned = hostlogic.learn_positions(fh)
        for line in hostlogic.positions_report(learned):
            print(line)
        learned.pop("_found")
        positions = learned
    dm = ssd_b200.Demultiplexer(whitelist, errors=int(errorThreshold), min_count=1, positions=positions)
    try:
        # file to file inside the library: pread into pinned rings, overlapped uploads, chunked download + append (V2:367)
        _, stats = dm.run_files(str(inputFastqF), str(inputFastqR), str(outputDir + "/MergedCells_1.fastq"), ssd_b200.EMIT_MATCHED, append=True)
        print("The linesInInputFastq value is set to " + str(4 * stats["n_records_r1"]))  # V2:88-90 (complete records)
    finally:
        dm.close()
    sys.stdout.flush()


def calc_demux_results(outputDir, performanceMetrics, readsPerCellThreshold):
    """V2:384-462: cells = distinct header.split("_")[1]; keep a cell when it has >= int(-m) distinct UMIs
    (header.split("_")[2]); copy the records of kept cells, in file order, to MergedCells_passing.fastq (append)."""
    if performanceMetrics is not True:  # V2:391: the whole body hangs off this flag
        return
    whitelist = hostlogic.load_whitelist(".")
    dm = ssd_b200.Demultiplexer(whitelist, errors=0, min_count=int(readsPerCellThreshold))
    try:
        _, stats = dm.filter_annotated_file(str(outputDir + "/MergedCells_1.fastq"), str(outputDir + "/MergedCells_passing.fastq"),
                                            append=True)                      # V2:415 (append mode)
    finally:
        dm.close()
    for line in hostlogic.summary_lines(stats["n_cells"], readsPerCellThreshold, stats["n_passing_cells"]):
        print(line)                                                             # V2:458-459
    sys.stdout.flush()


if __name__ == '__main__':
    demultiplex_fun()
    calc_demux_results()
