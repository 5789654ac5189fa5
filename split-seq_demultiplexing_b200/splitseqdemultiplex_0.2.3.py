#!/usr/bin/env python
"""Drop-in for the reference's python_only_tool/splitseqdemultiplex_0.2.3.py (CLI3): same options, same outputs
(<outputDir>/MergedCells_passing.fastq, the two summary lines), the demultiplexing done on B200 GPUs.

CLI3 splits the FASTQs into -n chunk files, runs demultiplex_fun per chunk in worker processes, filters the
concatenation, then deletes the intermediates (CLI3:52-81).  Here the whole pair of files goes through one fused
device pipeline (scan, match, filter, write); -n is accepted for compatibility and ignored: one process drives one
GPU, and `torchrun --nproc-per-node N splitseqdemultiplex_0.2.3.py ...` shares the pair of files between N GPUs
(ssd_b200.filebatch.run_files_sharded: contiguous runs of record-aligned batches per rank, -m decided globally).  The bash wrappers' extra flags (-v fast, -c, -t, -g) are accepted too.
"""
import argparse
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)

import ssd_b200  # noqa: E402
from ssd_b200 import hostlogic  # noqa: E402


def build_parser():
    p = argparse.ArgumentParser()
    # CLI3:22-38
    p.add_argument('-n', '--numCores', required=True, help='accepted for compatibility (the reference: number of CPUs)')
    p.add_argument('-e', '--errors', required=True, help='The max number of errors permissible per barcode "-e = 1" is recommended.')
    p.add_argument('-m', '--minReads', required=True, help='The minimum number of distinct UMIs per cell for a cell to be retained.')
    p.add_argument('-1', '--round1Barcodes', required=True, help='parsed and ignored like the reference (V2 opens Round1_barcodes_new5.txt in the CWD)')
    p.add_argument('-2', '--round2Barcodes', required=True, help='parsed and ignored like the reference')
    p.add_argument('-3', '--round3Barcodes', required=True, help='parsed and ignored like the reference')
    p.add_argument('-f', '--fastqF', required=True, help='filepath to the forward .fastq file')
    p.add_argument('-r', '--fastqR', required=True, help='filepath to the reverse .fastq file')
    p.add_argument('-o', '--outputDir', required=True, help='name of the output directory')
    p.add_argument('-a', '--align', required=True, help='any non-empty value asks for alignment, which is out of scope here: pass ""')
    p.add_argument('-x', '--starGenome', required=False)
    p.add_argument('-s', '--geneAnnotation', required=False)
    p.add_argument('-i', '--indexType', required=False)
    p.add_argument('-b', '--numReadsBin', required=True, help='accepted for compatibility (the reference: reads per flush)')
    p.add_argument('-p', '--performanceMetrics', required=True, help='any non-empty string is True, like the reference (CLI3:61)')
    p.add_argument('-l', '--lengthFastq', required=True, help='accepted for compatibility (the reference: lines of the input)')
    p.add_argument('-k', '--positionDetection', required=False, dest='positionDetection', action='store_false',
                   help='give -k to turn automated barcode position detection OFF (store_false, CLI3:38)')
    # bash wrapper vocabulary (splitseqdemultiplex_0.2.2.sh:83-193)
    p.add_argument('-v', '--version', required=False, default='fast',
                   help='"fast" (merged output) or "split" (the same records as one file per cell, legacy naming)')
    p.add_argument('-c', '--collapseRandomHexamers', required=False, default='false',
                   help='"true": fold RanHex Round1 barcodes into their OligoDT partners (Collapse_RanHex_Odt.py); the reference\'s fast path ignores it')
    p.add_argument('--useBarcodeFiles', action='store_true',
                   help='extension: read the whitelist from the -1/-2/-3 files instead of Round*_barcodes_new*.txt in the CWD')
    p.add_argument('--cellTable', required=False,
                   help='extension: also write cell / reads / distinct UMIs / passing as tab-separated text to this path')
    p.add_argument('--headerStyle', required=False, default='fast', choices=['fast', 'extract'],
                   help='extension: "extract" writes the headers of Extract_BC_UMI.py (id_cell_umi, blank, rest of the name) instead of '
                        'the fast path\'s id_cell_umi')
    p.add_argument('--cellUmiTable', required=False,
                   help='extension: also write cell / UMI / reads (one row per distinct pair of the passing cells) as tab-separated '
                        'text to this path (one GPU, input within one batch)')
    p.add_argument('-t', '--targetMemory', required=False)
    p.add_argument('-g', '--granularity', required=False)
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    if str(args.version) not in ('fast', 'split'):
        sys.exit('only --version fast (and its per-cell variant, split) is implemented')
    # Under torchrun (one process per GPU) the ranks share the pair of files: the reference's fan-out over chunk files
    # (CLI3:52-64) with GPUs in place of worker processes; rank 0 does the talking.
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    say = print if rank == 0 else (lambda *a, **k: None)
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    say("Starting Step1: demultiplexing on the GPU.")
    if rank == 0 and not os.path.exists(args.outputDir):
        os.makedirs(args.outputDir)                                            # CLI3:48-49
    whitelist2 = whitelist3 = None
    if args.useBarcodeFiles:
        # one list per round from the -1/-2/-3 files; rounds whose 8-mers equal Round 1's share its list (the bundled files)
        whitelist, w2, w3 = hostlogic.load_whitelists_from(args.round1Barcodes, args.round2Barcodes, args.round3Barcodes)
        whitelist2, whitelist3 = (None if w2 == whitelist else w2), (None if w3 == whitelist else w3)
    else:
        whitelist = hostlogic.load_whitelist(".")
    positions = dict(hostlogic.DEFAULT_POSITIONS)
    if args.positionDetection:                                                  # default True (CLI3:38, 63)
        say("Learning barcode positions to replace defaults...")
        # the three find()s per sequence line run on the GPU (ssd_anchor_positions); SSD_LEARNER=host: the host version
        learned = (hostlogic.learn_positions(hostlogic.learner_sample(args.fastqR)) if os.environ.get("SSD_LEARNER") == "host"
                   else hostlogic.learn_positions_device(args.fastqR, device=local if world > 1 else None))
        for line in hostlogic.positions_report(learned):
            say(line)
        learned.pop("_found")
        positions = learned
    collapse = 48 if str(args.collapseRandomHexamers).lower() == 'true' else 0
    dm = ssd_b200.Demultiplexer(whitelist, errors=int(args.errors), min_count=int(args.minReads), positions=positions,
                                collapse_pairs=collapse, device=local if world > 1 else None, whitelist2=whitelist2, whitelist3=whitelist3, header_style=args.headerStyle)
    # file to file inside the library (pinned rings, overlapped copies; batches of whole records above SSD_BATCH_BYTES).
    # Without -p the reference still demultiplexes (and raises on a missing mate) but never writes the passing file.
    wanted = bool(args.performanceMetrics)                                      # CLI3:70 + V2:391
    out_path = os.path.join(args.outputDir, "MergedCells_passing.fastq") if wanted else os.devnull
    try:
        if str(args.version) == 'split':
            # per-cell files instead of the merged one: the fast path's annotated records bucketed by cell, named like the
            # legacy pipeline's results (round1-round2-round3.fastq, demultiplex_using_barcodes.py:364); one GPU, one call
            if world > 1:
                # every rank buckets its record range; the ranks append to the per-cell files in rank order
                from ssd_b200 import filebatch
                if not wanted:
                    sys.exit('--version split under torchrun needs -p (the per-cell files are its only output)')
                n_mine, stats = filebatch.run_split_sharded(dm, args.fastqF, args.fastqR, args.outputDir, hostlogic.load_round_names("."), world=world)
                say("# of results files [{}]".format(stats["n_passing_cells"]))   # demultiplex_using_barcodes.py:408
            else:
                import numpy as np
                _, stats = dm.run_host(np.fromfile(args.fastqF, dtype=np.uint8), np.fromfile(args.fastqR, dtype=np.uint8), ssd_b200.EMIT_PASSING)
                if wanted:
                    files = dm.write_split_files(args.outputDir, hostlogic.load_round_names("."), stats)
                    say("# of results files [{}]".format(len(files)))               # demultiplex_using_barcodes.py:408
        elif world > 1:
            from ssd_b200 import filebatch
            if not wanted:
                out_path = os.path.join(args.outputDir, ".MergedCells_passing.discard")  # the ranks' parts need a real directory
            _, stats = filebatch.run_files_sharded(dm, args.fastqF, args.fastqR, out_path, ssd_b200.EMIT_PASSING, append=True, world=world)
            if not wanted and rank == 0:
                os.remove(out_path)
        else:
            _, stats = dm.run_files(args.fastqF, args.fastqR, out_path, ssd_b200.EMIT_PASSING, append=True)  # V2:415 (append mode)
        if args.cellTable and rank == 0 and (world == 1 or str(args.version) != 'split'):
            # the global tables are still in place after the (last) batch; on several GPUs every rank holds them (two_pass
            # leaves the exact distinct counts of the exchange by owner rank), rank 0 writes
            dm.write_cell_table(args.cellTable)
        if args.cellUmiTable and world == 1:
            from ssd_b200 import filebatch
            if max(os.path.getsize(args.fastqF), os.path.getsize(args.fastqR)) > filebatch.batch_bytes():
                sys.exit('--cellUmiTable needs the input within one batch (SSD_BATCH_BYTES)')
            dm.write_cell_umi_table(args.cellUmiTable)
    finally:
        dm.close()
    if wanted:
        for line in hostlogic.summary_lines(stats["n_cells"], args.minReads, stats["n_passing_cells"]):
            say(line)
    if bool(args.align):                                                        # CLI3:88
        sys.exit("Alignment setting = True, but alignment is outside the scope of this tool: rerun with -a \"\" and "
                 "feed MergedCells_passing.fastq to STAR yourself")
    say("Alignment setting = False, exiting run")
    return 0


if __name__ == '__main__':
    main()
