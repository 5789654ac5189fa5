#!/usr/bin/env python
"""Drop-in for the reference's python_only_tool/Splitseq_fun_lib.py (LIB).  The GPU pipeline shards by record
ranges on the device and does not need chunk files; split_fastq{F,R}_fun are kept (same files, same names, same bytes)
for callers that still invoke them, and copy byte ranges instead of lines whenever that gives the same bytes.  The aligner wrappers (LIB:126-181: STAR, featureCounts,
samtools, umi_tools, minimap2) are outside this project's scope and raise."""
import os
import shutil

from ssd_b200 import hostlogic


def _split_by_ranges(fastq_file, split_size, prefix):
    """The reference copies line.strip() + "\n" line by line (LIB:36, 83; ~0.4 M read pairs/s in CPython).  When every line
    of the file already is its own strip() + "\n" (ssd_fastq_lines_are_stripped: ASCII, no '\r', no blank next to a newline,
    final newline) the chunk files are byte ranges of the input: the range bounds come from the library's newline counter and
    the bytes are moved with sendfile.  Returns False when the file is not of that kind (the caller then copies line by line)."""
    import ctypes as C
    from ssd_b200 import _lib as L
    from ssd_b200 import filebatch
    plain, n_lines = C.c_int(0), C.c_uint64(0)
    if L.load().ssd_fastq_lines_are_stripped(os.fsencode(fastq_file), C.byref(plain), C.byref(n_lines)) != L.OK or not plain.value:
        return False
    n_chunks = -(-n_lines.value // split_size)
    starts = [k * split_size // 4 for k in range(1, n_chunks)]           # records at which chunks 1, 2, ... begin
    offs = filebatch.record_cuts(fastq_file, records=starts)[1] if starts else []
    bounds = [0] + offs + [os.path.getsize(fastq_file)]
    with open(fastq_file, "rb") as src:
        for k in range(n_chunks):
            with open("./%s_%d" % (prefix, k), "ab") as dst:            # append mode, LIB:35 / 82
                off, left = bounds[k], bounds[k + 1] - bounds[k]
                while left > 0:
                    try:
                        n = os.sendfile(dst.fileno(), src.fileno(), off, min(left, 1 << 30))
                    except OSError:
                        src.seek(off)
                        n = dst.write(src.read(min(left, 64 << 20)))
                    if n <= 0:
                        raise OSError("short copy while splitting %s" % fastq_file)
                    off += n
                    left -= n
    return True


def _split(split_num, fastq_file, lengthFastq, prefix, tries):
    split_size = hostlogic.tuned_split_size(int(lengthFastq), int(split_num), tries)
    print("The split size after tuning is " + str(split_size))
    if split_size % 4 != 0:
        print("WARNING!!! Unable to split input fastq without read loss, please try again with a different number of cores.")
    elif not os.environ.get("SSD_SPLIT_LINE_BY_LINE") and _split_by_ranges(fastq_file, split_size, prefix):
        return
    counter, split_counter, outfile = 0, 0, None
    with open(fastq_file, "r") as infile:
        for line in infile:
            if outfile is None or counter == split_size:
                if outfile is not None:
                    outfile.close()
                    split_counter += 1
                outfile = open("./%s_%d" % (prefix, split_counter), "a")   # append mode, LIB:35 / 82
                counter = 0
            outfile.write(line.strip() + "\n")                              # LIB:36 / 83
            counter += 1
    if outfile is not None:
        outfile.close()


def split_fastqF_fun(split_num, fastq_file, lengthFastq):
    """LIB:6-50."""
    _split(split_num, fastq_file, lengthFastq, "split_fastq_F", 100)


def split_fastqR_fun(split_num, fastq_file, lengthFastq):
    """LIB:53-110: the chunks plus the 1001-line sample the position learner reads."""
    _split(split_num, fastq_file, lengthFastq, "split_fastq_R", 10000)
    with open(hostlogic.LEARNER_FILE, "w") as outfile:
        outfile.writelines(hostlogic.learner_sample(fastq_file))


def remove_file_fun(filename):
    os.remove(filename)                                                      # LIB:113-115


def remove_dir_fun(filename):
    shutil.rmtree(filename)                                                  # LIB:117-119


def _out_of_scope(*_args, **_kwargs):
    raise NotImplementedError("alignment / quantification (STAR, featureCounts, samtools, umi_tools, minimap2) is outside the "
                              "scope of the B200 demultiplexer; run those tools on MergedCells_passing.fastq")


run_minimap2_alignment_fun = run_star_alignment_fun = run_featureCounts_SAF_fun = run_samtools_fun = run_umi_tools_fun = _out_of_scope
