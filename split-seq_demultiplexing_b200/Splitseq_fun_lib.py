#!/usr/bin/env python
"""Drop-in for the reference's python_only_tool/Splitseq_fun_lib.py (LIB).  The GPU pipeline shards by record
ranges on the device and does not need chunk files; split_fastq{F,R}_fun are kept (plain file I/O, same files,
same names) for callers that still invoke them.  The aligner wrappers (LIB:126-181: STAR, featureCounts,
samtools, umi_tools, minimap2) are outside this project's scope and raise."""
import os
import shutil

from ssd_b200 import hostlogic


def _split(split_num, fastq_file, lengthFastq, prefix, tries):
    split_size = hostlogic.tuned_split_size(int(lengthFastq), int(split_num), tries)
    print("The split size after tuning is " + str(split_size))
    if split_size % 4 != 0:
        print("WARNING!!! Unable to split input fastq without read loss, please try again with a different number of cores.")
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
