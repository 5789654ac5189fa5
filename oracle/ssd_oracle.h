/*
 * ssd_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, CPU restatement of the reference's `--version fast` hot path
 * (paulranum11/SPLiT-Seq_demultiplexing, python_only_tool/DemultiplexUsingBarcodes_New_V2.py and
 * python_only_tool/Splitseq_fun_lib.py).  It exists so that the CUDA product path can be checked
 * bit-for-bit at sizes the Python reference cannot reach.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it; nothing in the product imports,
 * links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this restatement against golden vectors
 * produced by running the unmodified reference in the authoring container
 * (tests/golden/make_golden.py): the 8 bundled-FASTQ digests of SURVEY.md Appendix A, the rule-level
 * cases of Appendix B, and seeded adversarial sets.
 *
 * Input contract (anything else returns SSD_ORACLE_EMALFORMED where the reference would misbehave):
 * ASCII; lines end in "\n" or "\r\n" (a lone "\r" is not a line break here, Python's universal
 * newlines would make it one); every record is 4 lines and its first line starts with '@'.
 */
#ifndef SSD_ORACLE_H
#define SSD_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSD_ORACLE_OK 0
#define SSD_ORACLE_EKEY (-1)       /* KeyError: matched read-2 id without a read-1 mate in the bin (V2:350) */
#define SSD_ORACLE_EMALFORMED (-2) /* outside the input contract */
#define SSD_ORACLE_EINDEX (-3)     /* IndexError: header without two '_' in calc_demux_results (V2:397-398) */
#define SSD_ORACLE_ENOMEM (-4)

typedef struct ssd_oracle_cfg {
    const char *whitelist;   /* n_whitelist entries, 8 bytes each, NUL padded when shorter        */
    const uint8_t *wl_len;   /* length of each entry (0..8); NULL = all 8                          */
    int n_whitelist;
    int umi_start, umi_end;  /* Python slice bounds into the read-2 sequence line, all >= 0       */
    int bc1_start, bc1_end;
    int bc2_start, bc2_end;
    int bc3_start, bc3_end;
    int errors;              /* -e : max Hamming distance per barcode                             */
    long bin_reads;          /* -b : reads per bin (V2:80)                                        */
    /* Per-round lists (extension beyond the reference's fast path, which matches all three windows against the one list above,
     * V2:60, 302-304): whitelist2 / whitelist3 != NULL give the Round2 (bc2) / Round3 (bc3) window its own list, matched by
     * the same first-hit rule.  Parity of this extension is pinned on the restated rule only (the reference cannot run it). */
    const char *whitelist2;
    const uint8_t *wl_len2;
    int n_whitelist2;
    const char *whitelist3;
    const uint8_t *wl_len3;
    int n_whitelist3;
} ssd_oracle_cfg;

typedef struct ssd_oracle_cell {
    char cell[64];           /* cell string (24 chars for 8-mer whitelists)                       */
    uint64_t reads;
    uint64_t distinct_umis;
    int passing;
} ssd_oracle_cell;

typedef struct ssd_oracle_out {
    char *merged1;           /* MergedCells_1.fastq text (records in read-2 order inside each bin) */
    size_t merged1_len;
    char *passing;           /* MergedCells_passing.fastq text                                     */
    size_t passing_len;
    int32_t *rec_cell;       /* per read-2 record (file order): i1*n*n+i2*n+i3, or -1 if rejected  */
    uint64_t n_records;      /* read-2 records parsed                                              */
    uint64_t n_matched, n_retained;
    ssd_oracle_cell *cells;  /* sorted by cell string                                              */
    uint64_t n_cells, n_passing_cells;
    int error;
    char error_msg[256];
} ssd_oracle_out;

/* DemultiplexUsingBarcodes_New_V2.demultiplex_fun (V2:32-381) on in-memory files.
 * Fills out->merged1, rec_cell, n_records, n_matched.  Appends to out->merged1 if it is already set
 * (the reference opens MergedCells_1.fastq in append mode, V2:367). */
int ssd_oracle_demultiplex(const ssd_oracle_cfg *cfg, const char *r1, size_t r1_len, const char *r2,
                           size_t r2_len, ssd_oracle_out *out);

/* DemultiplexUsingBarcodes_New_V2.calc_demux_results (V2:384-462) on out->merged1: fills cells,
 * n_cells, n_passing_cells, passing, n_retained.  min_umis is -m. */
int ssd_oracle_calc_demux_results(ssd_oracle_out *out, long min_umis);

/* Splitseq_fun_lib.split_fastq{F,R}_fun (LIB:6-97): cut `buf` into at most n_chunks pieces of
 * split_size lines, each line strip()ped and "\n"-terminated.  chunk[i]/chunk_len[i] are malloc'ed.
 * Returns the number of chunks produced (can be < n_chunks on tiny inputs). */
int ssd_oracle_split_fastq(const char *buf, size_t len, int n_chunks, long length_fastq, char **chunk,
                           size_t *chunk_len);

/* CLI3 (splitseqdemultiplex_0.2.3.py:52-70): split both files into n_workers chunks, run
 * demultiplex per chunk (OpenMP threads stand in for the joblib workers), concatenate in chunk
 * order, then calc_demux_results. */
int ssd_oracle_run_cli(const ssd_oracle_cfg *cfg, const char *r1, size_t r1_len, const char *r2,
                       size_t r2_len, long length_fastq, int n_workers, long min_umis,
                       ssd_oracle_out *out);

/* Position learner (V2:164-199) on the text of position_learner_fastqr.fastq.  Writes the three
 * found static-sequence positions; returns SSD_ORACLE_EMALFORMED where the reference raises
 * ValueError (empty candidate list).  Mode ties resolve to the smallest position. */
int ssd_oracle_learn_positions(const char *sample, size_t len, int pos[3]);

/* Collapse rule of Collapse_RanHex_Odt.py:44-71 on a set of cells given as (i1,i2,i3) index triples:
 * for pair k, if both (k + n_pairs, i2, i3) and (k, i2, i3) are present, the RanHex cell is merged
 * into the OligoDT one.  present[] is an n*n*n byte map; remap[] receives the target cell id. */
void ssd_oracle_collapse_map(const uint8_t *present, int n, int n_pairs, uint32_t *remap);

void ssd_oracle_free(ssd_oracle_out *out);

#ifdef __cplusplus
}
#endif
#endif
