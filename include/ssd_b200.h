/*
 * ssd_b200.h — C ABI of libssd_b200.so, the B200 (sm_100a) implementation of the `--version fast`
 * demultiplexing hot path of paulranum11/SPLiT-Seq_demultiplexing.
 *
 * The reference has no FFI: its seam is the Python function surface that
 * python_only_tool/splitseqdemultiplex_0.2.3.py calls.  Each entry point below names the reference
 * interface (file:line, relative to the reference root) whose work it takes over; INTEGRATION.md
 * shows the ctypes stub that binds them from the reference's own modules.
 *
 *   V2   = python_only_tool/DemultiplexUsingBarcodes_New_V2.py
 *   LIB  = python_only_tool/Splitseq_fun_lib.py
 *   CLI3 = python_only_tool/splitseqdemultiplex_0.2.3.py
 *
 * Conventions: plain pointers and sizes only.  Every function returns SSD_OK (0) or a negative
 * SSD_E* code; ssd_last_error() gives the message.  A context is bound to one CUDA device and is not
 * thread safe.  All device work is enqueued on the context's stream (ssd_set_stream); functions that
 * return host-visible numbers synchronise that stream.  There is no CPU fallback: ssd_create fails
 * when no CUDA device is usable.
 */
#ifndef SSD_B200_H
#define SSD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSD_ABI_VERSION 1

#define SSD_OK 0
#define SSD_EINVAL (-1)     /* bad argument / unsupported configuration                              */
#define SSD_ECUDA (-2)      /* CUDA runtime failure                       -> RuntimeError             */
#define SSD_EKEY (-3)       /* matched read-2 record without its read-1 mate (V2:350) -> KeyError     */
#define SSD_EMALFORMED (-4) /* FASTQ outside the input contract           -> ValueError               */
#define SSD_ENOMEM (-5)
#define SSD_ESTATE (-6)     /* call sequence violated                                                 */
#define SSD_ERANGE (-7)     /* output buffer too small                                                */
#define SSD_EINDEX (-8)     /* annotated header without two '_' (V2:397-398)           -> IndexError        */
#define SSD_EIO (-9)        /* a file could not be opened, read or written              -> OSError            */

/* -m semantics.  V2:411 keeps a cell when len(set(UMIs)) >= m; the bash pipeline's step 1
 * (demultiplex_using_barcodes.py:409) counts reads instead. */
#define SSD_FILTER_DISTINCT_UMI 0
#define SSD_FILTER_READS 1

/* ssd_emit selector */
#define SSD_EMIT_MATCHED 0 /* MergedCells_1.fastq      (V2:367-370)  every barcode-matched pair  */
#define SSD_EMIT_PASSING 1 /* MergedCells_passing.fastq (V2:415-454)  pairs of cells passing -m   */

/* Header line of an output record (read 1's name annotated with cell and UMI):
 *   SSD_HEADER_FAST     name.split("/")[0].replace(" ", "").strip() + "_" + cell + "_" + umi                 (V2:134-141)
 *   SSD_HEADER_EXTRACT  name.split()[0] + "_" + cell + "_" + umi + " " + name.split()[1]     (Extract_BC_UMI.py:78-85, what
 *                       users of the legacy `-v merged` pipeline feed to their aligner: the read id stays a token of its
 *                       own and the rest of the name survives as the comment; a name of one token is an IndexError there
 *                       and SSD_EINDEX here).  An extension of the fast path: every record then goes through the exact
 *                       record-by-record code of the scan and the writer (several times slower than the default), and the
 *                       text is for downstream tools, not for V2.calc_demux_results (ssd_annotated_*), whose split("_")
 *                       would take the comment for a part of the UMI. */
#define SSD_HEADER_FAST 0
#define SSD_HEADER_EXTRACT 1

#define SSD_MAX_WHITELIST 128
#define SSD_BARCODE_LEN 8

typedef struct ssd_ctx ssd_ctx;

/* Replaces the constants and file reads at the top of V2.demultiplex_fun (V2:51-68 whitelist,
 * V2:155-199 slice offsets) and the -e / -m options of CLI3:23-24. */
typedef struct ssd_config {
    uint32_t struct_size;     /* = sizeof(ssd_config), for ABI growth                                 */
    int32_t device;           /* CUDA ordinal; -1 = current device                                    */
    int32_t n_whitelist;      /* 1..SSD_MAX_WHITELIST; one list serves all three rounds (V2:302-304)
                                 unless whitelist2 / whitelist3 below say otherwise                   */
    const char *whitelist;    /* n_whitelist x 8 bytes (file order = match priority)                  */
    const uint8_t *whitelist_len; /* per-entry length 0..8, NULL = all 8                              */
    int32_t umi_start, umi_end;   /* Python slice bounds into the read-2 sequence, 0 <= start <= end  */
    int32_t bc1_start, bc1_end;   /* Round1 window (defaults 86:94)                                   */
    int32_t bc2_start, bc2_end;   /* Round2 window (defaults 48:56)                                   */
    int32_t bc3_start, bc3_end;   /* Round3 window (defaults 10:18)                                   */
    int32_t errors;           /* -e                                                                    */
    int64_t min_count;        /* -m                                                                    */
    int32_t filter_mode;      /* SSD_FILTER_*                                                          */
    int32_t collapse_pairs;   /* 0 = off; k = whitelist[k+i] (RanHex) folds into whitelist[i] (OligoDT)
                                 for i < k when both cells survive the filter (Collapse_RanHex_Odt.py:44-71) */
    int32_t record_bytes_hint;/* lower bound on bytes per FASTQ record used to size tables; 0 = 24     */
    int32_t header_style;     /* SSD_HEADER_*; 0 = the fast path's header (this field was `reserved`)  */
    /* Per-round whitelists (extension beyond the reference's fast path; the bash pipeline reads one file per
     * round, demultiplex_using_barcodes.py:350-372).  `whitelist` is matched against the Round1 window (bc1), whitelist2
     * against the Round2 window (bc2), whitelist3 against the Round3 window (bc3); NULL = `whitelist` serves that round
     * too.  Cell id = (i1 * n2 + i2) * n3 + i3, written bc1 bc2 bc3 (V2:139).  collapse_pairs refers to the round-1 list.
     * A caller compiled against the first layout (struct_size up to `header_style`) gets one list. */
    int32_t n_whitelist2, n_whitelist3;
    const char *whitelist2;
    const uint8_t *whitelist2_len;
    const char *whitelist3;
    const uint8_t *whitelist3_len;
} ssd_config;

typedef struct ssd_stats {
    uint64_t n_records_r1, n_records_r2; /* complete 4-line records found                             */
    uint64_t n_matched;        /* records whose three barcodes matched (MergedCells_1 records)         */
    uint64_t n_cells;          /* "total number of unique barcodes detected" (V2:458)                  */
    uint64_t n_passing_cells;  /* "barcodes passing the minimum UMI threshold" (V2:459)                */
    uint64_t n_retained;       /* records of passing cells                                             */
    uint64_t bytes_matched;    /* size of the SSD_EMIT_MATCHED text                                    */
    uint64_t bytes_passing;    /* size of the SSD_EMIT_PASSING text                                    */
    uint64_t n_regular_keys, n_irregular_keys; /* (cell,UMI) keys: ACGT-only full-length UMIs / others */
} ssd_stats;

/* ---- lifetime ------------------------------------------------------------------------------- */
int ssd_abi_version(void);
int ssd_create(const ssd_config *cfg, ssd_ctx **out);
void ssd_destroy(ssd_ctx *ctx);
const char *ssd_last_error(const ssd_ctx *ctx); /* ctx may be NULL: error of the last failed ssd_create */
/* cudaStream_t to enqueue on (0/NULL = the context's own stream). */
int ssd_set_stream(ssd_ctx *ctx, void *cuda_stream);

/* ---- the position learner on the device (V2:164-199, sample of LIB:99-110) ---------------------- */
/* Over the first n_lines (<= 4096; the reference takes 1001) lines of the read-2 text at d_text (DEVICE pointer, current
 * device): for every sequence line (line index % 4 == 1), the str.find() position of each of the three static sequences
 * (NUL-terminated, <= 16 bytes) in the strip()ped line, -1 when absent.  found[k * capacity + t] = anchor k in sequence
 * line t (host array of 3 * capacity), *n_sequence_lines = how many sequence lines the sample holds.  The caller keeps
 * the positions >= 17 and takes their mode (ssd_b200/hostlogic.py: positions_from_found, V2:175-183). */
int ssd_anchor_positions(const void *d_text, size_t len, uint32_t n_lines, const char *const anchors[3], int32_t *found,
                         uint32_t capacity, uint32_t *n_sequence_lines);

/* ---- phase 1: V2.demultiplex_fun minus the file writing (V2:207-358) ------------------------- */
/* FASTQ record-boundary scan of both mates, UMI/barcode extraction at the configured offsets,
 * Hamming<=e first-hit matching, mate-id check, per-cell read histogram, (cell,UMI) key capture.
 * d_r1/d_r2 are DEVICE pointers, 16-byte aligned, to the whole text of the two files (or of the same
 * record range of both); they must stay valid and unchanged until the last ssd_emit of this batch.
 * Starts a new batch: previous per-batch results are discarded. */
int ssd_demux_device(ssd_ctx *ctx, const void *d_r1, size_t r1_len, const void *d_r2, size_t r2_len);

/* V2.calc_demux_results' input side (V2:389-404): the batch is one already annotated FASTQ (MergedCells_1.fastq:
 * headers id_cell_umi).  Cell = the text between the first two '_' (it must be three whitelist entries back to
 * back), UMI = the text up to the next '_' or the end of the line (at most 10 bytes).  Afterwards phases 2 and 3
 * work as usual; SSD_EMIT_PASSING then copies the records of passing cells with every line strip()ped
 * (V2:415-454) and SSD_EMIT_MATCHED copies all records. */
int ssd_annotated_device(ssd_ctx *ctx, const void *d_text, size_t len);
/* The same from a HOST buffer, through to the filtered text (out == NULL and out_cap == 0: stop after the filter). */
int ssd_annotated_host(ssd_ctx *ctx, const void *text, size_t len, void *out, size_t out_cap, size_t *written,
                       ssd_stats *stats);

/* File to file.  ssd_run_files = what V2.demultiplex_fun does with one pair of (chunk) files (V2:32-381):
 * the two FASTQ files are read with pread on two host threads into pinned rings and uploaded as they arrive, the batch
 * is demultiplexed, and the selected output is downloaded in chunks and written to out_path while the next chunk is in
 * flight.  append != 0 appends like the reference (V2:367, 415), else the file is truncated first.
 * ssd_annotated_file = V2.calc_demux_results (V2:384-454) on an annotated file. */
int ssd_run_files(ssd_ctx *ctx, const char *path_r1, const char *path_r2, int which, const char *out_path, int append,
                  size_t *written, ssd_stats *stats);
int ssd_annotated_file(ssd_ctx *ctx, const char *path_in, const char *out_path, int append, size_t *written,
                       ssd_stats *stats);

/* The file path in pieces, for inputs that do not fit one call (more than 64 GiB per file, or more than the GPU holds): the
 * host cuts the files into batches of whole records and drives the batches itself -- one pass when no filter is involved
 * (V2.demultiplex_fun writes every matched pair, V2:367), two passes when -m is (V2.calc_demux_results, V2:389-454: per-cell
 * tables over all batches first, see ssd_b200/filebatch.py).
 *   ssd_load_file_range   bytes [offset, offset + length) of a file (clamped to the file) into the context's input buffer
 *                         `slot` (0: read 1 or the annotated text, 1: read 2), read with pread on eight host threads into
 *                         pinned rings and uploaded as they arrive
 *   ssd_loaded_newlines   the shard phase of several GPUs on one pair of files (what LIB:6-97 does with line counts):
 *                         '\n' bytes per tile of 2^tile_log2 bytes of the range loaded into `slot`, counted on the device,
 *                         copied into tile_counts[capacity]; *n_tiles = how many.  Every rank loads ITS byte range of the
 *                         file and counts; the counts of all ranks (all-gather) give the line index of every tile, hence
 *                         which tile holds the start of record r = line 4 r (ssd_b200/filebatch.py: plan_shards_device).
 *   ssd_demux_loaded      ssd_demux_device on the two loaded ranges; ssd_annotated_loaded: ssd_annotated_device on slot 0
 *   ssd_emit_file         after the filter (ssd_filter_single / ssd_build_filter): the selected output into out_path,
 *                         appended (V2:367, 415) or truncating
 *   ssd_fastq_record_cuts host only, no context: where a FASTQ file can be cut between records ('\n' ends a line, four
 *                         lines make a record).  records_in == NULL: entry 0 is (record 0, byte 0), after a cut at byte b
 *                         the next cut is the first record start at or after b + target_bytes, the last entry is (number
 *                         of complete records, file size).  records_in != NULL (ascending): the offsets at which those
 *                         records start (the file size for records past the end) -- the mate file cut at the same records.
 *                         SSD_ERANGE when `capacity` entries do not suffice. */
int ssd_load_file_range(ssd_ctx *ctx, int slot, const char *path, uint64_t offset, uint64_t length);
int ssd_loaded_newlines(ssd_ctx *ctx, int slot, int tile_log2, uint32_t *tile_counts, uint64_t capacity, uint64_t *n_tiles);
/* '\n' bytes of any text on the device (16-byte aligned): what a caller checks an output with (four lines per record). */
int ssd_count_newlines_device(ssd_ctx *ctx, const void *d_text, size_t len, uint64_t *total);
int ssd_demux_loaded(ssd_ctx *ctx);
int ssd_annotated_loaded(ssd_ctx *ctx);
int ssd_emit_file(ssd_ctx *ctx, int which, const char *out_path, int append, size_t *written);
/* Host only: *plain = 1 when every line of the file already equals Python's line.strip() + "\n" (ASCII, no '\r', no blank
 * next to a newline or at the start, final newline present), i.e. when the reference's splitter (LIB:36, 83) would copy the
 * file's bytes unchanged; *n_lines = its number of lines. */
int ssd_fastq_lines_are_stripped(const char *path, int *plain, uint64_t *n_lines);
int ssd_fastq_record_cuts(const char *path, uint64_t target_bytes, const uint64_t *records_in, uint64_t n_in,
                          uint64_t *records_out, uint64_t *offsets_out, uint64_t capacity, uint64_t *n_out,
                          uint64_t *n_records_total);

/* ---- phase 2: the -m decision, V2.calc_demux_results pass 1 (V2:389-412) ---------------------- */
/* Device pointers to the n^3 uint32 per-cell read counts / distinct-UMI counts of this context.
 * A multi-GPU host all-reduces (sums) them in place between the calls below. */
int ssd_cell_space(const ssd_ctx *ctx, uint64_t *n_cells_total);
int ssd_reads_histogram_device(ssd_ctx *ctx, uint32_t **d_hist);
int ssd_distinct_histogram_device(ssd_ctx *ctx, uint32_t **d_distinct);

/* Sort + deduplicate this batch's (cell,UMI) keys.  Regular keys are uint64 (cell << umi_bits | 2-bit
 * UMI code), irregular keys (UMI shorter than the window or containing non-ACGT) are 16-byte records
 * {uint64 hi = cell<<36 | len<<32 | bytes[0..4), uint64 lo = bytes[4..10)}.  Both come back sorted and
 * unique; the arrays stay owned by the context. */
int ssd_local_unique_keys(ssd_ctx *ctx, uint64_t **d_keys, uint64_t *n_keys, void **d_irregular,
                          uint64_t *n_irregular);
/* Per-cell x UMI table of the last batch (extension, SURVEY 8f.4: the table a downstream counter wants, without a FASTQ round
 * trip): the sorted unique (cell,UMI) keys as above, and next to them the number of READS each key was seen in (uint32).  The
 * arrays live on the device and belong to the context. */
int ssd_cell_umi_table(ssd_ctx *ctx, uint64_t **d_keys, uint32_t **d_reads, uint64_t *n_keys, void **d_irregular,
                       uint32_t **d_irregular_reads, uint64_t *n_irregular);
/* Count distinct keys per cell into the distinct histogram (overwrites it).  Inputs are device arrays
 * in the formats above, not necessarily unique or sorted (e.g. the concatenation of what peers sent
 * after an all-to-all by cell range); they are sorted in place. */
int ssd_count_distinct(ssd_ctx *ctx, uint64_t *d_keys, uint64_t n_keys, void *d_irregular,
                       uint64_t n_irregular);
/* Multi-GPU hosts, the cheaper exchange: this batch's (cell,UMI) keys WITH their duplicates, regular keys grouped by
 * the rank that owns their cell (one partition pass instead of a sort): owner k's keys are
 * (*d_keys)[offsets[k] .. offsets[k + 1]), offsets has world + 1 entries (host).  Irregular keys come back as an
 * unordered list; the caller routes them with ssd_owner_layout.  The arrays stay owned by the context.
 * Replaces the set-union of per-chunk UMI sets the reference gets for free from its single results file (V2:404-411). */
int ssd_partition_keys(ssd_ctx *ctx, int world, uint64_t **d_keys, uint64_t *offsets, void **d_irregular,
                       uint64_t *n_irregular);
/* ssd_count_distinct for a rank of a multi-GPU job: regular keys are what the peers sent for this rank's cells;
 * the irregular keys are everybody's (all-gathered, unordered), of which only those of cells this rank owns are
 * counted; records with hi == ~0 are padding. */
int ssd_count_distinct_owned(ssd_ctx *ctx, uint64_t *d_keys, uint64_t n_keys, void *d_irregular,
                             uint64_t n_irregular, int rank, int world);
/* Multi-GPU hosts, the -m decision without shipping every key (what ssd_b200.distributed.global_filter does):
 *   ssd_distinct_local   per-cell distinct-UMI counts of THIS batch into the distinct table (the read histogram must still be the
 *                        batch's own).
 *   ssd_pack_decision    *d_packed -> n^3 uint64 on the device, one word per cell: bits 0..39 = reads of the cell in this batch,
 *                        bit 40 = this batch alone counts >= m distinct UMIs.  The caller all-reduces (SUM) the array over the
 *                        ranks in place: bits 0..39 = global reads, bits 40.. = number of ranks that decide "pass" alone.  A
 *                        cell passes when that number is not zero and fails when the global reads are < m (distinct <= reads).
 *   ssd_undecided_keys   this batch's (cell,UMI) keys, duplicates kept, of the cells in between (given the all-reduced words) as
 *                        ONE packed list of capacity + 1 records in the irregular-key format above (regular UMIs spelled out as
 *                        bytes): *d_records -> record 0 is the header {hi = ~0, lo = number of keys this batch had}, records
 *                        1 .. capacity are keys, unused ones all ones (padding).  lo > capacity means the list is incomplete.
 *                        Nothing comes back to the host: the lists of all ranks are all-gathered as they are (fixed size), and
 *                        ssd_count_distinct(ctx, NULL, 0, gathered, world * (capacity + 1)) counts them (padding and headers
 *                        are skipped).  Every rank holds fewer than m distinct keys per such cell, so this decides those cells
 *                        exactly.
 *   ssd_adopt_decision   after that count: read histogram := global reads, distinct[cell] := max(distinct[cell], m) for the cells
 *                        some rank decided alone.  The distinct table is then a lower bound of the exact global one that decides
 *                        -m exactly like it, identical on all ranks; ssd_build_filter follows. */
int ssd_distinct_local(ssd_ctx *ctx);
int ssd_pack_decision(ssd_ctx *ctx, uint64_t **d_packed);
int ssd_undecided_keys(ssd_ctx *ctx, const uint64_t *d_packed_global, uint64_t capacity, void **d_records);
int ssd_adopt_decision(ssd_ctx *ctx, const uint64_t *d_packed_global);
/* Owner rank of a cell in a world of W ranks: ((cell >> *shift) * W) / *n_buckets. */
int ssd_owner_layout(const ssd_ctx *ctx, int *shift, int *n_buckets);
/* Bit width of the 2-bit-packed UMI inside a regular key (cell id = key >> umi_bits). */
int ssd_key_layout(const ssd_ctx *ctx, int *umi_bits, int *cell_bits);

/* Decide which cells pass from the (possibly all-reduced) histograms, apply the optional
 * RanHex->OligoDT collapse, size both outputs.  Fills `stats` (may be NULL).  n_cells and
 * n_passing_cells count over the whole (all-reduced) histogram; the record counts are this batch's. */
int ssd_build_filter(ssd_ctx *ctx, ssd_stats *stats);
/* The statistics of the current batch as the last ssd_build_filter left them. */
int ssd_get_stats(const ssd_ctx *ctx, ssd_stats *stats);

/* Single-GPU convenience: ssd_local_unique_keys + ssd_count_distinct (distinct mode only) +
 * ssd_build_filter. */
int ssd_filter_single(ssd_ctx *ctx, ssd_stats *stats);

/* ---- phase 3: the writers, V2:134-141 + V2:367-370 and V2:415-454 ----------------------------- */
/* Write the annotated FASTQ text of this batch, records in input order, into the DEVICE buffer d_out
 * (capacity `cap` bytes, 16-byte aligned).  *written receives stats.bytes_matched / bytes_passing. */
int ssd_emit_device(ssd_ctx *ctx, int which, void *d_out, size_t cap, size_t *written);

/* Split mode (BASELINE config 5; bucket model of demultiplex_using_barcodes.py:350-372, one buffer per cell): the
 * SSD_EMIT_PASSING records of the batch grouped into per-cell buckets — cells in ascending id, records in input
 * order inside a cell — written back to back into d_out.  The union of the buckets is the merged output. */
int ssd_emit_split_device(ssd_ctx *ctx, void *d_out, size_t cap, size_t *written, uint64_t *n_buckets);
/* Table of the last ssd_emit_split_device: per bucket its cell id ((i1*n + i2)*n + i3), first byte and first
 * record position in d_out; arrays of `capacity` >= n_buckets entries on the host (any may be NULL). */
int ssd_split_index(ssd_ctx *ctx, uint32_t *cells_host, uint64_t *offsets_host, uint64_t *first_host, uint64_t capacity);

/* ---- host-buffer convenience (the end-to-end path) -------------------------------------------- */
/* Whole pipeline on HOST buffers: chunked H2D into context-owned device staging, phases 1-3, D2H of
 * the selected output into `out` (capacity out_cap).  Pinned host memory makes the copies async. */
int ssd_run_host(ssd_ctx *ctx, const void *r1, size_t r1_len, const void *r2, size_t r2_len, int which,
                 void *out, size_t out_cap, size_t *written, ssd_stats *stats);
/* After ssd_run_host (or the phase calls): emit the selected output of the current batch into a HOST
 * buffer.  ssd_run_host with out == NULL and out_cap == 0 stops after the filter so the caller can size
 * the buffer from stats->bytes_* and then call this. */
int ssd_emit_host(ssd_ctx *ctx, int which, void *out, size_t out_cap, size_t *written);
/* Size query for ssd_run_host callers: upper bound of the output text for these input sizes. */
size_t ssd_output_bound(const ssd_ctx *ctx, size_t r1_len, size_t r2_len);

/* Pinned host memory helpers (cudaHostAlloc / cudaFreeHost) so that non-CUDA hosts can get async copies. */
int ssd_host_alloc(void **p, size_t bytes);
int ssd_host_free(void *p);

/* ---- per-record results (tests, split mode, downstream hand-off) ------------------------------ */
/* Copy the per-record cell ids of the batch to the host: (i1*n + i2)*n + i3, or -1 when rejected. */
int ssd_record_cells(ssd_ctx *ctx, int32_t *cells_host, uint64_t capacity, uint64_t *n_records);

/* ---- instrumentation -------------------------------------------------------------------------- */
/* Pipeline stages for ssd_read_timings. */
#define SSD_STAGE_SCAN_R1 0   /* record-boundary scan of read 1 + output-length planning              */
#define SSD_STAGE_SCAN_R2 1   /* scan of read 2 fused with extraction, matching, histogram, key capture */
#define SSD_STAGE_DISTINCT 2  /* radix sort + unique + per-cell distinct-UMI count                       */
#define SSD_STAGE_FILTER 3    /* pass decision, collapse map, output sizing                              */
#define SSD_STAGE_PLAN 4      /* exclusive scan of output lengths                                        */
#define SSD_STAGE_EMIT 5      /* the writer                                                              */
#define SSD_N_STAGES 6
/* When enabled, every stage is bracketed by a CUDA-event pair on the context's stream. */
int ssd_set_timing(ssd_ctx *ctx, int enabled);
/* Synchronises the stream, sums the recorded spans per stage (milliseconds, number of spans) and
 * clears them. */
int ssd_read_timings(ssd_ctx *ctx, double *ms_per_stage, uint32_t *spans_per_stage, int n_stages);
/* Number of kernels this context has launched since it was created. */
uint64_t ssd_launch_count(const ssd_ctx *ctx);
/* Diagnostics: how many times a scan had to be repeated without guessing the record phase. */
int ssd_scan_repeats(const ssd_ctx *ctx, uint64_t *n);
/* Diagnostics: which read-2 scan the NEXT batch of this context runs (*short_fast = 1: the variant that keeps short read-2
 * sequences -- V2:296-301 slices past the end, hamming() zips to the shorter operand, V2:72-73 -- and windows with several
 * non-ACGT bytes on the fast path), and how many read-2 records of the LAST batch took the exact record-by-record path.  A
 * context starts without and switches when a batch sent more than 1 record in 1 000 that way (SSD_SCAN_SHORT=0/1 pins it).
 * Results are the same bytes either way.  Either pointer may be NULL. */
int ssd_scan_variant(const ssd_ctx *ctx, int *short_fast, uint64_t *exact_records);

/* ---- synthetic read pairs (bench / tests) ---------------------------------------- */
typedef struct ssd_synth_spec {
    uint32_t struct_size;
    uint64_t seed;
    uint64_t first_index;     /* 1-based global index of the first pair                               */
    uint64_t n_pairs;
    int32_t read_len;         /* bases per mate (150)                                                  */
    int32_t n_whitelist;
    const char *whitelist;    /* n_whitelist x 8                                                       */
    uint32_t cell_pool;       /* number of distinct planted cells                                      */
    uint32_t p_sub_ppm;       /* per-base substitution probability inside barcode windows, ppm        */
    uint32_t p_n_ppm;         /* per-base N probability over read 2, ppm                               */
    uint32_t p_junk_ppm;      /* reads with random barcodes, ppm                                       */
    uint32_t p_dup_umi_ppm;   /* reads that reuse a low-entropy UMI of their cell, ppm                 */
    uint32_t p_tie_ppm;       /* reads with a barcode at the midpoint of a distance-4 whitelist pair   */
    uint32_t p_trunc_ppm;     /* read 2 truncated to 40..93 bases, ppm                                 */
    uint32_t p_atqual_ppm;    /* quality lines starting with '@', ppm                                  */
} ssd_synth_spec;

/* Exact byte sizes of the two files the spec describes. */
int ssd_synth_sizes(const ssd_synth_spec *spec, uint64_t *r1_bytes, uint64_t *r2_bytes);
/* Generate on the device (pointers from the caller, sizes from ssd_synth_sizes). */
int ssd_synth_generate_device(ssd_ctx *ctx, const ssd_synth_spec *spec, void *d_r1, void *d_r2);
/* Same bytes on the host (plain C++ twin of the device generator, for CPU-side tests). */
int ssd_synth_generate_host(const ssd_synth_spec *spec, void *r1, void *r2);

#ifdef __cplusplus
}
#endif
#endif
