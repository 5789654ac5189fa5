// ssd_b200.cu — host side of libssd_b200.so: the C ABI declared in include/ssd_b200.h, device memory
// management, kernel launches and the synthetic read generator.  Everything that touches bytes of the
// FASTQ files runs in the kernels of ssd_kernels.cuh; there is no CPU implementation of the path here.
#include "../../include/ssd_b200.h"

#include <cuda_runtime.h>
#include <stdarg.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <errno.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/vfs.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "ssd_kernels.cuh"
#include "ssd_synth.cuh"

using namespace ssd;

namespace {

thread_local char g_create_error[512] = "";

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

}  // namespace

namespace {
constexpr size_t kIoChunk = (size_t)16 << 20;  // bytes per pinned slot
constexpr int kIoLanes = 16;                   // host threads moving file data (8 per input file; 8 for the output)
constexpr int kIoHalf = kIoLanes / 2;
constexpr int kIoSlots = 2;                    // pinned slots per lane: one being filled / written, one in flight

struct IoLane {  // what one I/O thread owns
    uint8_t* slot[kIoSlots] = {nullptr, nullptr};
    cudaEvent_t done[kIoSlots] = {nullptr, nullptr};
    cudaStream_t stream = nullptr;
    int init()
    {
        if (stream) return SSD_OK;
        if (cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess) return SSD_ECUDA;
        for (int k = 0; k < kIoSlots; k++) {
            if (cudaHostAlloc((void**)&slot[k], kIoChunk, cudaHostAllocDefault) != cudaSuccess) return SSD_ENOMEM;
            if (cudaEventCreateWithFlags(&done[k], cudaEventDisableTiming) != cudaSuccess) return SSD_ECUDA;
        }
        return SSD_OK;
    }
    void destroy()
    {
        for (int k = 0; k < kIoSlots; k++) {
            if (slot[k]) cudaFreeHost(slot[k]);
            if (done[k]) cudaEventDestroy(done[k]);
            slot[k] = nullptr;
            done[k] = nullptr;
        }
        if (stream) cudaStreamDestroy(stream);
        stream = nullptr;
    }
};
}  // namespace

struct ssd_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    ssd_config cfg{};
    DevCfg dcfg{};
    uint64_t n_cells_total = 0;
    int record_hint = 48;
    int bc_const = 0;  // bytes of the three barcodes of a cell when every entry of a round has the same length
    int n_lut_sets = 0;  // distinct whitelists among the three rounds
    uint32_t scan_grid = 296;  // persistent scan blocks: SM count x resident blocks per SM
    char err[512] = "";

    DevCounters* d_cnt = nullptr;
    unsigned long long* d_err_off = nullptr;
    unsigned long long* d_ucnt = nullptr;  // [2] cursor of ssd_undecided_keys
    uint8_t* d_lut = nullptr;
    uint8_t* d_wl_len = nullptr;
    unsigned long long* d_wl64 = nullptr;  // whitelist entries as 8-byte words (tiled writer)
    uint32_t* d_hist = nullptr;
    uint32_t* d_distinct = nullptr;
    uint32_t* d_remap = nullptr;
    uint8_t* d_pass = nullptr;

    // current batch
    const uint8_t* r1 = nullptr;
    const uint8_t* r2 = nullptr;
    size_t r1_len = 0, r2_len = 0;
    bool wide = false;  // 64-bit record offsets
    uint64_t rec_cap = 0;
    uint64_t n_rec_r1 = 0, n_rec_r2 = 0, n_rec = 0, n_keys = 0, n_irr = 0;
    bool defer1 = true, defer2 = true;  // scan variant per file: park-and-commit, or (after a failed guess) wait for the line index
    // read-2 scan variant that keeps short sequences and windows with several non-ACGT bytes on the fast path (k_scan<.., SHORT>):
    // off until a batch sends more than 1 record in 1 000 down the exact path, then on for the next batches of this context
    // (SSD_SCAN_SHORT=1: on from the first batch, =0: never)
    bool short2 = false, short_auto = true;
    uint64_t n_exact_last = 0;  // read-2 records of the last batch that took the exact path
    uint32_t emit_grid = 148;
    uint64_t n_respec = 0;  // scans repeated without guessing (diagnostics)
    bool phase1_done = false, filter_done = false, unique_done = false;
    int planned = -1;  // which output b_out_off currently describes (-1 = none)
    bool annotated = false;  // the batch is an already annotated FASTQ (V2.calc_demux_results)
    uint64_t n_buckets = 0;  // split mode: buckets of the last ssd_emit_split_device
    uint64_t split_bytes = 0;
    uint64_t* uniq_keys = nullptr;
    uint64_t n_uniq = 0;
    IrrKey* uniq_irr = nullptr;
    uint64_t n_uniq_irr = 0;
    ssd_stats stats{};

    DevBuf b_status1, b_status2, b_r1_start, b_r1_meta, b_r1_base, b_r1_tok, b_ann, b_keys, b_keys_alt, b_irr, b_irr_alt, b_perm, b_perm_alt, b_tmp64a,
        b_tmp64b, b_counts, b_offsets, b_bsums, b_pos, b_out_off, b_cells;
    DevBuf b_in1, b_in2, b_out;
    size_t loaded[2] = {0, 0};  // bytes ssd_load_file_range put into b_in1 / b_in2
    IoLane io[kIoLanes];  // pinned slots + streams of the file-to-file entry points (created on first use)
    DevBuf b_first, b_runs;  // ssd_cell_umi_table: where each unique key's run starts, run lengths
    DevBuf b_und, b_und_bits, b_pack;  // multi-GPU -m decision: packed key list, undecided-cell bitmap, packed per-cell words
    DevBuf b_os, b_irr_table;  // one-kernel-per-pass sort: histograms, tickets, tile status; hash set of the irregular keys
    DevBuf b_seg;              // per-cell hash sets of the regular keys: key counts (when not the read histogram) and segment offsets
    DevBuf b_sp_cells, b_sp_cells_alt, b_sp_recs, b_sp_recs_alt, b_sp_off, b_bk_cell, b_bk_off, b_bk_first;
    // the irregular-key path runs on its own stream with its own scratch, next to the regular sort
    cudaStream_t side_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    DevBuf s_counts, s_offsets, s_bsums, s_pos;

    // instrumentation: kernel launch counter and optional per-stage CUDA-event timing
    uint64_t n_launches = 0;
    bool timing = false;
    struct Span {
        int stage;
        cudaEvent_t a, b;
    };
    std::vector<Span> spans;
};

namespace {

int fail(ssd_ctx* c, int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(c ? c->err : g_create_error, 512, fmt, ap);
    va_end(ap);
    return code;
}


struct StageTimer {  // records a CUDA-event pair around a pipeline stage when timing is on
    ssd_ctx* c;
    cudaEvent_t b = nullptr;
    StageTimer(ssd_ctx* ctx, int stage) : c(ctx)
    {
        if (!c->timing) return;
        cudaEvent_t a;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        cudaEventRecord(a, c->stream);
        c->spans.push_back({stage, a, b});
    }
    ~StageTimer()
    {
        if (b) cudaEventRecord(b, c->stream);
    }
};

#define CU(c, call)                                                                                          \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess) return fail((c), SSD_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

int ensure(ssd_ctx* c, DevBuf& b, size_t bytes)
{
    if (bytes <= b.cap) return SSD_OK;
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        e = cudaMalloc(&b.p, bytes);
        want = bytes;
    }
    if (e != cudaSuccess) return fail(c, SSD_ENOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    b.cap = want;
    return SSD_OK;
}

void release(DevBuf& b)
{
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}

#define OKR(x)                     \
    do {                           \
        int rc_ = (x);             \
        if (rc_ != SSD_OK) return rc_; \
    } while (0)

inline uint32_t blocks_for(uint64_t n, uint32_t per) { return (uint32_t)((n + per - 1) / per); }

// ---- device-wide exclusive scan (out has n + 1 entries) ----------------------------------------
template <typename T, class F>
int exclusive_scan(ssd_ctx* c, F f, uint64_t n, T* out)
{
    const uint32_t nb = (uint32_t)(n / kScanBlock + 1);
    OKR(ensure(c, c->b_bsums, (size_t)nb * sizeof(T)));
    T* bs = (T*)c->b_bsums.p;
    k_scan_reduce<T, F><<<nb, kThreads, 0, c->stream>>>(f, n, bs);
    k_scan_block_sums<T><<<1, kThreads, 0, c->stream>>>(bs, nb, (T*)nullptr);
    k_scan_apply<T, F><<<nb, kThreads, 0, c->stream>>>(f, n, bs, out);
    c->n_launches += 3;
    CU(c, cudaGetLastError());
    return SSD_OK;
}

struct LoadU32 {
    const uint32_t* p;
    __device__ __forceinline__ uint32_t operator()(uint64_t i) const { return p[i]; }
};

// ---- LSD radix sort of 64-bit keys over bits [0, bits); returns where the sorted data ended up ----
template <class Src, bool HV>
int radix_pass(ssd_ctx* c, Src src, uint64_t n, int bit, uint64_t* dst, const uint32_t* vsrc, uint32_t* vdst)
{
    const uint32_t nb = blocks_for(n, kSortBlock);
    const size_t ncount = (size_t)kRadix * nb;
    OKR(ensure(c, c->b_counts, ncount * sizeof(uint32_t)));
    OKR(ensure(c, c->b_offsets, (ncount + 1) * sizeof(uint32_t)));
    uint32_t* counts = (uint32_t*)c->b_counts.p;
    uint32_t* offsets = (uint32_t*)c->b_offsets.p;
    k_sort_hist<Src><<<nb, kThreads, 0, c->stream>>>(src, n, bit, counts, nb);
    if (ncount <= kSmallScanMax) {
        k_scan_small_u32<<<1, kSmallScanThreads, 0, c->stream>>>(counts, (uint32_t)ncount, offsets);
        c->n_launches++;
    } else {
        OKR(exclusive_scan<uint32_t>(c, LoadU32{counts}, ncount, offsets));
    }
    k_sort_scatter<Src, uint64_t, HV><<<nb, kThreads, 0, c->stream>>>(src, dst, vsrc, vdst, n, bit, offsets, nb);
    c->n_launches += 2;
    return SSD_OK;
}

template <bool HV>
int radix_sort64(ssd_ctx* c, uint64_t* a, uint64_t* b, uint32_t* va, uint32_t* vb, uint64_t n, int bits, uint64_t** sorted,
                 uint32_t** sorted_vals)
{
    *sorted = a;
    if (sorted_vals) *sorted_vals = va;
    if (n == 0) return SSD_OK;
    for (int bit = 0; bit < bits; bit += kRadixBits) {
        OKR((radix_pass<KeysArray, HV>(c, KeysArray{a}, n, bit, b, va, vb)));
        std::swap(a, b);
        std::swap(va, vb);
    }
    CU(c, cudaGetLastError());
    *sorted = a;
    if (sorted_vals) *sorted_vals = va;
    return SSD_OK;
}

// LSD radix sort of 64-bit keys over bits [0, bits), one kernel per 8-bit digit (ssd_onesweep.cuh).  from_ann: the keys are
// the regular (cell, UMI) keys of the batch's annotation words (n_keys of them), a and b are only buffers; otherwise a holds
// the keys.  *sorted is whichever of a / b the result ended up in.
int os_sort64(ssd_ctx* c, bool from_ann, uint64_t* a, uint64_t* b, uint64_t n_keys, int bits, uint64_t** sorted)
{
    *sorted = a;
    if (n_keys == 0) return SSD_OK;
    const int passes = std::max(1, (bits + kOsBits - 1) / kOsBits);
    if (passes > kOsMaxPasses || n_keys > (uint64_t)kOsValue) return fail(c, SSD_ERANGE, "key space too large for the radix sort");
    const uint64_t n_first = from_ann ? c->n_rec_r2 : n_keys;
    const uint64_t tiles_first = (n_first + kOsTile - 1) / kOsTile, tiles_next = (n_keys + kOsTile - 1) / kOsTile;
    const size_t head_words = (size_t)kOsMaxPasses * kOsRadix + 64;  // histograms, then one ticket per pass
    const size_t status_words = (size_t)(tiles_first + (uint64_t)(passes - 1) * tiles_next) * kOsRadix;
    OKR(ensure(c, c->b_os, (head_words + status_words) * sizeof(uint32_t)));
    uint32_t* ghist = (uint32_t*)c->b_os.p;
    uint32_t* ticket = ghist + (size_t)kOsMaxPasses * kOsRadix;
    uint32_t* status = ghist + head_words;
    CU(c, cudaMemsetAsync(c->b_os.p, 0, (head_words + status_words) * sizeof(uint32_t), c->stream));
    const uint32_t hist_blocks = (uint32_t)std::min<uint64_t>(blocks_for(n_first, kThreads), 148 * 8);
    uint64_t *src = a, *dst = b;
    if (from_ann) {
        const KeysFromAnn ka{(const uint64_t*)c->b_ann.p, c->dcfg.umi_bits};
        k_os_hist<KeysFromAnn><<<hist_blocks, kThreads, 0, c->stream>>>(ka, n_first, passes, ghist, 0);
        k_os_pass<KeysFromAnn><<<(uint32_t)tiles_first, kThreads, 0, c->stream>>>(ka, n_first, 0, a, ghist, status, ticket);
    } else {
        k_os_hist<KeysArray><<<hist_blocks, kThreads, 0, c->stream>>>(KeysArray{a}, n_first, passes, ghist, 0);
        k_os_pass<KeysArray><<<(uint32_t)tiles_first, kThreads, 0, c->stream>>>(KeysArray{a}, n_first, 0, b, ghist, status, ticket);
        std::swap(src, dst);
    }
    status += tiles_first * kOsRadix;
    for (int p = 1; p < passes; p++) {
        k_os_pass<KeysArray><<<(uint32_t)tiles_next, kThreads, 0, c->stream>>>(KeysArray{src}, n_keys, p * kOsBits, dst, ghist + (size_t)p * kOsRadix, status, ticket + p);
        status += tiles_next * kOsRadix;
        std::swap(src, dst);
    }
    c->n_launches += 1 + passes;
    CU(c, cudaGetLastError());
    *sorted = src;
    return SSD_OK;
}

// sort + unique of regular keys.  `keys` holds n keys (or, when from_ann, the keys are taken from the
// annotation words of the batch and `keys` only is the buffer they are compacted into by the first pass);
// result: *uniq (a context buffer or `keys`), *n_uniq
int sort_unique64(ssd_ctx* c, uint64_t* keys, uint64_t n, uint64_t** uniq, uint32_t* total_host, bool from_ann = false, uint32_t* first = nullptr)
{
    // enqueues only: *total_host (pinned or pageable) is valid after the caller synchronised c->stream
    *uniq = keys;
    *total_host = 0;
    if (n == 0) return SSD_OK;
    OKR(ensure(c, c->b_keys_alt, n * sizeof(uint64_t)));
    uint64_t* sorted = nullptr;
    const int bits = c->dcfg.umi_bits + c->dcfg.cell_bits;
    OKR(os_sort64(c, from_ann, keys, (uint64_t*)c->b_keys_alt.p, n, bits, &sorted));
    uint64_t* other = sorted == keys ? (uint64_t*)c->b_keys_alt.p : keys;
    OKR(ensure(c, c->b_pos, (n + 1) * sizeof(uint32_t)));
    uint32_t* pos = (uint32_t*)c->b_pos.p;
    OKR(exclusive_scan<uint32_t>(c, HeadFlag64{sorted}, n, pos));
    k_compact64<<<blocks_for(n, kThreads), kThreads, 0, c->stream>>>(sorted, pos, n, other, first);
    c->n_launches += 1;
    CU(c, cudaMemcpyAsync(total_host, pos + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    *uniq = other;
    return SSD_OK;
}

// sort + unique of irregular keys through an index permutation (128-bit LSD: low word, then high word)
int sort_unique_irr(ssd_ctx* c, IrrKey* keys, uint64_t n, IrrKey** uniq, uint32_t* total_host, uint32_t* first = nullptr)
{
    // enqueues only, like sort_unique64
    *uniq = keys;
    *total_host = 0;
    if (n == 0) return SSD_OK;
    OKR(ensure(c, c->b_perm, n * sizeof(uint32_t)));
    OKR(ensure(c, c->b_perm_alt, n * sizeof(uint32_t)));
    OKR(ensure(c, c->b_tmp64a, n * sizeof(uint64_t)));
    OKR(ensure(c, c->b_tmp64b, n * sizeof(uint64_t)));
    OKR(ensure(c, c->b_irr_alt, n * sizeof(IrrKey)));
    uint32_t* perm = (uint32_t*)c->b_perm.p;
    uint32_t* perm_alt = (uint32_t*)c->b_perm_alt.p;
    uint64_t* ta = (uint64_t*)c->b_tmp64a.p;
    uint64_t* tb = (uint64_t*)c->b_tmp64b.p;
    const uint32_t nb = blocks_for(n, kThreads);
    k_iota<<<nb, kThreads, 0, c->stream>>>(perm, n);
    c->n_launches += 4;  // iota, two gathers, compaction
    uint64_t* sk = nullptr;
    uint32_t* sv = nullptr;
    for (int word = 0; word < 2; word++) {
        k_gather_irr_word<<<nb, kThreads, 0, c->stream>>>(keys, perm, n, word, ta);
        OKR(radix_sort64<true>(c, ta, tb, perm, perm_alt, n, word ? kIrrHiBits : kIrrLoBits, &sk, &sv));
        if (sv != perm) std::swap(perm, perm_alt);
        if (sk != ta) std::swap(ta, tb);
    }
    OKR(ensure(c, c->b_pos, (n + 1) * sizeof(uint32_t)));
    uint32_t* pos = (uint32_t*)c->b_pos.p;
    OKR(exclusive_scan<uint32_t>(c, HeadFlagIrr{keys, perm}, n, pos));
    IrrKey* out = (IrrKey*)c->b_irr_alt.p;
    k_compact_irr<<<nb, kThreads, 0, c->stream>>>(keys, perm, pos, n, out, first);
    CU(c, cudaMemcpyAsync(total_host, pos + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    *uniq = out;
    return SSD_OK;
}

// Regular keys on the context's stream, irregular keys (few, launch-latency bound) next to them on the side
// stream with their own scratch; joined before returning.  from_ann: keys come from the batch's annotations.
int sort_unique_both(ssd_ctx* c, uint64_t* keys, uint64_t n_keys, IrrKey* irr, uint64_t n_irr, bool from_ann, uint64_t** uk, uint64_t* nuk,
                     IrrKey** ui, uint64_t* nui, uint32_t* first_k = nullptr, uint32_t* first_i = nullptr)
{
    uint32_t tot_k = 0, tot_i = 0;
    const bool side = n_irr != 0 && c->side_stream != nullptr;
    if (side) {
        CU(c, cudaEventRecord(c->ev_fork, c->stream));
        CU(c, cudaStreamWaitEvent(c->side_stream, c->ev_fork, 0));
        std::swap(c->stream, c->side_stream);
        std::swap(c->b_counts, c->s_counts);
        std::swap(c->b_offsets, c->s_offsets);
        std::swap(c->b_bsums, c->s_bsums);
        std::swap(c->b_pos, c->s_pos);
    }
    int rc = SSD_OK;
    if (n_irr) {
        if (from_ann) {
            unsigned long long* cursor = (unsigned long long*)c->d_err_off;  // free after phase 1 (errors were read)
            cudaMemsetAsync(cursor, 0, sizeof(unsigned long long), c->stream);
            k_collect_irr<<<blocks_for(c->n_rec_r2, kThreads), kThreads, 0, c->stream>>>((const uint64_t*)c->b_ann.p, c->n_rec_r2, c->r2, irr, cursor);
            c->n_launches++;
        }
        rc = sort_unique_irr(c, irr, n_irr, ui, &tot_i, first_i);
    } else {
        *ui = irr;
    }
    if (side) {
        cudaEventRecord(c->ev_join, c->stream);
        std::swap(c->stream, c->side_stream);
        std::swap(c->b_counts, c->s_counts);
        std::swap(c->b_offsets, c->s_offsets);
        std::swap(c->b_bsums, c->s_bsums);
        std::swap(c->b_pos, c->s_pos);
    }
    if (rc == SSD_OK) rc = sort_unique64(c, keys, n_keys, uk, &tot_k, from_ann, first_k);
    if (side) cudaStreamWaitEvent(c->stream, c->ev_join, 0);
    if (rc != SSD_OK) return rc;
    CU(c, cudaStreamSynchronize(c->stream));
    *nuk = tot_k;
    *nui = tot_i;
    return SSD_OK;
}

int count_cells(ssd_ctx* c, const uint64_t* uniq, uint64_t n_uniq, const IrrKey* uirr, uint64_t n_uirr)
{
    CU(c, cudaMemsetAsync(c->d_distinct, 0, c->n_cells_total * sizeof(uint32_t), c->stream));
    if (n_uniq) k_count_cells64<<<blocks_for(n_uniq, kThreads), kThreads, 0, c->stream>>>(uniq, n_uniq, c->dcfg.umi_bits, c->d_distinct);
    if (n_uirr) k_count_cells_irr<<<blocks_for(n_uirr, kThreads), kThreads, 0, c->stream>>>(uirr, n_uirr, c->d_distinct);
    c->n_launches += (n_uniq ? 1 : 0) + (n_uirr ? 1 : 0);
    CU(c, cudaGetLastError());
    return SSD_OK;
}

// Per-cell distinct-UMI counts of the batch straight from the annotation words (single GPU: no unique-key lists, no
// host round trip).  Regular keys: one-kernel-per-pass radix sort + first-occurrence count; irregular keys: hash set
// on the side stream.
// hash set of irregular keys on `st`: from the batch's annotation words (keys == nullptr) or from an array of n keys
void owner_layout(const ssd_ctx* c, int* shift, int* n_buckets);

int irr_hash_counts(ssd_ctx* c, const IrrKey* keys, uint64_t n, cudaStream_t st, int rank = 0, int world = 0)
{
    if (n == 0) return SSD_OK;
    uint64_t slots = 1024;
    while (slots < 2 * n) slots <<= 1;
    if (slots > (1ull << 32)) return fail(c, SSD_ERANGE, "too many irregular UMIs in one batch");
    OKR(ensure(c, c->b_irr_table, slots * sizeof(ulonglong2)));
    CU(c, cudaMemsetAsync(c->b_irr_table.p, 0, slots * sizeof(ulonglong2), st));
    int shift = 0, nb = 1;
    owner_layout(c, &shift, &nb);
    if (keys) k_irr_hash_keys<<<blocks_for(n, kThreads), kThreads, 0, st>>>(keys, n, (ulonglong2*)c->b_irr_table.p, (uint32_t)(slots - 1), c->d_distinct, rank, world,
                                                                          shift, nb);
    else k_irr_hash<<<blocks_for(c->n_rec_r2, kThreads), kThreads, 0, st>>>((const uint64_t*)c->b_ann.p, c->n_rec_r2, c->r2, (ulonglong2*)c->b_irr_table.p,
                                                                              (uint32_t)(slots - 1), c->d_distinct);
    c->n_launches++;
    CU(c, cudaGetLastError());
    return SSD_OK;
}

// Regular keys into per-cell hash sets (k_umi_sets): distinct[cell] += first occurrences.  `counts` = per-cell number of
// keys (an upper bound will do) or nullptr to count the keys here; `table` is scratch for 2 x n_keys_bound 32-bit slots.
inline bool use_sort_for_distinct(uint64_t n_keys_bound)
{
    static const bool v = getenv("SSD_DISTINCT_SORT") != nullptr;
    return v || n_keys_bound >= (1ull << 31);  // segment offsets are 32-bit
}

template <class Src>
int umi_set_counts(ssd_ctx* c, Src src, uint64_t n_src, uint64_t n_keys_bound, const uint32_t* counts, DevBuf* table)
{
    const uint64_t nc = c->n_cells_total;
    OKR(ensure(c, c->b_seg, (2 * nc + 2) * sizeof(uint32_t)));
    uint32_t* seg = (uint32_t*)c->b_seg.p;
    if (!counts) {
        uint32_t* own = seg + nc + 1;
        CU(c, cudaMemsetAsync(own, 0, nc * sizeof(uint32_t), c->stream));
        k_cell_key_counts<Src><<<(uint32_t)std::min<uint64_t>(blocks_for(n_src, kThreads), 148 * 8), kThreads, 0, c->stream>>>(src, n_src, c->dcfg.umi_bits, own);
        c->n_launches++;
        counts = own;
    }
    const uint32_t bitmap_words = c->dcfg.umi_bits >= 5 ? 1u << (c->dcfg.umi_bits - 5) : 1u;
    OKR(exclusive_scan<uint32_t>(c, UmiSetWords{counts, bitmap_words}, nc, seg));
    OKR(ensure(c, *table, 2 * n_keys_bound * sizeof(uint32_t)));
    CU(c, cudaMemsetAsync(table->p, 0, 2 * n_keys_bound * sizeof(uint32_t), c->stream));
    k_umi_sets<Src><<<(uint32_t)std::min<uint64_t>(blocks_for(n_src, kThreads), 148 * 8), kThreads, 0, c->stream>>>(src, n_src, c->dcfg.umi_bits, counts, seg, (uint32_t*)table->p, c->d_distinct);
    k_umi_bitmap_counts<<<148 * 4, kThreads, 0, c->stream>>>(counts, seg, (const uint32_t*)table->p, (uint32_t)nc, c->dcfg.umi_bits, c->d_distinct);
    c->n_launches += 2;
    CU(c, cudaGetLastError());
    return SSD_OK;
}

// Per-cell distinct counts of arbitrary key arrays (duplicates allowed; d_keys is reordered): what a rank does with the
// keys its peers sent for the cells it owns.
int distinct_counts_arrays(ssd_ctx* c, uint64_t* d_keys, uint64_t n_keys, const IrrKey* d_irr, uint64_t n_irr, int rank = 0, int world = 0)
{
    CU(c, cudaMemsetAsync(c->d_distinct, 0, c->n_cells_total * sizeof(uint32_t), c->stream));
    const bool side = n_irr != 0 && c->side_stream != nullptr;
    if (n_irr) {
        cudaStream_t st = c->stream;
        if (side) {
            CU(c, cudaEventRecord(c->ev_fork, c->stream));
            CU(c, cudaStreamWaitEvent(c->side_stream, c->ev_fork, 0));
            st = c->side_stream;
        }
        OKR(irr_hash_counts(c, d_irr, n_irr, st, rank, world));
        if (side) CU(c, cudaEventRecord(c->ev_join, st));
    }
    if (n_keys && use_sort_for_distinct(n_keys)) {
        OKR(ensure(c, c->b_keys_alt, n_keys * sizeof(uint64_t)));
        uint64_t* sorted = nullptr;
        OKR(os_sort64(c, false, d_keys, (uint64_t*)c->b_keys_alt.p, n_keys, c->dcfg.umi_bits + c->dcfg.cell_bits, &sorted));
        k_count_sorted<<<blocks_for(n_keys, kThreads), kThreads, 0, c->stream>>>(sorted, n_keys, c->dcfg.umi_bits, c->d_distinct);
        c->n_launches++;
    } else if (n_keys) {
        // scratch for the sets: whichever key buffer of the context the caller's keys do not live in
        const char* kp = (const char*)d_keys;
        const bool in_alt = c->b_keys_alt.p && kp >= (const char*)c->b_keys_alt.p && kp < (const char*)c->b_keys_alt.p + c->b_keys_alt.cap;
        OKR(umi_set_counts(c, KeysArray{d_keys}, n_keys, n_keys, nullptr, in_alt ? &c->b_keys : &c->b_keys_alt));
    }
    if (side) CU(c, cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    CU(c, cudaGetLastError());
    return SSD_OK;
}

int distinct_counts_local(ssd_ctx* c)
{
    CU(c, cudaMemsetAsync(c->d_distinct, 0, c->n_cells_total * sizeof(uint32_t), c->stream));
    const uint64_t n_keys = c->n_keys, n_irr = c->n_irr;
    const bool side = n_irr != 0 && c->side_stream != nullptr;
    if (n_irr) {
        cudaStream_t st = c->stream;
        if (side) {
            CU(c, cudaEventRecord(c->ev_fork, c->stream));
            CU(c, cudaStreamWaitEvent(c->side_stream, c->ev_fork, 0));
            st = c->side_stream;
        }
        OKR(irr_hash_counts(c, nullptr, n_irr, st));
        if (side) CU(c, cudaEventRecord(c->ev_join, st));
    }
    if (n_keys && !use_sort_for_distinct(n_keys + n_irr)) {
        // the read histogram of the batch is the per-cell key count (irregular keys included: those sets are a little larger)
        OKR(umi_set_counts(c, KeysFromAnnStream{(const uint64_t*)c->b_ann.p, c->dcfg.umi_bits}, c->n_rec_r2, n_keys + n_irr, c->d_hist, &c->b_keys));
    } else if (n_keys) {
        OKR(ensure(c, c->b_keys, n_keys * sizeof(uint64_t)));
        OKR(ensure(c, c->b_keys_alt, n_keys * sizeof(uint64_t)));
        uint64_t* sorted = nullptr;
        OKR(os_sort64(c, true, (uint64_t*)c->b_keys.p, (uint64_t*)c->b_keys_alt.p, n_keys, c->dcfg.umi_bits + c->dcfg.cell_bits, &sorted));
        k_count_sorted<<<blocks_for(n_keys, kThreads), kThreads, 0, c->stream>>>(sorted, n_keys, c->dcfg.umi_bits, c->d_distinct);
        c->n_launches++;
    }
    if (side) CU(c, cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    CU(c, cudaGetLastError());
    return SSD_OK;
}

// The scan kernel for a file: the one table the launches AND the shared-memory attributes are taken from.
template <typename OffT>
using ScanFn = void (*)(const ScanArgs<OffT>, const DevCfg);

template <typename OffT>
ScanFn<OffT> scan_kernel(int mode, bool defer, bool shrt)
{
    if (mode == kScanR1) return defer ? k_scan<kScanR1, OffT, true, false> : k_scan<kScanR1, OffT, false, false>;
    if (mode == kScanAnn) return defer ? k_scan<kScanAnn, OffT, true, false> : k_scan<kScanAnn, OffT, false, false>;
    if (shrt) return defer ? k_scan<kScanR2, OffT, true, true> : k_scan<kScanR2, OffT, false, true>;
    return defer ? k_scan<kScanR2, OffT, true, false> : k_scan<kScanR2, OffT, false, false>;
}

template <typename OffT>
cudaError_t scan_kernel_attributes()
{
    for (int mode : {kScanR1, kScanR2, kScanAnn})
        for (int defer = 0; defer < 2; defer++)
            for (int shrt = 0; shrt < 2; shrt++) {
                cudaError_t e = cudaFuncSetAttribute((const void*)scan_kernel<OffT>(mode, defer != 0, shrt != 0), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     (int)sizeof(ScanSmem));
                if (e != cudaSuccess) return e;
            }
    return cudaSuccess;
}

template <typename OffT>
int launch_scans(ssd_ctx* c, bool defer1, bool defer2)
{
    ScanArgs<OffT> a{};
    a.cnt = c->d_cnt;
    a.err_off = c->d_err_off;
    a.rec_cap = c->rec_cap;
    a.r1_start = (OffT*)c->b_r1_start.p;
    a.r1_meta = (uint4*)c->b_r1_meta.p;
    a.r1_base = (uint32_t*)c->b_r1_base.p;
    a.r1_tok = (unsigned long long*)c->b_r1_tok.p;
    a.r1buf = c->r1;
    a.r1_len = c->r1_len;
    a.ann = (uint64_t*)c->b_ann.p;
    a.hist = c->d_hist;
    a.irr = (IrrKey*)c->b_irr.p;
    a.lut = c->d_lut;

    const uint32_t nt1 = blocks_for(c->r1_len, kTile), nt2 = c->annotated ? 0u : blocks_for(c->r2_len, kTile);
    if (c->annotated) {
        if (nt1) {
            a.buf = c->r1;
            a.len = c->r1_len;
            a.status = (unsigned long long*)c->b_status1.p;
            a.n_tiles = nt1;
            StageTimer t(c, 0);
            scan_kernel<OffT>(kScanAnn, defer1, false)<<<std::min<uint32_t>(nt1, c->scan_grid), kScanThreads, sizeof(ScanSmem), c->stream>>>(a, c->dcfg);
            c->n_launches++;
        }
        CU(c, cudaGetLastError());
        return SSD_OK;
    }
    if (nt1) {
        a.buf = c->r1;
        a.len = c->r1_len;
        a.status = (unsigned long long*)c->b_status1.p;
        a.n_tiles = nt1;
        StageTimer t(c, 0);
        scan_kernel<OffT>(kScanR1, defer1, false)<<<std::min<uint32_t>(nt1, c->scan_grid), kScanThreads, sizeof(ScanSmem), c->stream>>>(a, c->dcfg);
        c->n_launches++;
    }
    if (nt2) {
        a.buf = c->r2;
        a.len = c->r2_len;
        a.status = (unsigned long long*)c->b_status2.p;
        a.n_tiles = nt2;
        StageTimer t(c, 1);
        scan_kernel<OffT>(kScanR2, defer2, c->short2)<<<std::min<uint32_t>(nt2, c->scan_grid), kScanThreads, sizeof(ScanSmem), c->stream>>>(a, c->dcfg);
        c->n_launches++;
    }
    CU(c, cudaGetLastError());
    return SSD_OK;
}

template <typename OffT>
int launch_emit(ssd_ctx* c, int which, uint8_t* d_out, const uint32_t* order = nullptr, const unsigned long long* out_off = nullptr,
                uint64_t n_positions = 0)
{
    EmitArgs<OffT> a{};
    a.order = order;
    a.r1buf = c->r1;
    a.r1_len = c->r1_len;
    a.r2buf = c->r2;
    a.r2_len = c->r2_len;
    a.r1_start = (const OffT*)c->b_r1_start.p;
    a.r1_meta = (const uint4*)c->b_r1_meta.p;
    a.ann = (const uint64_t*)c->b_ann.p;
    a.out_off = out_off ? (const uint64_t*)out_off : (const uint64_t*)c->b_out_off.p;
    a.remap = (which == SSD_EMIT_PASSING && c->cfg.collapse_pairs > 0) ? c->d_remap : nullptr;
    a.out = d_out;
    a.n_rec = order ? n_positions : c->n_rec;
    a.cnt = c->d_cnt;
    a.copy_mode = c->annotated ? 1 : 0;
    // records per block: the block's slice of the output must fit the staging buffer
    const uint64_t n_out = which == SSD_EMIT_MATCHED ? c->stats.n_matched : c->stats.n_retained;
    const uint64_t bytes = which == SSD_EMIT_MATCHED ? c->stats.bytes_matched : c->stats.bytes_passing;
    const uint64_t avg_out = n_out ? bytes / n_out + 1 : 1;
    uint64_t rpb = (uint64_t)(kEmitOut * 7 / 8) / avg_out;
    if (!order && n_out) rpb = rpb * c->n_rec / n_out;  // input-order blocks also walk over the records that are not written
    a.rpb = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(rpb, 1), 256);
    // input order, plain records, 8-byte whitelist entries: the tiled writer (bulk copies in and out, shared-to-shared assembly)
    const bool tiled = !order && !a.copy_mode && c->dcfg.bc8 && c->dcfg.header_style == 0 && c->n_rec && !getenv("SSD_EMIT_GROUPS");
    if (tiled) {
        const uint64_t avg_in = c->r1_len / c->n_rec + 1, avg_out_in = bytes / c->n_rec + 1;  // per input record
        const uint64_t rin = (uint64_t)(kTIn * 7 / 8) / avg_in, rout = (uint64_t)(kTOut * 7 / 8) / avg_out_in;
        a.rpb = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(std::min(rin, rout), 1), kTR);
        a.wl64 = c->d_wl64;
        a.cps = (uint32_t)std::min<uint64_t>(std::max<uint64_t>((avg_out > 40 ? avg_out - 40 : 0) / 16, 1), 64);
        a.cps_magic = ((1u << 20) + a.cps - 1) / a.cps;
    }
    const uint32_t nb = blocks_for(a.n_rec, a.rpb);
    StageTimer t(c, 5);
    if (nb && tiled && !getenv("SSD_EMIT_TILED_BLOCKS"))
        k_emit_stream<OffT><<<std::min<uint32_t>(nb, c->emit_grid), kStreamThreads, sizeof(StreamSmem), c->stream>>>(a, c->dcfg, nb);
    else if (nb && tiled) k_emit_tiled<OffT><<<nb, kThreads, sizeof(TileSmem), c->stream>>>(a, c->dcfg);
    else if (nb) k_emit<OffT><<<nb, kThreads, 0, c->stream>>>(a, c->dcfg);
    c->n_launches += nb ? 1 : 0;
    CU(c, cudaGetLastError());
    return SSD_OK;
}

OutLen make_outlen(ssd_ctx* c, int which)
{
    OutLen f{};
    f.ann = (const uint64_t*)c->b_ann.p;
    f.r1_base = (const uint32_t*)c->b_r1_base.p;
    f.pass = which == SSD_EMIT_PASSING ? c->d_pass : nullptr;
    f.remap = (which == SSD_EMIT_PASSING && c->cfg.collapse_pairs > 0) ? c->d_remap : nullptr;
    f.wl_len = c->d_wl_len;
    f.n2 = c->dcfg.n_wl[1];
    f.n3 = c->dcfg.n_wl[2];
    f.umi_len = c->dcfg.umi_len;
    f.bc_const = c->bc_const;
    f.copy_mode = c->annotated ? 1 : 0;
    return f;
}

int fetch_counters(ssd_ctx* c, DevCounters* h)
{
    CU(c, cudaMemcpyAsync(h, c->d_cnt, sizeof(DevCounters), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return SSD_OK;
}

std::string token_at(ssd_ctx* c, const uint8_t* d_buf, size_t len, unsigned long long off)
{
    char tmp[257];
    size_t n = std::min<size_t>(256, len > off ? len - off : 0);
    if (!n || cudaMemcpy(tmp, d_buf + off, n, cudaMemcpyDeviceToHost) != cudaSuccess) return "?";
    size_t k = 0;
    while (k < n && !((unsigned char)tmp[k] >= 9 && (unsigned char)tmp[k] <= 13) && !((unsigned char)tmp[k] >= 28 && (unsigned char)tmp[k] <= 32)) k++;
    (void)c;
    return std::string(tmp, k);
}

}  // namespace

// ================================================================================================
// lifetime
// ================================================================================================
extern "C" int ssd_abi_version(void) { return SSD_ABI_VERSION; }

extern "C" const char* ssd_last_error(const ssd_ctx* ctx) { return ctx ? ctx->err : g_create_error; }

extern "C" int ssd_create(const ssd_config* cfg, ssd_ctx** out)
{
    if (!cfg || !out) return fail(nullptr, SSD_EINVAL, "ssd_create: null argument");
    *out = nullptr;
    // the structure grew by the per-round lists: a caller compiled against the first layout passes the shorter size and gets one list
    const size_t v1_size = offsetof(ssd_config, n_whitelist2);
    if (cfg->struct_size != sizeof(ssd_config) && cfg->struct_size != v1_size)
        return fail(nullptr, SSD_EINVAL, "ssd_config.struct_size %u is neither %zu nor %zu", cfg->struct_size, sizeof(ssd_config), v1_size);
    ssd_config full;
    memset(&full, 0, sizeof full);
    memcpy(&full, cfg, cfg->struct_size);
    cfg = &full;
    if (cfg->n_whitelist < 1 || cfg->n_whitelist > SSD_MAX_WHITELIST || !cfg->whitelist)
        return fail(nullptr, SSD_EINVAL, "whitelist must hold 1..%d entries", SSD_MAX_WHITELIST);
    if ((cfg->whitelist2 && (cfg->n_whitelist2 < 1 || cfg->n_whitelist2 > SSD_MAX_WHITELIST)) ||
        (cfg->whitelist3 && (cfg->n_whitelist3 < 1 || cfg->n_whitelist3 > SSD_MAX_WHITELIST)))
        return fail(nullptr, SSD_EINVAL, "the round 2 / round 3 whitelists must hold 1..%d entries", SSD_MAX_WHITELIST);
    const int32_t bounds[8] = {cfg->umi_start, cfg->umi_end, cfg->bc1_start, cfg->bc1_end, cfg->bc2_start, cfg->bc2_end, cfg->bc3_start, cfg->bc3_end};
    for (int i = 0; i < 8; i += 2)
        if (bounds[i] < 0 || bounds[i + 1] < bounds[i])
            return fail(nullptr, SSD_EINVAL, "slice bounds must satisfy 0 <= start <= end (got %d:%d)", bounds[i], bounds[i + 1]);
    if (cfg->umi_end - cfg->umi_start > 10) return fail(nullptr, SSD_EINVAL, "UMI windows longer than 10 bases are not supported");
    for (int i = 2; i < 8; i += 2)
        if (bounds[i + 1] - bounds[i] > 8) return fail(nullptr, SSD_EINVAL, "barcode windows longer than 8 bases are not supported");
    if (cfg->errors < 0 || cfg->errors > 8) return fail(nullptr, SSD_EINVAL, "errors must be in 0..8");
    if (cfg->filter_mode != SSD_FILTER_DISTINCT_UMI && cfg->filter_mode != SSD_FILTER_READS) return fail(nullptr, SSD_EINVAL, "unknown filter_mode");
    if (cfg->collapse_pairs < 0 || 2 * cfg->collapse_pairs > cfg->n_whitelist) return fail(nullptr, SSD_EINVAL, "collapse_pairs out of range");  // round-1 list
    if (cfg->header_style != SSD_HEADER_FAST && cfg->header_style != SSD_HEADER_EXTRACT) return fail(nullptr, SSD_EINVAL, "unknown header_style");

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, SSD_ECUDA, "no usable CUDA device (%s); libssd_b200 has no CPU fallback", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));

    ssd_ctx* c = new ssd_ctx();
    c->cfg = *cfg;
    c->cfg.struct_size = sizeof(ssd_config);
    c->cfg.whitelist = c->cfg.whitelist2 = c->cfg.whitelist3 = nullptr;
    c->cfg.whitelist_len = c->cfg.whitelist2_len = c->cfg.whitelist3_len = nullptr;
    if (cfg->device >= 0) {
        c->device = cfg->device;
    } else if (cudaGetDevice(&c->device) != cudaSuccess) {
        c->device = 0;
    }
    auto bail = [&](int code) {
        strncpy(g_create_error, c->err, sizeof g_create_error - 1);
        ssd_destroy(c);
        return code;
    };
#define CUB_(call)                                                                                   \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            fail(c, SSD_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e_));                      \
            return bail(SSD_ECUDA);                                                                  \
        }                                                                                            \
    } while (0)
    CUB_(cudaSetDevice(c->device));
    CUB_(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    CUB_(cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking));
    CUB_(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CUB_(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));

    DevCfg& d = c->dcfg;
    memset(&d, 0, sizeof d);
    d.e = cfg->errors;
    d.us = cfg->umi_start;
    d.ue = cfg->umi_end;
    d.b1s = cfg->bc1_start;
    d.b1e = cfg->bc1_end;
    d.b2s = cfg->bc2_start;
    d.b2e = cfg->bc2_end;
    d.b3s = cfg->bc3_start;
    d.b3e = cfg->bc3_end;
    d.umi_len = d.ue - d.us;
    d.umi_bits = 2 * d.umi_len;
    d.header_style = cfg->header_style;
    // round k: its own list when given, else the first one (V2:60: one list for all rounds)
    const char* lists[3] = {cfg->whitelist, cfg->whitelist2 ? cfg->whitelist2 : cfg->whitelist, cfg->whitelist3 ? cfg->whitelist3 : cfg->whitelist};
    const uint8_t* lens[3] = {cfg->whitelist_len, cfg->whitelist2 ? cfg->whitelist2_len : cfg->whitelist_len,
                              cfg->whitelist3 ? cfg->whitelist3_len : cfg->whitelist_len};
    const int counts[3] = {cfg->n_whitelist, cfg->whitelist2 ? cfg->n_whitelist2 : cfg->n_whitelist, cfg->whitelist3 ? cfg->n_whitelist3 : cfg->n_whitelist};
    d.fast_wl = 1;
    d.bc8 = 1;
    c->bc_const = 0;
    bool same_len = true;
    for (int k = 0; k < 3; k++) {
        d.n_wl[k] = counts[k];
        for (int i = 0; i < d.n_wl[k]; i++) {
            int len = lens[k] ? std::min<int>(lens[k][i], 8) : 8;
            d.wl_len[k][i] = (uint8_t)len;
            uint32_t code = 0;
            for (int j = 0; j < 8; j++) {
                unsigned char ch = j < len ? (unsigned char)lists[k][i * 8 + j] : 0;
                d.wl[k][i * 8 + j] = ch;
                bool acgt = ch == 'A' || ch == 'C' || ch == 'G' || ch == 'T';
                if (!acgt) d.fast_wl = 0;
                code |= ((uint32_t)(ch >> 1) & 3u) << (2 * j);
            }
            if (len != 8) d.fast_wl = d.bc8 = 0;
            if (len != d.wl_len[k][0]) same_len = false;
            d.wl_code[k][i] = (uint16_t)code;
        }
        // rounds with the same list share their lookup tables (sets are numbered densely in order of first use)
        d.lut_set[k] = -1;
        for (int j = 0; j < k && d.lut_set[k] < 0; j++)
            if (d.n_wl[j] == d.n_wl[k] && memcmp(d.wl[j], d.wl[k], sizeof d.wl[k]) == 0 && memcmp(d.wl_len[j], d.wl_len[k], sizeof d.wl_len[k]) == 0)
                d.lut_set[k] = d.lut_set[j];
        if (d.lut_set[k] < 0) d.lut_set[k] = c->n_lut_sets++;
    }
    c->n_cells_total = (uint64_t)d.n_wl[0] * d.n_wl[1] * d.n_wl[2];
    d.cell_bits = 1;
    while ((1ull << d.cell_bits) < c->n_cells_total) d.cell_bits++;
    d.fast_geom = (d.b1e - d.b1s == 8 && d.b2e - d.b2s == 8 && d.b3e - d.b3s == 8) ? 1 : 0;
    d.need_len = std::max(std::max(d.ue, d.b1e), std::max(d.b2e, d.b3e));
    const unsigned long long n3 = (unsigned long long)d.n_wl[2], n23 = (unsigned long long)d.n_wl[1] * n3;
    d.magic_n = ((1ull << 43) + n3 - 1) / n3;
    d.magic_nn = ((1ull << 43) + n23 - 1) / n23;
    if (same_len) c->bc_const = d.wl_len[0][0] + d.wl_len[1][0] + d.wl_len[2][0];
    c->record_hint = cfg->record_bytes_hint > 0 ? cfg->record_bytes_hint : 48;

    {
        // the scan kernels are persistent: one wave of blocks, each looping over tiles
        cudaDeviceProp prop;
        CUB_(cudaGetDeviceProperties(&prop, c->device));
        CUB_(scan_kernel_attributes<uint32_t>());  // every kernel scan_kernel() can return
        CUB_(scan_kernel_attributes<uint64_t>());
        CUB_(cudaFuncSetAttribute(k_emit_tiled<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileSmem)));
        CUB_(cudaFuncSetAttribute(k_emit_tiled<uint64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileSmem)));
        CUB_(cudaFuncSetAttribute(k_emit_stream<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StreamSmem)));
        CUB_(cudaFuncSetAttribute(k_emit_stream<uint64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StreamSmem)));
        int per_sm = 0;
        CUB_(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_emit_stream<uint32_t>, kStreamThreads, sizeof(StreamSmem)));
        c->emit_grid = (uint32_t)prop.multiProcessorCount * (uint32_t)std::max(per_sm, 1);
        CUB_(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, scan_kernel<uint32_t>(kScanR2, true, false), kScanThreads, sizeof(ScanSmem)));
        c->scan_grid = (uint32_t)prop.multiProcessorCount * (uint32_t)std::max(per_sm, 1);
    }
    CUB_(cudaMalloc(&c->d_cnt, sizeof(DevCounters)));
    CUB_(cudaMalloc(&c->d_err_off, 2 * sizeof(unsigned long long)));
    CUB_(cudaMalloc(&c->d_ucnt, 2 * sizeof(unsigned long long)));
    CUB_(cudaMalloc(&c->d_lut, (size_t)c->n_lut_sets * (size_t)(d.e + 1) * 65536u));
    CUB_(cudaMalloc(&c->d_wl_len, 3 * kMaxWl));
    CUB_(cudaMalloc(&c->d_wl64, 3 * kMaxWl * 8));
    CUB_(cudaMalloc(&c->d_hist, c->n_cells_total * sizeof(uint32_t)));
    CUB_(cudaMalloc(&c->d_distinct, c->n_cells_total * sizeof(uint32_t)));
    CUB_(cudaMalloc(&c->d_remap, c->n_cells_total * sizeof(uint32_t)));
    CUB_(cudaMalloc(&c->d_pass, c->n_cells_total));
    CUB_(cudaMemcpyAsync(c->d_wl_len, d.wl_len, 3 * kMaxWl, cudaMemcpyHostToDevice, c->stream));
    CUB_(cudaMemcpyAsync(c->d_wl64, d.wl, 3 * kMaxWl * 8, cudaMemcpyHostToDevice, c->stream));
    CUB_(cudaMemsetAsync(c->d_hist, 0, c->n_cells_total * sizeof(uint32_t), c->stream));
    CUB_(cudaMemsetAsync(c->d_distinct, 0, c->n_cells_total * sizeof(uint32_t), c->stream));
    CUB_(cudaMemsetAsync(c->d_pass, 0, c->n_cells_total, c->stream));
    CUB_(cudaMemsetAsync(c->d_cnt, 0, sizeof(DevCounters), c->stream));
    if (d.fast_wl)
        for (int k = 0, built = 0; k < 3; k++)
            if (d.lut_set[k] == built) {  // the first round of every set builds it
                k_build_lut<<<256, 256, 0, c->stream>>>(d, k, c->d_lut + (size_t)built * (size_t)(d.e + 1) * 65536u, d.e + 1);
                built++;
            }
    CUB_(cudaGetLastError());
    CUB_(cudaStreamSynchronize(c->stream));
#undef CUB_
    *out = c;
    return SSD_OK;
}

extern "C" void ssd_destroy(ssd_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    DevBuf* bufs[] = {&c->b_status1, &c->b_status2, &c->b_r1_start, &c->b_r1_meta, &c->b_r1_base, &c->b_r1_tok, &c->b_ann, &c->b_keys, &c->b_keys_alt, &c->b_irr,
                      &c->b_irr_alt, &c->b_perm, &c->b_perm_alt, &c->b_tmp64a, &c->b_tmp64b, &c->b_counts, &c->b_offsets, &c->b_bsums,
                      &c->b_pos, &c->b_out_off, &c->b_cells, &c->b_in1, &c->b_in2, &c->b_out, &c->b_sp_cells, &c->b_sp_cells_alt, &c->b_sp_recs,
                      &c->b_sp_recs_alt, &c->b_sp_off, &c->b_bk_cell, &c->b_bk_off, &c->b_bk_first, &c->b_os, &c->b_irr_table, &c->b_seg, &c->b_und, &c->b_und_bits, &c->b_pack, &c->b_first, &c->b_runs};
    for (DevBuf* b : bufs) release(*b);
    cudaFree(c->d_cnt);
    cudaFree(c->d_err_off);
    cudaFree(c->d_ucnt);
    cudaFree(c->d_lut);
    cudaFree(c->d_wl_len);
    cudaFree(c->d_wl64);
    cudaFree(c->d_hist);
    cudaFree(c->d_distinct);
    cudaFree(c->d_remap);
    cudaFree(c->d_pass);
    release(c->s_counts);
    release(c->s_offsets);
    release(c->s_bsums);
    release(c->s_pos);
    for (int k = 0; k < kIoLanes; k++) c->io[k].destroy();
    if (c->side_stream) cudaStreamDestroy(c->side_stream);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

extern "C" int ssd_set_stream(ssd_ctx* c, void* cuda_stream)
{
    if (!c) return SSD_EINVAL;
    c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
    return SSD_OK;
}

// ================================================================================================
// phase 1
// ================================================================================================
namespace {

int run_phase1(ssd_ctx* c, const void* d_r1, size_t r1_len, const void* d_r2, size_t r2_len, bool annotated)
{
    if ((r1_len && !d_r1) || (r2_len && !d_r2)) return fail(c, SSD_EINVAL, "null input pointer");
    if (((uintptr_t)d_r1 & 15u) || ((uintptr_t)d_r2 & 15u)) return fail(c, SSD_EINVAL, "device inputs must be 16-byte aligned");
    CU(c, cudaSetDevice(c->device));
    c->r1 = (const uint8_t*)d_r1;
    c->r2 = (const uint8_t*)d_r2;
    c->r1_len = r1_len;
    c->r2_len = r2_len;
    c->annotated = annotated;
    c->wide = std::max(r1_len, r2_len) >= (1ull << 32) || getenv("SSD_FORCE_WIDE") != nullptr;
    c->phase1_done = c->filter_done = c->unique_done = false;
    c->planned = -1;
    if (std::max(r1_len, r2_len) >= (1ull << 36)) return fail(c, SSD_EINVAL, "a batch is limited to 64 GiB per file");

    const uint32_t nt1 = blocks_for(r1_len, kTile), nt2 = blocks_for(r2_len, kTile);
    uint64_t cap = std::max(r1_len, r2_len) / (uint64_t)c->record_hint + 1024;
    DevCounters h{};
    // the scans park their per-record results until the record numbers arrive and guess the record phase of each tile meanwhile;
    // a wrong guess (lines that look like record starts where none is) repeats the scan with the variant that waits instead
    // (records so short that a tile holds more of them than stage B parks at once count as a wrong guess).  The choice is
    // remembered per file for the next batches of this context.
    if (getenv("SSD_SCAN_NODEFER")) c->defer1 = c->defer2 = false;
    if (const char* sv = getenv("SSD_SCAN_SHORT")) {
        c->short2 = sv[0] == '1';
        c->short_auto = false;
    }
    for (int attempt = 0, overflows = 0; attempt < 5; attempt++) {
        c->rec_cap = cap;
        OKR(ensure(c, c->b_status1, (size_t)(nt1 + 1) * 8));
        OKR(ensure(c, c->b_status2, (size_t)(nt2 + 1) * 8));
        OKR(ensure(c, c->b_r1_start, (cap + 1) * (c->wide ? 8 : 4)));
        OKR(ensure(c, c->b_r1_meta, cap * 16));
        OKR(ensure(c, c->b_r1_base, cap * 4));
        OKR(ensure(c, c->b_r1_tok, cap * 8));
        OKR(ensure(c, c->b_ann, cap * 8));
        OKR(ensure(c, c->b_keys, cap * 8));
        OKR(ensure(c, c->b_irr, cap * sizeof(IrrKey)));
        CU(c, cudaMemsetAsync(c->d_cnt, 0, sizeof(DevCounters), c->stream));
        CU(c, cudaMemsetAsync(c->d_err_off, 0xFF, 2 * sizeof(unsigned long long), c->stream));
        CU(c, cudaMemsetAsync(c->d_hist, 0, c->n_cells_total * sizeof(uint32_t), c->stream));
        CU(c, cudaMemsetAsync(c->b_status1.p, 0, (size_t)(nt1 + 1) * 8, c->stream));
        CU(c, cudaMemsetAsync(c->b_status2.p, 0, (size_t)(nt2 + 1) * 8, c->stream));
        OKR(c->wide ? launch_scans<uint64_t>(c, c->defer1, c->defer2) : launch_scans<uint32_t>(c, c->defer1, c->defer2));
        OKR(fetch_counters(c, &h));
        if (h.err_flags & (kErrRespec | (kErrRespec << 1))) {
            if (h.err_flags & kErrRespec) c->defer1 = false;
            if (h.err_flags & (kErrRespec << 1)) c->defer2 = false;
            c->n_respec++;
            continue;
        }
        if (!(h.err_flags & kErrOverflow)) break;
        if (++overflows == 2) return fail(c, SSD_EINVAL, "record tables overflowed twice (internal)");
        cap = std::max(h.n_records[0], h.n_records[1]) + 1;
    }
    if (h.err_flags & (kErrRespec | (kErrRespec << 1) | kErrOverflow)) return fail(c, SSD_EINVAL, "scan did not settle (internal)");
    if (h.err_flags & (kErrNoAt | kErrKey | kErrDense | kErrIndex | kErrForeign)) {
        unsigned long long off[2];
        CU(c, cudaMemcpy(off, c->d_err_off, sizeof off, cudaMemcpyDeviceToHost));
        if (h.err_flags & kErrDense) return fail(c, SSD_EMALFORMED, "FASTQ records shorter than 32 bytes on average are not supported");
        if (h.err_flags & kErrIndex) return fail(c, SSD_EINDEX, "list index out of range");
        if (h.err_flags & kErrForeign)
            return fail(c, SSD_EINVAL, "record at byte %llu: the cell/UMI fields of its header are not made of this whitelist's barcodes", off[0]);
        if (h.err_flags & kErrNoAt) {
            int f = off[0] != ~0ull ? 0 : 1;
            return fail(c, SSD_EMALFORMED, "read %d: record at byte %llu does not start with '@'", f + 1, off[f]);
        }
        std::string tok = token_at(c, c->r2, c->r2_len, off[1]);
        return fail(c, SSD_EKEY, "%s", tok.c_str());
    }
    // Many read-2 records on the exact path (truncated reads: one thread each, against global memory): the next batches of this
    // context take the scan variant that keeps them on the fast path.  This batch's results are exact either way.
    c->n_exact_last = h.n_exact;
    if (!annotated && c->short_auto && !c->short2 && h.n_exact > 64u && (uint64_t)h.n_exact * 1000u > h.n_records[1]) c->short2 = true;
    c->n_rec_r1 = h.n_records[0];
    c->n_rec_r2 = annotated ? h.n_records[0] : h.n_records[1];
    c->n_rec = std::min(c->n_rec_r1, c->n_rec_r2);
    c->n_keys = h.n_keys;
    c->n_irr = h.n_irr;
    c->phase1_done = true;
    memset(&c->stats, 0, sizeof c->stats);
    c->stats.n_records_r1 = c->n_rec_r1;
    c->stats.n_records_r2 = c->n_rec_r2;
    c->stats.n_regular_keys = c->n_keys;
    c->stats.n_irregular_keys = c->n_irr;
    return SSD_OK;
}

}  // namespace

extern "C" int ssd_demux_device(ssd_ctx* c, const void* d_r1, size_t r1_len, const void* d_r2, size_t r2_len)
{
    if (!c) return SSD_EINVAL;
    return run_phase1(c, d_r1, r1_len, d_r2, r2_len, false);
}

extern "C" int ssd_annotated_device(ssd_ctx* c, const void* d_text, size_t len)
{
    if (!c) return SSD_EINVAL;
    return run_phase1(c, d_text, len, d_text, len, true);
}

// ================================================================================================
// phase 2
// ================================================================================================
extern "C" int ssd_cell_space(const ssd_ctx* c, uint64_t* n)
{
    if (!c || !n) return SSD_EINVAL;
    *n = c->n_cells_total;
    return SSD_OK;
}

extern "C" int ssd_reads_histogram_device(ssd_ctx* c, uint32_t** p)
{
    if (!c || !p) return SSD_EINVAL;
    *p = c->d_hist;
    return SSD_OK;
}

extern "C" int ssd_distinct_histogram_device(ssd_ctx* c, uint32_t** p)
{
    if (!c || !p) return SSD_EINVAL;
    *p = c->d_distinct;
    return SSD_OK;
}

extern "C" int ssd_key_layout(const ssd_ctx* c, int* umi_bits, int* cell_bits)
{
    if (!c) return SSD_EINVAL;
    if (umi_bits) *umi_bits = c->dcfg.umi_bits;
    if (cell_bits) *cell_bits = c->dcfg.cell_bits;
    return SSD_OK;
}

extern "C" int ssd_local_unique_keys(ssd_ctx* c, uint64_t** d_keys, uint64_t* n_keys, void** d_irregular, uint64_t* n_irregular)
{
    if (!c) return SSD_EINVAL;
    if (!c->phase1_done) return fail(c, SSD_ESTATE, "ssd_local_unique_keys before ssd_demux_device");
    CU(c, cudaSetDevice(c->device));
    if (!c->unique_done) {
        StageTimer t(c, 2);
        OKR(sort_unique_both(c, (uint64_t*)c->b_keys.p, c->n_keys, (IrrKey*)c->b_irr.p, c->n_irr, true, &c->uniq_keys, &c->n_uniq, &c->uniq_irr,
                             &c->n_uniq_irr));
        c->unique_done = true;
    }
    if (d_keys) *d_keys = c->uniq_keys;
    if (n_keys) *n_keys = c->n_uniq;
    if (d_irregular) *d_irregular = c->uniq_irr;
    if (n_irregular) *n_irregular = c->n_uniq_irr;
    return SSD_OK;
}

extern "C" int ssd_cell_umi_table(ssd_ctx* c, uint64_t** d_keys, uint32_t** d_reads, uint64_t* n_keys, void** d_irregular, uint32_t** d_irregular_reads,
                                  uint64_t* n_irregular)
{
    if (!c) return SSD_EINVAL;
    if (!c->phase1_done) return fail(c, SSD_ESTATE, "ssd_cell_umi_table before ssd_demux_device");
    CU(c, cudaSetDevice(c->device));
    StageTimer t(c, 2);
    OKR(ensure(c, c->b_keys, std::max<uint64_t>(c->n_keys, 1) * sizeof(uint64_t)));
    OKR(ensure(c, c->b_irr, std::max<uint64_t>(c->n_irr, 1) * sizeof(IrrKey)));
    OKR(ensure(c, c->b_first, (c->n_keys + c->n_irr + 2) * sizeof(uint32_t)));
    uint32_t *first_k = (uint32_t*)c->b_first.p, *first_i = first_k + c->n_keys + 1;
    OKR(sort_unique_both(c, (uint64_t*)c->b_keys.p, c->n_keys, (IrrKey*)c->b_irr.p, c->n_irr, true, &c->uniq_keys, &c->n_uniq, &c->uniq_irr,
                         &c->n_uniq_irr, first_k, first_i));
    c->unique_done = true;
    OKR(ensure(c, c->b_runs, (c->n_uniq + c->n_uniq_irr + 2) * sizeof(uint32_t)));
    uint32_t *runs_k = (uint32_t*)c->b_runs.p, *runs_i = runs_k + c->n_uniq + 1;
    if (c->n_uniq) k_run_lengths<<<blocks_for(c->n_uniq, kThreads), kThreads, 0, c->stream>>>(first_k, c->n_uniq, c->n_keys, runs_k);
    if (c->n_uniq_irr) k_run_lengths<<<blocks_for(c->n_uniq_irr, kThreads), kThreads, 0, c->stream>>>(first_i, c->n_uniq_irr, c->n_irr, runs_i);
    c->n_launches += (c->n_uniq ? 1 : 0) + (c->n_uniq_irr ? 1 : 0);
    CU(c, cudaGetLastError());
    if (d_keys) *d_keys = c->uniq_keys;
    if (d_reads) *d_reads = runs_k;
    if (n_keys) *n_keys = c->n_uniq;
    if (d_irregular) *d_irregular = c->uniq_irr;
    if (d_irregular_reads) *d_irregular_reads = runs_i;
    if (n_irregular) *n_irregular = c->n_uniq_irr;
    return SSD_OK;
}

namespace {
// buckets of the owner partition: cell >> shift, at most 256 of them
void owner_layout(const ssd_ctx* c, int* shift, int* n_buckets)
{
    int s = 0;
    while (((c->n_cells_total - 1) >> s) >= (uint64_t)kOsRadix) s++;
    *shift = s;
    *n_buckets = (int)((c->n_cells_total - 1) >> s) + 1;
}
}  // namespace

extern "C" int ssd_owner_layout(const ssd_ctx* c, int* shift, int* n_buckets)
{
    if (!c) return SSD_EINVAL;
    int s, nb;
    owner_layout(c, &s, &nb);
    if (shift) *shift = s;
    if (n_buckets) *n_buckets = nb;
    return SSD_OK;
}

extern "C" int ssd_partition_keys(ssd_ctx* c, int world, uint64_t** d_keys, uint64_t* offsets, void** d_irregular, uint64_t* n_irregular)
{
    if (!c || world < 1 || !offsets) return SSD_EINVAL;
    if (!c->phase1_done) return fail(c, SSD_ESTATE, "ssd_partition_keys before ssd_demux_device");
    CU(c, cudaSetDevice(c->device));
    StageTimer t(c, 2);
    c->unique_done = false;
    int shift, nb;
    owner_layout(c, &shift, &nb);
    const uint64_t n_keys = c->n_keys, n_rec = c->n_rec_r2;
    if (c->n_irr) {
        // next to the partition pass, on the side stream
        unsigned long long* cursor = (unsigned long long*)c->d_err_off;  // free after phase 1 (errors were read)
        OKR(ensure(c, c->b_irr, c->n_irr * sizeof(IrrKey)));
        cudaStream_t st = c->side_stream ? c->side_stream : c->stream;
        if (c->side_stream) {
            CU(c, cudaEventRecord(c->ev_fork, c->stream));
            CU(c, cudaStreamWaitEvent(st, c->ev_fork, 0));
        }
        CU(c, cudaMemsetAsync(cursor, 0, sizeof(unsigned long long), st));
        k_collect_irr<<<blocks_for(n_rec, kThreads), kThreads, 0, st>>>((const uint64_t*)c->b_ann.p, n_rec, c->r2, (IrrKey*)c->b_irr.p, cursor);
        c->n_launches++;
    }
    uint32_t hist[kOsRadix] = {0};
    if (n_keys) {
        // one scatter pass on the bucket digit: keys come out grouped by bucket, hence by owner
        const uint64_t tiles = (n_rec + kOsTile - 1) / kOsTile;
        const size_t head_words = (size_t)kOsMaxPasses * kOsRadix + 64, status_words = (size_t)tiles * kOsRadix;
        OKR(ensure(c, c->b_os, (head_words + status_words) * sizeof(uint32_t)));
        OKR(ensure(c, c->b_keys, n_keys * sizeof(uint64_t)));
        uint32_t* ghist = (uint32_t*)c->b_os.p;
        CU(c, cudaMemsetAsync(c->b_os.p, 0, (head_words + status_words) * sizeof(uint32_t), c->stream));
        const KeysFromAnn ka{(const uint64_t*)c->b_ann.p, c->dcfg.umi_bits};
        const int sh = c->dcfg.umi_bits + shift;
        k_os_hist<KeysFromAnn><<<(uint32_t)std::min<uint64_t>(blocks_for(n_rec, kThreads), 148 * 8), kThreads, 0, c->stream>>>(ka, n_rec, 1, ghist, sh);
        k_os_pass<KeysFromAnn><<<(uint32_t)tiles, kThreads, 0, c->stream>>>(ka, n_rec, sh, (uint64_t*)c->b_keys.p, ghist, ghist + head_words,
                                                                              ghist + (size_t)kOsMaxPasses * kOsRadix);
        c->n_launches += 2;
        CU(c, cudaMemcpyAsync(hist, ghist, sizeof hist, cudaMemcpyDeviceToHost, c->stream));
    }
    if (c->n_irr && c->side_stream) {
        CU(c, cudaEventRecord(c->ev_join, c->side_stream));
        CU(c, cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    }
    CU(c, cudaGetLastError());
    CU(c, cudaStreamSynchronize(c->stream));
    uint64_t run = 0;
    int b = 0;
    for (int k = 0; k < world; k++) {
        offsets[k] = run;
        for (; b < nb && (uint64_t)b * (uint64_t)world / (uint64_t)nb == (uint64_t)k; b++) run += hist[b];
    }
    offsets[world] = run;
    if (run != n_keys) return fail(c, SSD_EINVAL, "key partition lost keys (internal)");
    if (d_keys) *d_keys = (uint64_t*)c->b_keys.p;
    if (d_irregular) *d_irregular = c->b_irr.p;
    if (n_irregular) *n_irregular = c->n_irr;
    return SSD_OK;
}

extern "C" int ssd_count_distinct(ssd_ctx* c, uint64_t* d_keys, uint64_t n_keys, void* d_irregular, uint64_t n_irregular)
{
    if (!c) return SSD_EINVAL;
    CU(c, cudaSetDevice(c->device));
    uint64_t *uk = nullptr, nuk = 0, nui = 0;
    IrrKey* ui = nullptr;
    StageTimer t(c, 2);
    if (c->unique_done && d_keys == c->uniq_keys && n_keys == c->n_uniq && d_irregular == (void*)c->uniq_irr && n_irregular == c->n_uniq_irr) {
        uk = c->uniq_keys;  // already sorted and unique: only count
        nuk = c->n_uniq;
        ui = c->uniq_irr;
        nui = c->n_uniq_irr;
    } else {
        c->unique_done = false;  // the context's buffers are reused as scratch below
        return distinct_counts_arrays(c, d_keys, n_keys, (const IrrKey*)d_irregular, n_irregular);
    }
    return count_cells(c, uk, nuk, ui, nui);
}

extern "C" int ssd_count_distinct_owned(ssd_ctx* c, uint64_t* d_keys, uint64_t n_keys, void* d_irregular, uint64_t n_irregular, int rank, int world)
{
    if (!c || world < 1 || rank < 0 || rank >= world) return SSD_EINVAL;
    CU(c, cudaSetDevice(c->device));
    StageTimer t(c, 2);
    c->unique_done = false;  // the context's buffers are reused as scratch
    return distinct_counts_arrays(c, d_keys, n_keys, (const IrrKey*)d_irregular, n_irregular, rank, world);
}

extern "C" int ssd_build_filter(ssd_ctx* c, ssd_stats* stats)
{
    if (!c) return SSD_EINVAL;
    if (!c->phase1_done) return fail(c, SSD_ESTATE, "ssd_build_filter before ssd_demux_device");
    CU(c, cudaSetDevice(c->device));
    DevCounters h{};
    StageTimer t(c, 3);
    c->n_launches += 2 + (c->cfg.collapse_pairs > 0 ? 1 : 0);
    CU(c, cudaMemsetAsync((char*)c->d_cnt + offsetof(DevCounters, n_matched), 0, 6 * sizeof(unsigned long long), c->stream));
    const uint32_t nbc = blocks_for(c->n_cells_total, kThreads);
    k_filter<<<nbc, kThreads, 0, c->stream>>>(c->d_hist, c->d_distinct, c->n_cells_total, (long long)c->cfg.min_count, c->cfg.filter_mode, c->d_pass,
                                              c->d_cnt);
    if (c->cfg.collapse_pairs > 0) k_collapse_map<<<nbc, kThreads, 0, c->stream>>>(c->d_pass, c->dcfg.n_wl[0], c->dcfg.n_wl[1] * c->dcfg.n_wl[2], c->cfg.collapse_pairs, c->d_remap);
    {
        // exclusive scan of the post-filter output lengths; its first kernel also counts the retained records
        // and sizes the pre-filter output
        const uint32_t nb = (uint32_t)(c->n_rec / kScanBlock + 1);
        OKR(ensure(c, c->b_bsums, ((size_t)nb + 2) * sizeof(unsigned long long)));
        OKR(ensure(c, c->b_out_off, (c->n_rec + 1) * sizeof(uint64_t)));
        unsigned long long* bs = (unsigned long long*)c->b_bsums.p;  // [0]: ticket, [1..]: tile status
        OutLen f = make_outlen(c, SSD_EMIT_PASSING);
        CU(c, cudaMemsetAsync(bs, 0, ((size_t)nb + 2) * sizeof(unsigned long long), c->stream));
        k_plan_onepass<<<nb, kThreads, 0, c->stream>>>(f, c->n_rec, (unsigned long long*)c->b_out_off.p, bs + 1, (uint32_t*)bs, c->d_cnt);
        c->planned = SSD_EMIT_PASSING;
    }
    CU(c, cudaGetLastError());
    OKR(fetch_counters(c, &h));
    c->stats.n_matched = c->n_keys + c->n_irr;
    c->stats.n_cells = h.n_cells;
    c->stats.n_passing_cells = h.n_passing_cells;
    c->stats.n_retained = h.n_retained;
    c->stats.bytes_matched = h.bytes_matched;
    c->stats.bytes_passing = h.bytes_passing;
    c->filter_done = true;
    if (stats) *stats = c->stats;
    return SSD_OK;
}

extern "C" int ssd_distinct_local(ssd_ctx* c)
{
    if (!c) return SSD_EINVAL;
    if (!c->phase1_done) return fail(c, SSD_ESTATE, "ssd_distinct_local before ssd_demux_device");
    CU(c, cudaSetDevice(c->device));
    c->unique_done = false;
    StageTimer t(c, 2);
    return distinct_counts_local(c);
}

namespace {
// -m as the kernels of the exchange compare it (distinct counts are 32-bit): never = no count can reach it
inline uint32_t min_count_u32(const ssd_ctx* c, bool* never)
{
    *never = c->cfg.min_count > 0xFFFFFFFFll;
    return (uint32_t)std::min<long long>(std::max<long long>(c->cfg.min_count, 0), 0xFFFFFFFFll);
}
}  // namespace

extern "C" int ssd_pack_decision(ssd_ctx* c, uint64_t** d_packed)
{
    if (!c || !d_packed) return SSD_EINVAL;
    if (!c->phase1_done) return fail(c, SSD_ESTATE, "ssd_pack_decision before ssd_demux_device");
    CU(c, cudaSetDevice(c->device));
    StageTimer t(c, 2);
    const uint64_t nc = c->n_cells_total;
    OKR(ensure(c, c->b_pack, nc * sizeof(uint64_t)));
    bool never;
    const uint32_t m = min_count_u32(c, &never);
    k_pack_decision<<<blocks_for(nc, kThreads), kThreads, 0, c->stream>>>(c->d_hist, c->d_distinct, nc, m, never, (unsigned long long*)c->b_pack.p);
    c->n_launches++;
    CU(c, cudaGetLastError());
    *d_packed = (uint64_t*)c->b_pack.p;
    return SSD_OK;
}

extern "C" int ssd_undecided_keys(ssd_ctx* c, const uint64_t* d_packed_global, uint64_t capacity, void** d_records)
{
    if (!c || !d_packed_global || !d_records || capacity == 0) return SSD_EINVAL;
    if (!c->phase1_done) return fail(c, SSD_ESTATE, "ssd_undecided_keys before ssd_demux_device");
    CU(c, cudaSetDevice(c->device));
    StageTimer t(c, 2);
    const uint64_t n = c->n_rec_r2, nc = c->n_cells_total, words = (nc + 31) / 32;
    OKR(ensure(c, c->b_und, (capacity + 1) * sizeof(IrrKey)));
    OKR(ensure(c, c->b_und_bits, words * sizeof(uint32_t)));
    CU(c, cudaMemsetAsync(c->b_und.p, 0xFF, (capacity + 1) * sizeof(IrrKey), c->stream));  // unused records are padding
    CU(c, cudaMemsetAsync(c->d_ucnt, 0, sizeof(unsigned long long), c->stream));
    bool never;
    const uint32_t m = min_count_u32(c, &never);
    k_mark_undecided<<<blocks_for(nc, kThreads), kThreads, 0, c->stream>>>((const unsigned long long*)d_packed_global, nc, m, (uint32_t*)c->b_und_bits.p);
    c->n_launches++;
    if (n) {
        k_undecided_packed<<<(uint32_t)std::min<uint64_t>(blocks_for(n, 1024), 148 * 8), 256, 0, c->stream>>>(
            (const uint64_t*)c->b_ann.p, n, c->r2, (const uint32_t*)c->b_und_bits.p, c->dcfg.umi_len, (IrrKey*)c->b_und.p, capacity, c->d_ucnt);
        c->n_launches++;
    }
    k_undecided_header<<<1, 1, 0, c->stream>>>((IrrKey*)c->b_und.p, c->d_ucnt);
    c->n_launches++;
    CU(c, cudaGetLastError());
    *d_records = c->b_und.p;
    return SSD_OK;
}

extern "C" int ssd_adopt_decision(ssd_ctx* c, const uint64_t* d_packed_global)
{
    if (!c || !d_packed_global) return SSD_EINVAL;
    if (!c->phase1_done) return fail(c, SSD_ESTATE, "ssd_adopt_decision before ssd_demux_device");
    CU(c, cudaSetDevice(c->device));
    StageTimer t(c, 2);
    bool never;
    const uint32_t m = min_count_u32(c, &never);
    k_adopt_decision<<<blocks_for(c->n_cells_total, kThreads), kThreads, 0, c->stream>>>((const unsigned long long*)d_packed_global, c->n_cells_total, m, c->d_hist,
                                                                                        c->d_distinct);
    c->n_launches++;
    CU(c, cudaGetLastError());
    return SSD_OK;
}

extern "C" int ssd_get_stats(const ssd_ctx* c, ssd_stats* stats)
{
    if (!c || !stats) return SSD_EINVAL;
    *stats = c->stats;
    return SSD_OK;
}

extern "C" int ssd_filter_single(ssd_ctx* c, ssd_stats* stats)
{
    if (!c) return SSD_EINVAL;
    if (c->cfg.filter_mode == SSD_FILTER_DISTINCT_UMI) {
        if (!c->phase1_done) return fail(c, SSD_ESTATE, "ssd_filter_single before ssd_demux_device");
        CU(c, cudaSetDevice(c->device));
        if (c->unique_done || getenv("SSD_DISTINCT_LISTS")) {  // the unique-key lists exist (or are asked for): count those
            uint64_t *k = nullptr, nk = 0, ni = 0;
            void* ik = nullptr;
            OKR(ssd_local_unique_keys(c, &k, &nk, &ik, &ni));
            OKR(ssd_count_distinct(c, k, nk, ik, ni));
        } else {
            StageTimer t(c, 2);
            OKR(distinct_counts_local(c));
        }
    }
    return ssd_build_filter(c, stats);
}

// ================================================================================================
// phase 3
// ================================================================================================
extern "C" int ssd_emit_device(ssd_ctx* c, int which, void* d_out, size_t cap, size_t* written)
{
    if (!c) return SSD_EINVAL;
    if (!c->filter_done) return fail(c, SSD_ESTATE, "ssd_emit_device before ssd_build_filter");
    if (which != SSD_EMIT_MATCHED && which != SSD_EMIT_PASSING) return fail(c, SSD_EINVAL, "unknown output selector");
    CU(c, cudaSetDevice(c->device));
    const uint64_t need = which == SSD_EMIT_MATCHED ? c->stats.bytes_matched : c->stats.bytes_passing;
    if (written) *written = (size_t)need;
    if (need > cap) return fail(c, SSD_ERANGE, "output needs %llu bytes, buffer has %zu", (unsigned long long)need, cap);
    if (need == 0) return SSD_OK;
    if (!d_out) return fail(c, SSD_EINVAL, "null output pointer");
    if (c->planned != which) {
        StageTimer t(c, 4);
        OKR(ensure(c, c->b_out_off, (c->n_rec + 1) * sizeof(uint64_t)));
        OKR(exclusive_scan<unsigned long long>(c, make_outlen(c, which), c->n_rec, (unsigned long long*)c->b_out_off.p));
        c->planned = which;
    }
    OKR(c->wide ? launch_emit<uint64_t>(c, which, (uint8_t*)d_out) : launch_emit<uint32_t>(c, which, (uint8_t*)d_out));
    DevCounters h{};
    OKR(fetch_counters(c, &h));
    if (h.err_flags & kErrEmitLen) return fail(c, SSD_ECUDA, "internal: emitted record length differs from the planned length");
    return SSD_OK;
}

// Split mode: the post-filter records grouped into per-cell buckets (cells ascending, input order inside a cell).
extern "C" int ssd_emit_split_device(ssd_ctx* c, void* d_out, size_t cap, size_t* written, uint64_t* n_buckets)
{
    if (!c) return SSD_EINVAL;
    if (!c->filter_done) return fail(c, SSD_ESTATE, "ssd_emit_split_device before ssd_build_filter");
    CU(c, cudaSetDevice(c->device));
    const uint64_t need = c->stats.bytes_passing, n = c->stats.n_retained;
    if (written) *written = (size_t)need;
    if (n_buckets) *n_buckets = 0;
    c->n_buckets = 0;
    c->split_bytes = need;
    if (need > cap) return fail(c, SSD_ERANGE, "output needs %llu bytes, buffer has %zu", (unsigned long long)need, cap);
    if (n == 0) return SSD_OK;
    if (!d_out) return fail(c, SSD_EINVAL, "null output pointer");
    if (n >= (1ull << 32)) return fail(c, SSD_EINVAL, "split mode is limited to 2^32 records per batch");
    OutLen f = make_outlen(c, SSD_EMIT_PASSING);
    // 1. (cell, record) pairs of the written records, in input order
    OKR(ensure(c, c->b_pos, (c->n_rec + 1) * sizeof(uint32_t)));
    OKR(ensure(c, c->b_sp_cells, n * 8));
    OKR(ensure(c, c->b_sp_cells_alt, n * 8));
    OKR(ensure(c, c->b_sp_recs, n * 4));
    OKR(ensure(c, c->b_sp_recs_alt, n * 4));
    OKR(ensure(c, c->b_sp_off, (n + 1) * 8));
    uint32_t* pos = (uint32_t*)c->b_pos.p;
    OKR(exclusive_scan<uint32_t>(c, PassFlag{f}, c->n_rec, pos));
    k_split_pairs<<<blocks_for(c->n_rec, kThreads), kThreads, 0, c->stream>>>(f, pos, c->n_rec, (uint64_t*)c->b_sp_cells.p, (uint32_t*)c->b_sp_recs.p);
    // 2. stable sort by cell
    uint64_t* cells = nullptr;
    uint32_t* order = nullptr;
    OKR(radix_sort64<true>(c, (uint64_t*)c->b_sp_cells.p, (uint64_t*)c->b_sp_cells_alt.p, (uint32_t*)c->b_sp_recs.p, (uint32_t*)c->b_sp_recs_alt.p, n,
                           c->dcfg.cell_bits, &cells, &order));
    // 3. byte offsets in bucket order, 4. the writer through the permutation
    unsigned long long* off = (unsigned long long*)c->b_sp_off.p;
    OKR(exclusive_scan<unsigned long long>(c, SplitLen{f, order}, n, off));
    OKR(c->wide ? launch_emit<uint64_t>(c, SSD_EMIT_PASSING, (uint8_t*)d_out, order, off, n)
                : launch_emit<uint32_t>(c, SSD_EMIT_PASSING, (uint8_t*)d_out, order, off, n));
    // 5. bucket table
    OKR(exclusive_scan<uint32_t>(c, BucketHead{cells}, n, pos));
    uint32_t nb = 0;
    CU(c, cudaMemcpyAsync(&nb, pos + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    OKR(ensure(c, c->b_bk_cell, (size_t)nb * 4));
    OKR(ensure(c, c->b_bk_off, (size_t)nb * 8));
    OKR(ensure(c, c->b_bk_first, (size_t)nb * 8));
    k_bucket_table<<<blocks_for(n, kThreads), kThreads, 0, c->stream>>>(cells, pos, off, n, (uint32_t*)c->b_bk_cell.p, (unsigned long long*)c->b_bk_off.p,
                                                                       (unsigned long long*)c->b_bk_first.p);
    c->n_launches += 3;
    DevCounters h{};
    OKR(fetch_counters(c, &h));
    if (h.err_flags & kErrEmitLen) return fail(c, SSD_ECUDA, "internal: emitted record length differs from the planned length");
    c->n_buckets = nb;
    if (n_buckets) *n_buckets = nb;
    return SSD_OK;
}

// Bucket table of the last ssd_emit_split_device: cell id, first byte and first record position of every bucket
// (bucket b spans bytes [offsets[b], offsets[b+1]) and positions [first[b], first[b+1]); the last one ends at the totals).
extern "C" int ssd_split_index(ssd_ctx* c, uint32_t* cells_host, uint64_t* offsets_host, uint64_t* first_host, uint64_t capacity)
{
    if (!c) return SSD_EINVAL;
    if (capacity < c->n_buckets) return fail(c, SSD_ERANGE, "need room for %llu buckets", (unsigned long long)c->n_buckets);
    if (!c->n_buckets) return SSD_OK;
    CU(c, cudaSetDevice(c->device));
    if (cells_host) CU(c, cudaMemcpyAsync(cells_host, c->b_bk_cell.p, c->n_buckets * 4, cudaMemcpyDeviceToHost, c->stream));
    if (offsets_host) CU(c, cudaMemcpyAsync(offsets_host, c->b_bk_off.p, c->n_buckets * 8, cudaMemcpyDeviceToHost, c->stream));
    if (first_host) CU(c, cudaMemcpyAsync(first_host, c->b_bk_first.p, c->n_buckets * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return SSD_OK;
}

extern "C" size_t ssd_output_bound(const ssd_ctx* c, size_t r1_len, size_t r2_len)
{
    (void)c;
    (void)r2_len;
    return r1_len + 48 * (r1_len / 16 + 1);
}

extern "C" int ssd_host_alloc(void** p, size_t bytes)
{
    if (!p) return SSD_EINVAL;
    return cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocDefault) == cudaSuccess ? SSD_OK : SSD_ENOMEM;
}

extern "C" int ssd_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? SSD_OK : SSD_ECUDA; }

namespace {
// Contexts used from different host threads pipeline over PCIe: one batch uploads while another computes and a third
// downloads.  Copies in the same direction share one DMA engine; letting two uploads interleave only delays both, so
// each direction is handed to one context at a time.
std::mutex g_h2d_turn, g_d2h_turn;
}  // namespace

extern "C" int ssd_emit_host(ssd_ctx* c, int which, void* out, size_t out_cap, size_t* written)
{
    if (!c) return SSD_EINVAL;
    if (!c->filter_done) return fail(c, SSD_ESTATE, "ssd_emit_host before the filter was built");
    const uint64_t need = which == SSD_EMIT_MATCHED ? c->stats.bytes_matched : c->stats.bytes_passing;
    if (written) *written = (size_t)need;
    if (need > out_cap) return fail(c, SSD_ERANGE, "output needs %llu bytes, buffer has %zu", (unsigned long long)need, out_cap);
    if (!need) return SSD_OK;
    OKR(ensure(c, c->b_out, need + 32));
    OKR(ssd_emit_device(c, which, c->b_out.p, c->b_out.cap, nullptr));
    std::lock_guard<std::mutex> turn(g_d2h_turn);
    CU(c, cudaMemcpyAsync(out, c->b_out.p, need, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return SSD_OK;
}

extern "C" int ssd_run_host(ssd_ctx* c, const void* r1, size_t r1_len, const void* r2, size_t r2_len, int which, void* out, size_t out_cap,
                            size_t* written, ssd_stats* stats)
{
    if (!c) return SSD_EINVAL;
    CU(c, cudaSetDevice(c->device));
    OKR(ensure(c, c->b_in1, r1_len + 32));
    OKR(ensure(c, c->b_in2, r2_len + 32));
    // chunked so that a pageable source does not stall the whole transfer behind one staging copy
    const size_t chunk = (size_t)64 << 20;
    {
        std::lock_guard<std::mutex> turn(g_h2d_turn);
        for (size_t o = 0; o < r1_len; o += chunk)
            CU(c, cudaMemcpyAsync((uint8_t*)c->b_in1.p + o, (const uint8_t*)r1 + o, std::min(chunk, r1_len - o), cudaMemcpyHostToDevice, c->stream));
        for (size_t o = 0; o < r2_len; o += chunk)
            CU(c, cudaMemcpyAsync((uint8_t*)c->b_in2.p + o, (const uint8_t*)r2 + o, std::min(chunk, r2_len - o), cudaMemcpyHostToDevice, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
    }
    OKR(ssd_demux_device(c, c->b_in1.p, r1_len, c->b_in2.p, r2_len));
    OKR(ssd_filter_single(c, stats));
    if (!out && !out_cap) {
        if (written) *written = (size_t)(which == SSD_EMIT_MATCHED ? c->stats.bytes_matched : c->stats.bytes_passing);
        return SSD_OK;
    }
    return ssd_emit_host(c, which, out, out_cap, written);
}

// ================================================================================================
// file to file
// ================================================================================================
namespace {
// A byte range of a file mapped into memory.  Measured on the GPU box (profiles/r02_host_io.md, 16 host threads, RAM-backed
// files): READING is fastest with plain pread into the pinned slots (42 GB/s; through a mapping 34 GB/s, and tearing down the
// page tables of the mapping afterwards costs another 50 ms per 6 GB), WRITING is fastest through a shared mapping (memcpy
// from the pinned slots: 15-19 GB/s over pages the file already has, 6.6-8 GB/s when every page is new, against 5.7 / 3.5
// GB/s with pwrite).  So inputs are read with pread (SSD_IO_MMAP_INPUT=1 maps them instead) and the output range is mapped
// when the file system is RAM-backed (on a disk the dirty pages of a mapping are written back at the kernel's leisure,
// pwrite is the better citizen there).  SSD_IO_NOMMAP=1 forces pread / pwrite everywhere.
struct RangeMap {
    void* p = MAP_FAILED;
    size_t len = 0;
    uint8_t* data = nullptr;  // first byte of the range
    bool open_range(int fd, size_t base, size_t size, bool writable)
    {
        static const bool off = getenv("SSD_IO_NOMMAP") != nullptr, map_input = getenv("SSD_IO_MMAP_INPUT") != nullptr;
        if (off || size == 0 || (!writable && !map_input)) return false;
        if (writable) {
            struct statfs fs;
            if (fstatfs(fd, &fs) != 0 || (unsigned long)fs.f_type != 0x01021994ul) return false;  // TMPFS_MAGIC
            if (ftruncate(fd, (off_t)(base + size)) != 0) return false;
        }
        const size_t page = (size_t)sysconf(_SC_PAGESIZE), lo = base & ~(page - 1);
        len = size + (base - lo);
        p = mmap(nullptr, len, writable ? PROT_READ | PROT_WRITE : PROT_READ, MAP_SHARED, fd, (off_t)lo);
        if (p == MAP_FAILED) return false;
        madvise(p, len, MADV_SEQUENTIAL);
        data = (uint8_t*)p + (base - lo);
        return true;
    }
    void close_range()
    {
        if (p != MAP_FAILED) munmap(p, len);
        p = MAP_FAILED;
        data = nullptr;
    }
    // tearing down the page tables of a multi-gigabyte mapping takes tens of milliseconds: off the caller's critical path
    void close_range_later()
    {
        if (p != MAP_FAILED) {
            void* q = p;
            const size_t n = len;
            std::thread([q, n] { munmap(q, n); }).detach();
        }
        p = MAP_FAILED;
        data = nullptr;
    }
};

// one host thread of `stride`: chunks first, first + stride, ... of the range into its pinned slots (memcpy from the mapping, or
// pread), each uploaded as soon as it is there
void upload_file_worker(int device, int fd, const uint8_t* map, size_t base, size_t size, uint8_t* d_dst, IoLane* lane, int first, int stride, int* rc,
                        int* err_no)
{
    *rc = SSD_OK;
    if (cudaSetDevice(device) != cudaSuccess) {
        *rc = SSD_ECUDA;
        return;
    }
    int n = 0;
    for (size_t k = (size_t)first; k * kIoChunk < size; k += (size_t)stride, n++) {
        const int s = n % kIoSlots;
        if (n >= kIoSlots && cudaEventSynchronize(lane->done[s]) != cudaSuccess) {
            *rc = SSD_ECUDA;
            return;
        }
        const size_t off = k * kIoChunk, want = std::min(kIoChunk, size - off);
        size_t got = 0;
        if (map) {
            memcpy(lane->slot[s], map + off, want);
            got = want;
        }
        while (got < want) {
            const ssize_t r = pread(fd, lane->slot[s] + got, want - got, (off_t)(base + off + got));
            if (r < 0 && errno == EINTR) continue;
            if (r <= 0) {
                *rc = SSD_EIO;
                *err_no = r < 0 ? errno : EIO;
                return;
            }
            got += (size_t)r;
        }
        if (cudaMemcpyAsync(d_dst + off, lane->slot[s], want, cudaMemcpyHostToDevice, lane->stream) != cudaSuccess ||
            cudaEventRecord(lane->done[s], lane->stream) != cudaSuccess) {
            *rc = SSD_ECUDA;
            return;
        }
    }
    if (cudaStreamSynchronize(lane->stream) != cudaSuccess) *rc = SSD_ECUDA;
}

// one host thread of `stride`: download chunks first, first + stride, ... and put them at base + offset (memcpy into the mapping, or pwrite)
void download_file_worker(int device, int fd, uint8_t* map, size_t base, size_t size, const uint8_t* d_src, IoLane* lane, int first, int stride, int* rc,
                          int* err_no)
{
    *rc = SSD_OK;
    if (cudaSetDevice(device) != cudaSuccess) {
        *rc = SSD_ECUDA;
        return;
    }
    auto issue = [&](size_t k, int s) {
        const size_t off = k * kIoChunk;
        return cudaMemcpyAsync(lane->slot[s], d_src + off, std::min(kIoChunk, size - off), cudaMemcpyDeviceToHost, lane->stream) == cudaSuccess &&
               cudaEventRecord(lane->done[s], lane->stream) == cudaSuccess;
    };
    size_t k = (size_t)first;
    if (k * kIoChunk < size && !issue(k, 0)) {
        *rc = SSD_ECUDA;
        return;
    }
    for (int n = 0; k * kIoChunk < size; k += (size_t)stride, n++) {
        const int s = n % kIoSlots;
        const size_t next = k + (size_t)stride;
        if (next * kIoChunk < size && !issue(next, (n + 1) % kIoSlots)) {  // the other slot was written out in the previous round
            *rc = SSD_ECUDA;
            return;
        }
        if (cudaEventSynchronize(lane->done[s]) != cudaSuccess) {
            *rc = SSD_ECUDA;
            return;
        }
        const size_t off = k * kIoChunk, len = std::min(kIoChunk, size - off);
        size_t put = 0;
        if (map) {
            memcpy(map + off, lane->slot[s], len);
            put = len;
        }
        while (put < len) {
            const ssize_t w = pwrite(fd, lane->slot[s] + put, len - put, (off_t)(base + off + put));
            if (w < 0 && errno == EINTR) continue;
            if (w <= 0) {
                *rc = SSD_EIO;
                *err_no = w < 0 ? errno : EIO;
                return;
            }
            put += (size_t)w;
        }
    }
    if (cudaStreamSynchronize(lane->stream) != cudaSuccess) *rc = SSD_ECUDA;
}

// SSD_IO_TRACE=1: wall-clock milliseconds of the phases of the file-to-file calls on stderr (profiles/files_probe.py reads them)
struct IoTrace {
    bool on;
    std::chrono::steady_clock::time_point t;
    IoTrace() : t(std::chrono::steady_clock::now())
    {
        static const bool v = getenv("SSD_IO_TRACE") != nullptr;
        on = v;
    }
    void lap(const char* what, size_t bytes = 0)
    {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        const double ms = std::chrono::duration<double, std::milli>(now - t).count();
        if (bytes) fprintf(stderr, "[ssd io] %-18s %9.2f ms  %7.2f GB/s\n", what, ms, (double)bytes / ms / 1e6);
        else fprintf(stderr, "[ssd io] %-18s %9.2f ms\n", what, ms);
        t = now;
    }
};

// all lanes in [lane0, lane0 + n) work on one file; returns the first failure
template <class Worker, class... A>
int run_lanes(ssd_ctx* c, const char* path, int lane0, int n, Worker worker, A... a)
{
    int rc[kIoLanes], err[kIoLanes];
    std::thread th[kIoLanes];
    for (int t = 0; t < n; t++) {
        rc[t] = SSD_OK;
        err[t] = 0;
        th[t] = std::thread(worker, a..., &c->io[lane0 + t], t, n, &rc[t], &err[t]);
    }
    for (int t = 0; t < n; t++) th[t].join();
    for (int t = 0; t < n; t++) {
        if (rc[t] == SSD_EIO) return fail(c, SSD_EIO, "%s: %s", path, strerror(err[t]));
        if (rc[t] != SSD_OK) return fail(c, rc[t], "moving %s failed", path);
    }
    return SSD_OK;
}

int open_input(ssd_ctx* c, const char* path, int* fd, size_t* size)
{
    *fd = open(path, O_RDONLY);
    struct stat st;
    if (*fd < 0 || fstat(*fd, &st) != 0) {
        const int e = errno;
        if (*fd >= 0) close(*fd);
        *fd = -1;
        return fail(c, SSD_EIO, "%s: %s", path, strerror(e));
    }
    *size = (size_t)st.st_size;
    return SSD_OK;
}

// emit `which` on the device, then download it with all I/O lanes, each overlapping its copies with its writes into the file
int download_to_file(ssd_ctx* c, int which, const char* out_path, int append, size_t* written)
{
    const uint64_t need = which == SSD_EMIT_MATCHED ? c->stats.bytes_matched : c->stats.bytes_passing;
    if (written) *written = (size_t)need;
    IoTrace tr;
    // O_RDWR: a shared writable mapping needs it.  Not O_TRUNC: a file that is being replaced keeps the pages it has (cutting
    // it to the new length below frees only the excess), so that writing over it costs a copy, not a page allocation per 4 KiB
    const int fd = open(out_path, O_RDWR | O_CREAT, 0644);
    if (fd < 0) return fail(c, SSD_EIO, "%s: %s", out_path, strerror(errno));
    const off_t base = append ? lseek(fd, 0, SEEK_END) : 0;  // V2:367, 415 open in append mode
    int rc = base < 0 ? fail(c, SSD_EIO, "%s: %s", out_path, strerror(errno)) : SSD_OK;
    if (rc == SSD_OK && !append && ftruncate(fd, (off_t)need) != 0) rc = fail(c, SSD_EIO, "%s: %s", out_path, strerror(errno));
    tr.lap("open output");
    if (need && rc == SSD_OK) {
        rc = ensure(c, c->b_out, need + 32);
        if (rc == SSD_OK) rc = ssd_emit_device(c, which, c->b_out.p, c->b_out.cap, nullptr);
        for (int k = 0; k < kIoLanes && rc == SSD_OK; k++) rc = c->io[k].init();
        if (rc == SSD_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = fail(c, SSD_ECUDA, "emit failed");  // the emit kernel ran on the context's stream
        tr.lap("emit kernel");
        if (rc == SSD_OK) {
            RangeMap m;
            m.open_range(fd, (size_t)base, (size_t)need, true);
            tr.lap("map output");
            {
                std::lock_guard<std::mutex> turn(g_d2h_turn);
                rc = run_lanes(c, out_path, 0, kIoLanes, download_file_worker, c->device, fd, m.data, (size_t)base, (size_t)need, (const uint8_t*)c->b_out.p);
            }
            tr.lap("download + write", (size_t)need);
            m.close_range_later();
            tr.lap("unmap output");
        }
    }
    if (close(fd) != 0 && rc == SSD_OK) rc = fail(c, SSD_EIO, "%s: %s", out_path, strerror(errno));
    tr.lap("close output");
    return rc;
}
}  // namespace

extern "C" int ssd_run_files(ssd_ctx* c, const char* path_r1, const char* path_r2, int which, const char* out_path, int append, size_t* written,
                             ssd_stats* stats)
{
    if (!c || !path_r1 || !path_r2 || !out_path) return SSD_EINVAL;
    if (which != SSD_EMIT_MATCHED && which != SSD_EMIT_PASSING) return fail(c, SSD_EINVAL, "unknown output selector");
    CU(c, cudaSetDevice(c->device));
    int fd[2] = {-1, -1};
    size_t size[2] = {0, 0};
    IoTrace tr;
    OKR(open_input(c, path_r1, &fd[0], &size[0]));
    int rc = open_input(c, path_r2, &fd[1], &size[1]);
    if (rc == SSD_OK) rc = ensure(c, c->b_in1, size[0] + 32);
    if (rc == SSD_OK) rc = ensure(c, c->b_in2, size[1] + 32);
    for (int k = 0; k < kIoLanes && rc == SSD_OK; k++) rc = c->io[k].init();
    tr.lap("open + buffers");
    if (rc == SSD_OK) {
        std::lock_guard<std::mutex> turn(g_h2d_turn);
        int rc2 = SSD_OK;
        RangeMap m1, m2;
        m1.open_range(fd[0], 0, size[0], false);
        m2.open_range(fd[1], 0, size[1], false);
        std::thread other([&] {
            rc2 = run_lanes(c, path_r2, kIoHalf, kIoHalf, upload_file_worker, c->device, fd[1], (const uint8_t*)m2.data, (size_t)0, size[1], (uint8_t*)c->b_in2.p);
        });
        rc = run_lanes(c, path_r1, 0, kIoHalf, upload_file_worker, c->device, fd[0], (const uint8_t*)m1.data, (size_t)0, size[0], (uint8_t*)c->b_in1.p);
        other.join();
        tr.lap("read + upload", size[0] + size[1]);
        m1.close_range_later();
        m2.close_range_later();
        if (rc == SSD_OK) rc = rc2;
    }
    for (int k = 0; k < 2; k++)
        if (fd[k] >= 0) close(fd[k]);
    tr.lap("unmap + close");
    if (rc != SSD_OK) return rc;
    OKR(ssd_demux_device(c, c->b_in1.p, size[0], c->b_in2.p, size[1]));
    OKR(ssd_filter_single(c, stats));
    tr.lap("kernels");
    return download_to_file(c, which, out_path, append, written);
}

extern "C" int ssd_annotated_file(ssd_ctx* c, const char* path_in, const char* out_path, int append, size_t* written, ssd_stats* stats)
{
    if (!c || !path_in || !out_path) return SSD_EINVAL;
    CU(c, cudaSetDevice(c->device));
    int fd = -1;
    size_t size = 0;
    OKR(open_input(c, path_in, &fd, &size));
    int rc = ensure(c, c->b_in1, size + 32);
    for (int k = 0; k < kIoLanes && rc == SSD_OK; k++) rc = c->io[k].init();
    if (rc == SSD_OK) {
        std::lock_guard<std::mutex> turn(g_h2d_turn);
        RangeMap m;
        m.open_range(fd, 0, size, false);
        rc = run_lanes(c, path_in, 0, kIoLanes, upload_file_worker, c->device, fd, (const uint8_t*)m.data, (size_t)0, size, (uint8_t*)c->b_in1.p);
        m.close_range_later();
    }
    close(fd);
    if (rc != SSD_OK) return rc;
    OKR(ssd_annotated_device(c, c->b_in1.p, size));
    OKR(ssd_filter_single(c, stats));
    return download_to_file(c, SSD_EMIT_PASSING, out_path, append, written);
}

// ---- the same in pieces, for inputs that have to go through the GPU in batches (ssd_b200.filebatch) ----
extern "C" int ssd_load_file_range(ssd_ctx* c, int slot, const char* path, uint64_t offset, uint64_t length)
{
    if (!c || !path || slot < 0 || slot > 1) return SSD_EINVAL;
    CU(c, cudaSetDevice(c->device));
    int fd = -1;
    size_t size = 0;
    OKR(open_input(c, path, &fd, &size));
    const size_t off = std::min<uint64_t>(offset, size), len = (size_t)std::min<uint64_t>(length, size - off);
    DevBuf& buf = slot == 0 ? c->b_in1 : c->b_in2;
    int rc = ensure(c, buf, len + 32);
    for (int k = 0; k < kIoLanes && rc == SSD_OK; k++) rc = c->io[k].init();
    if (rc == SSD_OK) {
        std::lock_guard<std::mutex> turn(g_h2d_turn);
        RangeMap m;
        m.open_range(fd, off, len, false);
        rc = run_lanes(c, path, 0, kIoLanes, upload_file_worker, c->device, fd, (const uint8_t*)m.data, off, len, (uint8_t*)buf.p);
        m.close_range_later();
    }
    close(fd);
    if (rc == SSD_OK) c->loaded[slot] = len;
    return rc;
}

extern "C" int ssd_anchor_positions(const void* d_text, size_t len, uint32_t n_lines, const char* const anchors[3], int32_t* found, uint32_t capacity,
                                    uint32_t* n_sequence_lines)
{
    if (!anchors || !found || !n_sequence_lines || n_lines == 0 || n_lines > (uint32_t)kLearnMaxLines || (len && !d_text)) return SSD_EINVAL;
    uint8_t h_anchor[48];
    int h_len[3];
    memset(h_anchor, 0, sizeof h_anchor);
    for (int k = 0; k < 3; k++) {
        if (!anchors[k] || strlen(anchors[k]) > 16) return SSD_EINVAL;
        h_len[k] = (int)strlen(anchors[k]);
        memcpy(h_anchor + 16 * k, anchors[k], (size_t)h_len[k]);
    }
    uint8_t* d_anchor = nullptr;
    int32_t* d_found = nullptr;
    const size_t found_bytes = (size_t)3 * capacity * sizeof(int32_t);
    if (cudaMalloc(&d_anchor, 64 + 16 + 16) != cudaSuccess) return SSD_ECUDA;
    if (cudaMalloc(&d_found, std::max<size_t>(found_bytes, 16)) != cudaSuccess) {
        cudaFree(d_anchor);
        return SSD_ECUDA;
    }
    int* d_len = (int*)(d_anchor + 64);
    uint32_t* d_nseq = (uint32_t*)(d_anchor + 80);
    bool ok = cudaMemcpy(d_anchor, h_anchor, sizeof h_anchor, cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMemcpy(d_len, h_len, sizeof h_len, cudaMemcpyHostToDevice) == cudaSuccess && cudaMemset(d_found, 0xFF, std::max<size_t>(found_bytes, 16)) == cudaSuccess;
    if (ok) {
        k_anchor_positions<<<1, 256>>>((const uint8_t*)d_text, (unsigned long long)len, n_lines, d_anchor, d_len, d_found, capacity, d_nseq);
        ok = cudaGetLastError() == cudaSuccess && cudaMemcpy(found, d_found, found_bytes, cudaMemcpyDeviceToHost) == cudaSuccess &&
             cudaMemcpy(n_sequence_lines, d_nseq, sizeof(uint32_t), cudaMemcpyDeviceToHost) == cudaSuccess;
    }
    cudaFree(d_anchor);
    cudaFree(d_found);
    return ok ? SSD_OK : SSD_ECUDA;
}

extern "C" int ssd_count_newlines_device(ssd_ctx* c, const void* d_text, size_t len, uint64_t* total)
{
    if (!c || !total || (len && !d_text)) return SSD_EINVAL;
    if ((uintptr_t)d_text & 15u) return fail(c, SSD_EINVAL, "ssd_count_newlines_device: the buffer must be 16-byte aligned");
    CU(c, cudaSetDevice(c->device));
    *total = 0;
    const int tile_log2 = 22;
    const uint64_t tiles = (len + (1ull << tile_log2) - 1) >> tile_log2;
    if (tiles == 0) return SSD_OK;
    OKR(ensure(c, c->b_first, tiles * sizeof(uint32_t)));
    k_count_newlines<<<(uint32_t)tiles, 256, 0, c->stream>>>((const uint8_t*)d_text, len, tile_log2, (uint32_t*)c->b_first.p);
    c->n_launches++;
    CU(c, cudaGetLastError());
    std::vector<uint32_t> h(tiles);
    CU(c, cudaMemcpyAsync(h.data(), c->b_first.p, tiles * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    for (uint32_t v : h) *total += v;
    return SSD_OK;
}

extern "C" int ssd_loaded_newlines(ssd_ctx* c, int slot, int tile_log2, uint32_t* tile_counts, uint64_t capacity, uint64_t* n_tiles)
{
    if (!c || slot < 0 || slot > 1 || tile_log2 < 12 || tile_log2 > 30 || !n_tiles) return SSD_EINVAL;
    CU(c, cudaSetDevice(c->device));
    const uint64_t len = c->loaded[slot], tiles = (len + (1ull << tile_log2) - 1) >> tile_log2;
    *n_tiles = tiles;
    if (tiles == 0) return SSD_OK;
    if (!tile_counts || capacity < tiles) return fail(c, SSD_ERANGE, "ssd_loaded_newlines: %llu tiles, room for %llu", (unsigned long long)tiles,
                                                      (unsigned long long)capacity);
    OKR(ensure(c, c->b_first, tiles * sizeof(uint32_t)));
    const DevBuf& buf = slot == 0 ? c->b_in1 : c->b_in2;
    k_count_newlines<<<(uint32_t)tiles, 256, 0, c->stream>>>((const uint8_t*)buf.p, len, tile_log2, (uint32_t*)c->b_first.p);
    c->n_launches++;
    CU(c, cudaGetLastError());
    CU(c, cudaMemcpyAsync(tile_counts, c->b_first.p, tiles * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return SSD_OK;
}

extern "C" int ssd_demux_loaded(ssd_ctx* c)
{
    if (!c) return SSD_EINVAL;
    if (!c->b_in1.p || !c->b_in2.p) return fail(c, SSD_ESTATE, "ssd_demux_loaded before ssd_load_file_range of both slots");
    return ssd_demux_device(c, c->b_in1.p, c->loaded[0], c->b_in2.p, c->loaded[1]);
}

extern "C" int ssd_annotated_loaded(ssd_ctx* c)
{
    if (!c) return SSD_EINVAL;
    if (!c->b_in1.p) return fail(c, SSD_ESTATE, "ssd_annotated_loaded before ssd_load_file_range of slot 0");
    return ssd_annotated_device(c, c->b_in1.p, c->loaded[0]);
}

extern "C" int ssd_emit_file(ssd_ctx* c, int which, const char* out_path, int append, size_t* written)
{
    if (!c || !out_path) return SSD_EINVAL;
    if (which != SSD_EMIT_MATCHED && which != SSD_EMIT_PASSING) return fail(c, SSD_EINVAL, "unknown output selector");
    if (!c->filter_done) return fail(c, SSD_ESTATE, "ssd_emit_file before the filter was built");
    CU(c, cudaSetDevice(c->device));
    return download_to_file(c, which, out_path, append, written);
}

// Where a FASTQ file can be cut into batches of whole records (host only; a line ends at '\n', a record is four lines).
//   records_in == NULL: cuts by size.  Entry 0 is (record 0, byte 0); after a cut at byte b the next one is the first record
//                       start at or after b + target_bytes; the last entry is (number of complete records, file size).
//   records_in != NULL: the byte offsets at which the given records (ascending indices) start; an index past the end of
//                       the file gives the file size.  This is how the mate file is cut at the same records.
// n_out receives the number of entries written (at most cap; SSD_ERANGE when cap is too small).
// The newlines of every 4 MiB block are counted once, on up to 16 host threads (pread + SWAR count); a cut is then located
// through the prefix sums of the block counts and a walk inside one block.
namespace {
constexpr size_t kCutBlock = (size_t)4 << 20;

inline uint64_t count_newlines(const uint8_t* p, size_t n)
{
    uint64_t k = 0;
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        uint64_t x;
        memcpy(&x, p + i, 8);
        x ^= 0x0A0A0A0A0A0A0A0Aull;
        const uint64_t z = ~(((x & 0x7F7F7F7F7F7F7F7Full) + 0x7F7F7F7F7F7F7F7Full) | x | 0x7F7F7F7F7F7F7F7Full);  // 0x80 where the byte was '\n'
        k += ((z >> 7) * 0x0101010101010101ull) >> 56;  // the eight 0/1 bytes summed into the top byte (no popcount instruction assumed)
    }
    for (; i < n; i++) k += p[i] == '\n';
    return k;
}

bool pread_all(int fd, uint8_t* dst, size_t want, uint64_t off)
{
    size_t got = 0;
    while (got < want) {
        const ssize_t r = pread(fd, dst + got, want - got, (off_t)(off + got));
        if (r < 0 && errno == EINTR) continue;
        if (r <= 0) return false;
        got += (size_t)r;
    }
    return true;
}

struct CutFile {
    int fd = -1;
    uint64_t size = 0, total = 0;       // bytes, newlines
    std::vector<uint64_t> before;       // newlines before block b (n_blocks + 1 entries)
    std::vector<uint8_t> blk;           // the block looked at last
    uint64_t blk_index = ~0ull;
    size_t blk_len = 0;
    bool ok = true;

    const uint8_t* block(uint64_t b)
    {
        if (b != blk_index) {
            blk_len = (size_t)std::min<uint64_t>(kCutBlock, size - b * kCutBlock);
            if (!pread_all(fd, blk.data(), blk_len, b * kCutBlock)) ok = false;
            blk_index = b;
        }
        return blk.data();
    }
    // a cursor inside the current block, so that a run of cuts in one block costs one pass over it: cur_lines = newlines in [0, cur_off)
    uint64_t cur_off = 0, cur_lines = 0;

    // newlines in [0, off) and whether byte off - 1 is one (0 < off <= size)
    uint64_t lines_before(uint64_t off, bool* ends_line)
    {
        const uint64_t b = (off - 1) / kCutBlock, lo = b * kCutBlock;
        const uint8_t* p = block(b);
        if (cur_off < lo || cur_off > off) {
            cur_off = lo;
            cur_lines = before[b];
        }
        cur_lines += count_newlines(p + (cur_off - lo), (size_t)(off - cur_off));
        cur_off = off;
        *ends_line = p[off - lo - 1] == '\n';
        return cur_lines;
    }
    // the byte offset right after newline number g (1 <= g <= total)
    uint64_t after_newline(uint64_t g)
    {
        const uint64_t b = (uint64_t)(std::lower_bound(before.begin(), before.end(), g) - before.begin()) - 1;  // before[b] < g <= before[b + 1]
        const uint64_t lo = b * kCutBlock;
        const uint8_t* p = block(b);
        if (cur_off < lo || cur_off > lo + blk_len || cur_lines >= g) {
            cur_off = lo;
            cur_lines = before[b];
        }
        size_t i = (size_t)(cur_off - lo);
        while (cur_lines < g) {
            const uint8_t* q = (const uint8_t*)memchr(p + i, '\n', blk_len - i);
            if (!q) {  // cannot happen: the block holds before[b + 1] - before[b] newlines
                ok = false;
                return size;
            }
            i = (size_t)(q - p) + 1;
            cur_lines++;
        }
        cur_off = lo + i;
        return cur_off;
    }
};
}  // namespace

extern "C" int ssd_fastq_record_cuts(const char* path, uint64_t target_bytes, const uint64_t* records_in, uint64_t n_in, uint64_t* records_out,
                                     uint64_t* offsets_out, uint64_t cap, uint64_t* n_out, uint64_t* n_records_total)
{
    if (!path || !records_out || !offsets_out || !n_out || (!records_in && target_bytes == 0)) return SSD_EINVAL;
    CutFile f;
    f.fd = open(path, O_RDONLY);
    struct stat st;
    if (f.fd < 0 || fstat(f.fd, &st) != 0) {
        if (f.fd >= 0) close(f.fd);
        return SSD_EIO;
    }
    f.size = (uint64_t)st.st_size;
    const uint64_t n_blocks = (f.size + kCutBlock - 1) / kCutBlock;
    f.before.assign(n_blocks + 1, 0);
    f.blk.resize(kCutBlock);
    // newlines per block, in parallel
    {
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        const int n_threads = (int)std::min<uint64_t>(std::min<unsigned>(hw, 16u), std::max<uint64_t>(n_blocks, 1));
        std::vector<std::thread> th;
        std::vector<int> bad(n_threads, 0);
        for (int t = 0; t < n_threads; t++)
            th.emplace_back([&, t] {
                std::vector<uint8_t> buf(kCutBlock);
                for (uint64_t b = (uint64_t)t; b < n_blocks; b += (uint64_t)n_threads) {
                    const size_t len = (size_t)std::min<uint64_t>(kCutBlock, f.size - b * kCutBlock);
                    if (!pread_all(f.fd, buf.data(), len, b * kCutBlock)) {
                        bad[t] = 1;
                        return;
                    }
                    f.before[b + 1] = count_newlines(buf.data(), len);
                }
            });
        for (auto& x : th) x.join();
        for (int t = 0; t < n_threads; t++)
            if (bad[t]) {
                close(f.fd);
                return SSD_EIO;
            }
    }
    for (uint64_t b = 0; b < n_blocks; b++) f.before[b + 1] += f.before[b];
    f.total = f.before[n_blocks];
    bool last_is_newline = true;
    if (f.size) f.lines_before(f.size, &last_is_newline);
    const uint64_t n_lines = f.total + (f.size && !last_is_newline ? 1 : 0);

    uint64_t n = 0;
    int rc = SSD_OK;
    auto put = [&](uint64_t rec, uint64_t off) {
        if (n >= cap) {
            rc = SSD_ERANGE;
            return;
        }
        records_out[n] = rec;
        offsets_out[n] = off;
        n++;
    };
    if (!records_in) {
        put(0, 0);
        uint64_t last_cut = 0;
        while (rc == SSD_OK && f.ok) {
            const uint64_t limit = last_cut + target_bytes;
            if (limit >= f.size || limit < last_cut) break;
            bool at_line_end = false;
            const uint64_t before = f.lines_before(limit, &at_line_end);
            uint64_t cut, g;
            if (before && (before & 3u) == 0 && at_line_end) {  // a record starts exactly at the limit
                g = before;
                cut = limit;
            } else {
                g = (before / 4 + 1) * 4;
                if (g > f.total) break;
                cut = f.after_newline(g);
            }
            if (cut >= f.size) break;
            put(g >> 2, cut);
            last_cut = cut;
        }
        if (rc == SSD_OK) put(n_lines >> 2, f.size);
    } else {
        for (uint64_t k = 0; k < n_in && rc == SSD_OK && f.ok; k++) {
            const uint64_t r = records_in[k];
            put(r, r == 0 ? 0 : (r > f.total / 4 ? f.size : f.after_newline(4 * r)));
        }
    }
    close(f.fd);
    *n_out = n;
    if (n_records_total) *n_records_total = n_lines >> 2;
    return f.ok ? rc : SSD_EIO;
}

// Host only: is every line of the file already what Python's line.strip() + "\n" would make of it?  That is the case when
// the file is ASCII, has no '\r' (text mode would translate it), no blank (str.isspace) next to a '\n' or at the very
// start, and ends with '\n' (or is empty).  The reference's splitter copies line.strip() + "\n" (LIB:36, 83); for such a
// file that is a plain copy of byte ranges.  *plain = 1 / 0; *n_lines = number of lines.
extern "C" int ssd_fastq_lines_are_stripped(const char* path, int* plain, uint64_t* n_lines)
{
    if (!path || !plain) return SSD_EINVAL;
    const int fd = open(path, O_RDONLY);
    struct stat st;
    if (fd < 0 || fstat(fd, &st) != 0) {
        if (fd >= 0) close(fd);
        return SSD_EIO;
    }
    const uint64_t size = (uint64_t)st.st_size;
    const uint64_t n_blocks = (size + kCutBlock - 1) / kCutBlock;
    uint8_t blank[256];
    for (int b = 0; b < 256; b++) blank[b] = (b == ' ' || (b >= 9 && b <= 13 && b != '\n') || (b >= 0x1c && b <= 0x1f)) ? 1 : 0;
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const int n_threads = (int)std::min<uint64_t>(std::min<unsigned>(hw, 16u), std::max<uint64_t>(n_blocks, 1));
    std::vector<int> verdict(n_threads, 1);  // 1 plain so far, 0 not plain, -1 read error
    std::vector<uint64_t> lines(n_threads, 0);
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++)
        th.emplace_back([&, t] {
            std::vector<uint8_t> buf(kCutBlock + 2);
            for (uint64_t b = (uint64_t)t; b < n_blocks && verdict[t] == 1; b += (uint64_t)n_threads) {
                // the block with one byte of context on either side (0 stands for "outside the file")
                const uint64_t lo = b * kCutBlock, len = std::min<uint64_t>(kCutBlock, size - lo);
                const uint64_t from = lo ? lo - 1 : 0, to = std::min<uint64_t>(size, lo + len + 1);
                uint8_t* p = buf.data() + (lo ? 0 : 1);
                buf[0] = 0;
                if (!pread_all(fd, p, (size_t)(to - from), from)) {
                    verdict[t] = -1;
                    return;
                }
                uint8_t* q = buf.data() + 1;  // q[i] = byte lo + i, q[-1] / q[len] = the neighbours
                if (to == lo + len) q[len] = 0;
                bool ok = true;
                uint64_t nl = 0, i = 0;
                auto one = [&](uint64_t k) {
                    const uint8_t c = q[k];
                    if (c >= 0x80 || c == '\r') ok = false;
                    if (c == '\n') {
                        nl++;
                        if (blank[q[k - 1]] || blank[q[k + 1]]) ok = false;
                    }
                };
                for (; i + 8 <= len; i += 8) {  // eight bytes at a time; only words that hold a newline are looked at bytewise
                    uint64_t x;
                    memcpy(&x, q + i, 8);
                    const uint64_t cr = x ^ 0x0D0D0D0D0D0D0D0Dull, lf = x ^ 0x0A0A0A0A0A0A0A0Aull;
                    if ((x & 0x8080808080808080ull) || ((cr - 0x0101010101010101ull) & ~cr & 0x8080808080808080ull)) ok = false;
                    if ((lf - 0x0101010101010101ull) & ~lf & 0x8080808080808080ull)
                        for (uint64_t k = i; k < i + 8; k++) one(k);
                }
                for (; i < len; i++) one(i);
                lines[t] += nl;
                if (!ok) verdict[t] = 0;
            }
        });
    for (auto& x : th) x.join();
    int v = 1;
    uint64_t total = 0;
    for (int t = 0; t < n_threads; t++) {
        if (verdict[t] < 0) v = -1;
        else if (verdict[t] == 0 && v == 1) v = 0;
        total += lines[t];
    }
    if (v == 1 && size) {
        uint8_t first = 0, last = 0;
        if (!pread_all(fd, &first, 1, 0) || !pread_all(fd, &last, 1, size - 1)) v = -1;
        else if (blank[first] || last != '\n') v = 0;
    }
    close(fd);
    if (v < 0) return SSD_EIO;
    *plain = v;
    if (n_lines) *n_lines = total;  // exact when plain (every line ends with its newline)
    return SSD_OK;
}

extern "C" int ssd_annotated_host(ssd_ctx* c, const void* text, size_t len, void* out, size_t out_cap, size_t* written, ssd_stats* stats)
{
    if (!c) return SSD_EINVAL;
    CU(c, cudaSetDevice(c->device));
    OKR(ensure(c, c->b_in1, len + 32));
    const size_t chunk = (size_t)64 << 20;
    for (size_t o = 0; o < len; o += chunk)
        CU(c, cudaMemcpyAsync((uint8_t*)c->b_in1.p + o, (const uint8_t*)text + o, std::min(chunk, len - o), cudaMemcpyHostToDevice, c->stream));
    OKR(ssd_annotated_device(c, c->b_in1.p, len));
    OKR(ssd_filter_single(c, stats));
    if (!out && !out_cap) {
        if (written) *written = (size_t)c->stats.bytes_passing;
        return SSD_OK;
    }
    return ssd_emit_host(c, SSD_EMIT_PASSING, out, out_cap, written);
}

extern "C" int ssd_record_cells(ssd_ctx* c, int32_t* cells_host, uint64_t capacity, uint64_t* n_records)
{
    if (!c) return SSD_EINVAL;
    if (!c->phase1_done) return fail(c, SSD_ESTATE, "ssd_record_cells before ssd_demux_device");
    if (n_records) *n_records = c->n_rec_r2;
    if (!cells_host) return SSD_OK;
    if (capacity < c->n_rec_r2) return fail(c, SSD_ERANGE, "need room for %llu records", (unsigned long long)c->n_rec_r2);
    if (!c->n_rec_r2) return SSD_OK;
    CU(c, cudaSetDevice(c->device));
    OKR(ensure(c, c->b_cells, c->n_rec_r2 * sizeof(int32_t)));
    c->n_launches++;
    k_record_cells<<<blocks_for(c->n_rec_r2, kThreads), kThreads, 0, c->stream>>>((const uint64_t*)c->b_ann.p, c->n_rec_r2, (int32_t*)c->b_cells.p);
    CU(c, cudaMemcpyAsync(cells_host, c->b_cells.p, c->n_rec_r2 * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return SSD_OK;
}

// ================================================================================================
// synthetic read pairs
// ================================================================================================
namespace {

struct SynthPrep {
    ssd_synth::Params p;
    std::vector<char> ties;
};

int synth_prepare(const ssd_synth_spec* s, SynthPrep& out)
{
    if (!s || s->struct_size != sizeof(ssd_synth_spec) || !s->whitelist || s->n_whitelist < 1 || s->n_whitelist > SSD_MAX_WHITELIST) return SSD_EINVAL;
    if (s->read_len < 94 || s->read_len > 1000 || s->cell_pool < 1) return SSD_EINVAL;
    ssd_synth::Params& p = out.p;
    p.seed = s->seed;
    p.read_len = s->read_len;
    p.n_wl = s->n_whitelist;
    p.cell_pool = s->cell_pool;
    p.cell_octaves = 1;
    while ((1ull << p.cell_octaves) <= s->cell_pool) p.cell_octaves++;
    p.p_sub = s->p_sub_ppm;
    p.p_n = s->p_n_ppm;
    p.p_junk = s->p_junk_ppm;
    p.p_dup = s->p_dup_umi_ppm;
    p.p_tie = s->p_tie_ppm;
    p.p_trunc = s->p_trunc_ppm;
    p.p_atqual = s->p_atqual_ppm;
    // midpoints of whitelist pairs at distance exactly 4: two of the differing positions taken from b
    out.ties.clear();
    for (int a = 0; a < s->n_whitelist && out.ties.size() < 8 * 512; a++)
        for (int b = a + 1; b < s->n_whitelist && out.ties.size() < 8 * 512; b++) {
            const char *x = s->whitelist + 8 * a, *y = s->whitelist + 8 * b;
            int diff[8], nd = 0;
            for (int j = 0; j < 8; j++)
                if (x[j] != y[j]) diff[nd++] = j;
            if (nd != 4) continue;
            char t[8];
            memcpy(t, x, 8);
            t[diff[0]] = y[diff[0]];
            t[diff[2]] = y[diff[2]];
            out.ties.insert(out.ties.end(), t, t + 8);
        }
    p.n_ties = (uint32_t)(out.ties.size() / 8);
    return SSD_OK;
}

struct SynthLen {
    ssd_synth::Params p;
    uint64_t first;
    int mate;
    __device__ __forceinline__ uint32_t operator()(uint64_t k) const
    {
        return mate == 1 ? ssd_synth::r1_record_len(p, first + k) : ssd_synth::r2_record_len(p, first + k);
    }
};

__global__ void __launch_bounds__(kThreads) k_synth_write(ssd_synth::Params p, uint64_t first, uint64_t n, const uint64_t* off1, const uint64_t* off2,
                                                          const char* wl, const char* ties, uint8_t* r1, uint8_t* r2)
{
    const int lane = threadIdx.x & 31;
    const uint64_t k = ((uint64_t)blockIdx.x * kThreads + threadIdx.x) >> 5;
    if (k >= n) return;
    const uint64_t i = first + k;
    {
        uint8_t* dst = r1 + off1[k];
        const uint32_t len = (uint32_t)(off1[k + 1] - off1[k]);
        for (uint32_t b = lane; b < len; b += 32) dst[b] = (uint8_t)ssd_synth::r1_byte(p, i, b);
    }
    {
        ssd_synth::R2Plan pl = ssd_synth::plan_r2(p, i, wl, ties);
        uint8_t* dst = r2 + off2[k];
        const uint32_t len = (uint32_t)(off2[k + 1] - off2[k]);
        for (uint32_t b = lane; b < len; b += 32) dst[b] = (uint8_t)ssd_synth::r2_byte(p, i, pl, b);
    }
}

}  // namespace

extern "C" int ssd_synth_sizes(const ssd_synth_spec* spec, uint64_t* r1_bytes, uint64_t* r2_bytes)
{
    SynthPrep sp;
    if (synth_prepare(spec, sp) != SSD_OK) return SSD_EINVAL;
    uint64_t a = 0, b = 0;
    for (uint64_t k = 0; k < spec->n_pairs; k++) {
        a += ssd_synth::r1_record_len(sp.p, spec->first_index + k);
        b += ssd_synth::r2_record_len(sp.p, spec->first_index + k);
    }
    if (r1_bytes) *r1_bytes = a;
    if (r2_bytes) *r2_bytes = b;
    return SSD_OK;
}

extern "C" int ssd_synth_generate_host(const ssd_synth_spec* spec, void* r1, void* r2)
{
    SynthPrep sp;
    if (synth_prepare(spec, sp) != SSD_OK || !r1 || !r2) return SSD_EINVAL;
    char *o1 = (char*)r1, *o2 = (char*)r2;
    for (uint64_t k = 0; k < spec->n_pairs; k++) {
        const uint64_t i = spec->first_index + k;
        uint32_t l1 = ssd_synth::r1_record_len(sp.p, i);
        for (uint32_t b = 0; b < l1; b++) *o1++ = ssd_synth::r1_byte(sp.p, i, b);
        ssd_synth::R2Plan pl = ssd_synth::plan_r2(sp.p, i, spec->whitelist, sp.ties.data());
        uint32_t l2 = ssd_synth::r2_record_len(sp.p, i);
        for (uint32_t b = 0; b < l2; b++) *o2++ = ssd_synth::r2_byte(sp.p, i, pl, b);
    }
    return SSD_OK;
}

extern "C" int ssd_synth_generate_device(ssd_ctx* c, const ssd_synth_spec* spec, void* d_r1, void* d_r2)
{
    if (!c) return SSD_EINVAL;
    SynthPrep sp;
    if (synth_prepare(spec, sp) != SSD_OK || !d_r1 || !d_r2) return fail(c, SSD_EINVAL, "bad synthetic spec");
    CU(c, cudaSetDevice(c->device));
    const uint64_t n = spec->n_pairs;
    if (!n) return SSD_OK;
    DevBuf off1, off2, tab;
    int rc = SSD_OK;
    do {
        if ((rc = ensure(c, off1, (n + 1) * 8)) != SSD_OK) break;
        if ((rc = ensure(c, off2, (n + 1) * 8)) != SSD_OK) break;
        const size_t wl_bytes = (size_t)spec->n_whitelist * 8, tie_bytes = sp.ties.size();
        if ((rc = ensure(c, tab, wl_bytes + tie_bytes + 16)) != SSD_OK) break;
        cudaMemcpyAsync(tab.p, spec->whitelist, wl_bytes, cudaMemcpyHostToDevice, c->stream);
        if (tie_bytes) cudaMemcpyAsync((char*)tab.p + wl_bytes, sp.ties.data(), tie_bytes, cudaMemcpyHostToDevice, c->stream);
        if ((rc = exclusive_scan<uint64_t>(c, SynthLen{sp.p, spec->first_index, 1}, n, (uint64_t*)off1.p)) != SSD_OK) break;
        if ((rc = exclusive_scan<uint64_t>(c, SynthLen{sp.p, spec->first_index, 2}, n, (uint64_t*)off2.p)) != SSD_OK) break;
        k_synth_write<<<blocks_for(n * 32, kThreads), kThreads, 0, c->stream>>>(sp.p, spec->first_index, n, (const uint64_t*)off1.p, (const uint64_t*)off2.p,
                                                                              (const char*)tab.p, (const char*)tab.p + wl_bytes, (uint8_t*)d_r1, (uint8_t*)d_r2);
        if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) rc = fail(c, SSD_ECUDA, "synthetic generation failed");
    } while (0);
    release(off1);
    release(off2);
    release(tab);
    return rc;
}

// ================================================================================================
// instrumentation
// ================================================================================================
extern "C" int ssd_set_timing(ssd_ctx* c, int enabled)
{
    if (!c) return SSD_EINVAL;
    c->timing = enabled != 0;
    return SSD_OK;
}

extern "C" int ssd_read_timings(ssd_ctx* c, double* ms_per_stage, uint32_t* launches_per_stage, int n_stages)
{
    if (!c) return SSD_EINVAL;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->stream));
    for (int i = 0; i < n_stages; i++) {
        if (ms_per_stage) ms_per_stage[i] = 0.0;
        if (launches_per_stage) launches_per_stage[i] = 0;
    }
    for (auto& sp : c->spans) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess && sp.stage < n_stages) {
            if (ms_per_stage) ms_per_stage[sp.stage] += ms;
            if (launches_per_stage) launches_per_stage[sp.stage]++;
        }
        cudaEventDestroy(sp.a);
        cudaEventDestroy(sp.b);
    }
    c->spans.clear();
    return SSD_OK;
}

extern "C" int ssd_scan_repeats(const ssd_ctx* c, uint64_t* n)
{
    if (!c || !n) return SSD_EINVAL;
    *n = c->n_respec;
    return SSD_OK;
}

extern "C" int ssd_scan_variant(const ssd_ctx* c, int* short_fast, uint64_t* exact_records)
{
    if (!c) return SSD_EINVAL;
    if (short_fast) *short_fast = c->short2 ? 1 : 0;
    if (exact_records) *exact_records = c->n_exact_last;
    return SSD_OK;
}

extern "C" uint64_t ssd_launch_count(const ssd_ctx* c) { return c ? c->n_launches : 0; }

#ifdef SSD_SCAN_PROFILE
// debug builds only: per-phase clock totals of the scan kernels (phases 0..4, tiles in [7]) for read 1 / read 2
extern "C" int ssd_debug_scan_profile(unsigned long long* out16, int reset)
{
    if (out16 && cudaMemcpyFromSymbol(out16, ssd::g_scan_prof, sizeof(unsigned long long) * 32) != cudaSuccess) return SSD_ECUDA;
    if (reset) {
        unsigned long long z[32] = {0};
        if (cudaMemcpyToSymbol(ssd::g_scan_prof, z, sizeof z) != cudaSuccess) return SSD_ECUDA;
    }
    return SSD_OK;
}
#endif
