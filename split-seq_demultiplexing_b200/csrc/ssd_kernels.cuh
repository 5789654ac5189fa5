// ssd_kernels.cuh — the CUDA kernels of the --version fast hot path (sm_100a).
//
//   k_scan<R1|R2>   single-pass FASTQ record-boundary scan: a tile of the file is bulk-copied (TMA) into
//                   shared memory, newlines become a bitmap (16-byte vector loads), per-tile newline counts
//                   are chained with a decoupled look-back so every tile learns its global line index, and
//                   the records that start in the tile are processed straight from shared memory:
//                     R1: record offset + the length its output record will have (V2:134-141)
//                     R2: UMI / barcode extraction (V2:293-301), first-hit Hamming matching (V2:302-304,
//                         327-329), mate-id check (V2:350), per-cell histogram, (cell,UMI) keys
//   k_build_lut     first-hit whitelist index for every 16-bit 2-bit-packed 8-mer
//   radix sort / unique / per-cell distinct count      the len(set(UMIs)) of V2:411
//   k_filter        per-cell pass decision + optional RanHex->OligoDT collapse map
//   k_emit          prefix-sum-compacted writer of the annotated FASTQ (V2:134-141, 367-370, 415-454)
#pragma once
#include "ssd_device.cuh"

namespace ssd {

constexpr int kTile = 32768;              // bytes of the file owned by one scan block
constexpr int kHalo = 2048;               // extra bytes loaded so records starting in the tile are whole
constexpr int kWin = kTile + kHalo;
constexpr int kWinWords = kWin / 32;
constexpr int kTileWords = kTile / 32;
constexpr int kMaxRecTile = 2048;         // records starting in one tile (>= 16 bytes each on average)

constexpr unsigned long long kFlagAgg = 1ull << 62;
constexpr unsigned long long kFlagPrefix = 2ull << 62;
constexpr unsigned long long kStatusValue = (1ull << 62) - 1;

template <typename OffT>
struct ScanArgs {
    const uint8_t* buf;        // file being scanned
    size_t len;
    unsigned long long* status;  // one word per tile: flag << 62 | newline count (aggregate or inclusive prefix)
    uint32_t n_tiles;
    DevCounters* cnt;
    unsigned long long* err_off;  // [2] smallest offending byte offset per file
    uint64_t rec_cap;
    // read 1 outputs
    OffT* r1_start;
    uint32_t* r1_base;
    // read 2 inputs / outputs
    const uint8_t* r1buf;
    size_t r1_len;
    uint64_t* ann;
    uint32_t* hist;
    uint64_t* keys;
    IrrKey* irr;
    const uint8_t* lut;
};

__device__ __forceinline__ void raise(DevCounters* cnt, unsigned long long* err_off, int file, uint32_t flag, size_t off)
{
    atomicOr(&cnt->err_flags, flag);
    atomicMin(&err_off[file], (unsigned long long)off);
}

// ---- read 1: where the record starts and how long its output record is going to be ----------------
template <class Acc, typename OffT>
__device__ __forceinline__ bool r1_record(const Acc& a, const ScanArgs<OffT>& args, uint64_t r, size_t start)
{
    Lines L;
    if (!find_lines(a, start, args.len, true, L)) return false;
    if (!L.complete) return true;
    if (r >= args.rec_cap) {
        atomicOr(&args.cnt->err_flags, kErrOverflow);
        return true;
    }
    if (a.byte(L.s[0]) != '@') raise(args.cnt, args.err_off, 0, kErrNoAt, start);
    // name.split("/")[0].replace(" ", "").strip()  (V2:135-138): count the bytes that survive
    uint32_t kept = 0, kept_at_last_nonspace = 0;
    for (size_t p = L.s[0]; p < L.e[0]; p++) {
        uint32_t c = a.byte(p);
        if (c == '/') break;
        if (c != ' ') kept++;
        if (!py_space(c)) kept_at_last_nonspace = kept;
    }
    size_t seq_e = rstrip_end(a, L.s[1], L.e[1]);
    size_t qual_e = rstrip_end(a, L.s[3], L.e[3]);
    // header' + '_' + barcodes + '_' + umi + '\n' + seq + "\n+\n" + qual + '\n'  -> 7 fixed bytes here,
    // barcode and UMI lengths are added once the read-2 annotation is known
    args.r1_start[r] = (OffT)start;
    args.r1_base[r] = kept_at_last_nonspace + (uint32_t)(seq_e - L.s[1]) + (uint32_t)(qual_e - L.s[3]) + 7u;
    return true;
}

// ---- read 2: extract, match, check the mate, count --------------------------------------------------
struct R2Out {
    uint64_t ann;      // 0 = rejected
    uint64_t key;      // regular key
    IrrKey irr;
};

template <class Acc, typename OffT>
__device__ __forceinline__ bool r2_record(const Acc& a, const ScanArgs<OffT>& args, const DevCfg& cfg, uint64_t r, size_t start,
                                          R2Out& out, bool& complete)
{
    Lines L;
    out.ann = 0;
    complete = false;
    if (!find_lines(a, start, args.len, false, L)) return false;
    if (!L.complete) return true;
    complete = true;
    if (r >= args.rec_cap) {
        atomicOr(&args.cnt->err_flags, kErrOverflow);
        complete = false;
        return true;
    }
    if (a.byte(L.s[0]) != '@') raise(args.cnt, args.err_off, 1, kErrNoAt, start);
    const size_t seq_s = L.s[1];
    const size_t seq_len = rstrip_end(a, seq_s, L.e[1]) - seq_s;  // lineRead = line.rstrip()  (V2:293)

    uint8_t raw[8];
    Packed p1 = pack_window(a, seq_s, seq_len, cfg.b1s, cfg.b1e, raw);
    int i1 = match_window(cfg, p1, raw, args.lut);
    if (i1 < 0) return true;  // V2:310-313
    Packed p2 = pack_window(a, seq_s, seq_len, cfg.b2s, cfg.b2e, raw);
    int i2 = match_window(cfg, p2, raw, args.lut);
    if (i2 < 0) return true;  // V2:314-317
    Packed p3 = pack_window(a, seq_s, seq_len, cfg.b3s, cfg.b3e, raw);
    int i3 = match_window(cfg, p3, raw, args.lut);
    if (i3 < 0) return true;  // V2:318-321
    const uint32_t cell = (uint32_t)((i1 * cfg.n_wl + i2) * cfg.n_wl + i3);

    // mate join by read id (V2:349-350): the id is the first whitespace-delimited token of the header
    {
        bool ok = r < args.cnt->n_records[0];
        if (ok) {
            size_t tok_e = L.s[0];
            while (tok_e < L.e[0] && !py_space(a.byte(tok_e))) tok_e++;
            size_t o1 = (size_t)args.r1_start[r];
            for (size_t j = 0;; j++) {
                bool e2 = L.s[0] + j >= tok_e;
                uint32_t c1 = o1 + j < args.r1_len ? args.r1buf[o1 + j] : (uint32_t)'\n';
                bool e1 = py_space(c1);
                if (e1 || e2) {
                    ok = e1 && e2;
                    break;
                }
                if (c1 != a.byte(L.s[0] + j)) {
                    ok = false;
                    break;
                }
            }
        }
        if (!ok) {
            raise(args.cnt, args.err_off, 1, kErrKey, start);
            return true;
        }
    }

    // UMI = lineRead[umi_start:umi_end] verbatim (V2:295); regular = full length and ACGT only
    size_t ulo = (size_t)cfg.us < seq_len ? (size_t)cfg.us : seq_len;
    size_t uhi = (size_t)cfg.ue < seq_len ? (size_t)cfg.ue : seq_len;
    int ulen = uhi > ulo ? (int)(uhi - ulo) : 0;
    uint64_t code = 0, hi = 0, lo = 0;
    bool regular = ulen == cfg.umi_len;
    for (int j = 0; j < ulen; j++) {
        uint32_t c = a.byte(seq_s + ulo + j);
        regular = regular && is_acgt(c);
        code |= (uint64_t)base2(c) << (2 * j);
        if (j < 4) hi |= (uint64_t)c << (8 * (3 - j));
        else lo |= (uint64_t)c << (8 * (9 - j));
    }
    if (regular) {
        out.ann = kAnnRegular | ((uint64_t)cell << kAnnCellShift) | code;
        out.key = ((uint64_t)cell << cfg.umi_bits) | code;
    } else {
        out.ann = kAnnIrregular | ((uint64_t)cell << kAnnCellShift) | ((uint64_t)(seq_s + ulo) << 5) | (uint64_t)ulen;
        out.irr.hi = ((uint64_t)cell << kIrrCellShift) | ((uint64_t)ulen << 32) | hi;
        out.irr.lo = lo;
    }
    return true;
}

constexpr int kScanR1 = 0;
constexpr int kScanR2 = 1;

template <int MODE, typename OffT>
__global__ void __launch_bounds__(kThreads) k_scan(const __grid_constant__ ScanArgs<OffT> args, const __grid_constant__ DevCfg cfg)
{
    __shared__ __align__(128) uint8_t s_win[kWin];
    __shared__ __align__(16) uint32_t s_mask[kWinWords];
    __shared__ uint16_t s_wpre[kTileWords];
    __shared__ uint16_t s_rec[kMaxRecTile];
    __shared__ uint32_t s_warp[kThreads / 32];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ unsigned int s_tile;
    __shared__ unsigned long long s_line_base;
    __shared__ unsigned int s_nk, s_ni;
    __shared__ unsigned long long s_kbase, s_ibase;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        s_tile = atomicAdd(&args.cnt->ticket[MODE], 1u);  // tiles are claimed in order: look-back never waits on an unstarted tile
        mbar_init(&s_bar, 1);
    }
    __syncthreads();
    const uint32_t tile = s_tile;
    const size_t tile_lo = (size_t)tile * kTile;
    const uint32_t win_bytes = (uint32_t)(args.len - tile_lo < (size_t)kWin ? args.len - tile_lo : (size_t)kWin);

    load_window(s_win, args.buf + tile_lo, win_bytes, &s_bar, 0);
    build_nl_mask(s_win, win_bytes, reinterpret_cast<uint16_t*>(s_mask), kWinWords * 2);
    __syncthreads();

    // ---- newlines owned by this tile: words [0, kTileWords) ----
    uint32_t m[4], cnt = 0;
#pragma unroll
    for (int w = 0; w < 4; w++) {
        m[w] = s_mask[4 * tid + w];
        cnt += __popc(m[w]);
    }
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t warp_base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; w++) {
        uint32_t v = s_warp[w];
        if (w < warp) warp_base += v;
        total += v;
    }
    const uint32_t excl = warp_base + incl - cnt;
    {
        uint32_t run = excl;
#pragma unroll
        for (int w = 0; w < 4; w++) {
            s_wpre[4 * tid + w] = (uint16_t)run;
            run += __popc(m[w]);
        }
    }

    // ---- decoupled look-back over the tile newline counts ----
    if (tid == 0) {
        volatile unsigned long long* st = args.status;
        unsigned long long excl_lines = 0;
        if (tile == 0) {
            st[0] = kFlagPrefix | total;
        } else {
            st[tile] = kFlagAgg | total;
            for (long long look = (long long)tile - 1;; look--) {
                unsigned long long v;
                do {
                    v = st[look];
                } while ((v >> 62) == 0);
                excl_lines += v & kStatusValue;
                if ((v >> 62) == 2) break;
            }
            st[tile] = kFlagPrefix | (excl_lines + total);
        }
        s_line_base = excl_lines;
        s_nk = 0;
        s_ni = 0;
        if (tile == args.n_tiles - 1) {
            unsigned long long n_lines = excl_lines + total + ((args.len && s_win[win_bytes - 1] != '\n') ? 1ull : 0ull);
            args.cnt->n_lines[MODE] = n_lines;
            args.cnt->n_records[MODE] = n_lines >> 2;
        }
    }
    __syncthreads();
    const unsigned long long line_base = s_line_base;

    // ---- records that start in this tile: the line after a newline whose next-line index is 0 mod 4 ----
    const unsigned long long r_first = tile == 0 ? 0ull : (line_base + 4) >> 2;
    const unsigned long long r_end = ((line_base + total) >> 2) + 1;  // exclusive
    uint32_t n_rec = r_end > r_first ? (uint32_t)(r_end - r_first) : 0u;
    if (tile == 0 && args.len == 0) n_rec = 0;
    {
        uint32_t rank = excl;
#pragma unroll
        for (int w = 0; w < 4; w++) {
            uint32_t mm = m[w];
            while (mm) {
                uint32_t b = (uint32_t)__ffs((int)mm) - 1u;
                mm &= mm - 1u;
                unsigned long long Lidx = line_base + rank + 1;
                rank++;
                if ((Lidx & 3ull) == 0) {
                    unsigned long long idx = (Lidx >> 2) - r_first;
                    if (idx < (unsigned long long)kMaxRecTile) s_rec[idx] = (uint16_t)((4 * tid + w) * 32 + b + 1);
                }
            }
        }
        if (tile == 0 && tid == 0) s_rec[0] = 0;
    }
    if (n_rec > (uint32_t)kMaxRecTile) {
        if (tid == 0) atomicOr(&args.cnt->err_flags, kErrDense);
        n_rec = kMaxRecTile;
    }
    __syncthreads();

    WinAcc wa{s_win, s_mask, tile_lo, win_bytes, args.len};
    GlobAcc ga{args.buf, args.len};

    if (MODE == kScanR1) {
        for (uint32_t i = tid; i < n_rec; i += kThreads) {
            const uint64_t r = r_first + i;
            const size_t start = tile_lo + s_rec[i];
            if (!r1_record(wa, args, r, start)) r1_record(ga, args, r, start);
        }
    } else {
        // batches of kThreads records so the key appends can be aggregated per block
        for (uint32_t base = 0; base < n_rec; base += kThreads) {
            const uint32_t i = base + tid;
            R2Out o;
            o.ann = 0;
            bool complete = false;
            uint32_t slot = 0;
            const uint64_t r = r_first + i;
            if (i < n_rec) {
                const size_t start = tile_lo + s_rec[i];
                if (!r2_record(wa, args, cfg, r, start, o, complete)) r2_record(ga, args, cfg, r, start, o, complete);
                if (complete) args.ann[r] = o.ann;
                if (ann_state(o.ann) == 1) slot = atomicAdd(&s_nk, 1u);
                else if (ann_state(o.ann) == 2) slot = atomicAdd(&s_ni, 1u);
                if (o.ann) atomicAdd(&args.hist[ann_cell(o.ann)], 1u);
            }
            __syncthreads();
            if (tid == 0) {
                s_kbase = s_nk ? atomicAdd(&args.cnt->n_keys, (unsigned long long)s_nk) : 0ull;
                s_ibase = s_ni ? atomicAdd(&args.cnt->n_irr, (unsigned long long)s_ni) : 0ull;
            }
            __syncthreads();
            if (ann_state(o.ann) == 1) args.keys[s_kbase + slot] = o.key;
            else if (ann_state(o.ann) == 2) args.irr[s_ibase + slot] = o.irr;
            __syncthreads();
            if (tid == 0) {
                s_nk = 0;
                s_ni = 0;
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// first-hit LUT: lut[e'][code] for e' = 0..e
// ---------------------------------------------------------------------------------------------
__global__ void k_build_lut(const __grid_constant__ DevCfg cfg, uint8_t* lut, int n_levels)
{
    uint32_t code = blockIdx.x * blockDim.x + threadIdx.x;
    if (code >= 65536u) return;
    for (int lev = 0; lev < n_levels; lev++) {
        uint32_t hit = 0xFFu;
        for (int i = 0; i < cfg.n_wl; i++) {
            uint32_t x = code ^ cfg.wl_code[i];
            if (__popc((x | (x >> 1)) & 0x5555u) <= lev) {
                hit = (uint32_t)i;
                break;
            }
        }
        lut[(size_t)lev * 65536u + code] = (uint8_t)hit;
    }
}

// ---------------------------------------------------------------------------------------------
// device-wide exclusive scan: out[i] = sum_{j<i} f(j), out[n] = total   (3 kernels)
// ---------------------------------------------------------------------------------------------
constexpr int kScanItems = 8;
constexpr int kScanBlock = kThreads * kScanItems;

template <typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T* s_tmp, T& total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T u = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += u;
    }
    if (lane == 31) s_tmp[warp] = incl;
    __syncthreads();
    T base = 0;
    total = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; w++) {
        T u = s_tmp[w];
        if (w < warp) base += u;
        total += u;
    }
    __syncthreads();
    return base + incl - v;
}

template <typename T, class F>
__global__ void __launch_bounds__(kThreads) k_scan_reduce(F f, uint64_t n, T* block_sums)
{
    __shared__ T s_tmp[kThreads / 32];
    const uint64_t base = (uint64_t)blockIdx.x * kScanBlock + (uint64_t)threadIdx.x * kScanItems;
    T v = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++)
        if (base + k < n) v += (T)f(base + k);
    T total;
    block_exclusive_scan(v, s_tmp, total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

template <typename T>
__global__ void __launch_bounds__(kThreads) k_scan_block_sums(T* block_sums, uint32_t n_blocks, T* total_out)
{
    __shared__ T s_tmp[kThreads / 32];
    __shared__ T s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_blocks; base += kThreads) {
        uint32_t i = base + threadIdx.x;
        T v = i < n_blocks ? block_sums[i] : (T)0;
        T total;
        T ex = block_exclusive_scan(v, s_tmp, total);
        T carry = s_carry;
        if (i < n_blocks) block_sums[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = s_carry;
}

template <typename T, class F>
__global__ void __launch_bounds__(kThreads) k_scan_apply(F f, uint64_t n, const T* block_sums, T* out)
{
    __shared__ T s_tmp[kThreads / 32];
    const uint64_t base = (uint64_t)blockIdx.x * kScanBlock + (uint64_t)threadIdx.x * kScanItems;
    T item[kScanItems];
    T v = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        item[k] = base + k < n ? (T)f(base + k) : (T)0;
        v += item[k];
    }
    T total;
    T ex = block_exclusive_scan(v, s_tmp, total) + block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        if (base + k < n) out[base + k] = ex;
        ex += item[k];
    }
    if (base <= n && n < base + kScanItems) out[n] = ex;  // items past n are zero, so this is the grand total
}

// ---------------------------------------------------------------------------------------------
// LSD radix sort, 8-bit digits, stable.  Pass = per-block digit histogram -> scan -> ranked scatter.
// ---------------------------------------------------------------------------------------------
constexpr int kSortItems = 16;
constexpr int kSortBlock = kThreads * kSortItems;  // 4096 keys per block

template <typename K>
__device__ __forceinline__ uint32_t digit_of(const K& k, int bit);
template <>
__device__ __forceinline__ uint32_t digit_of<uint64_t>(const uint64_t& k, int bit)
{
    return (uint32_t)(k >> bit) & 0xFFu;
}

template <typename K>
__global__ void __launch_bounds__(kThreads) k_sort_hist(const K* keys, uint64_t n, int bit, uint32_t* counts, uint32_t n_blocks)
{
    __shared__ uint32_t s_hist[256];
    s_hist[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * kSortBlock;
#pragma unroll
    for (int i = 0; i < kSortItems; i++) {
        uint64_t j = base + (uint64_t)i * kThreads + threadIdx.x;
        if (j < n) atomicAdd(&s_hist[digit_of(keys[j], bit)], 1u);
    }
    __syncthreads();
    counts[(size_t)threadIdx.x * n_blocks + blockIdx.x] = s_hist[threadIdx.x];
}

template <typename K, bool HAS_VALS>
__global__ void __launch_bounds__(kThreads) k_sort_scatter(const K* keys_in, K* keys_out, const uint32_t* vals_in, uint32_t* vals_out,
                                                           uint64_t n, int bit, const uint32_t* offsets, uint32_t n_blocks)
{
    __shared__ uint32_t s_wcount[kThreads / 32][257];
    __shared__ uint32_t s_base[256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < (kThreads / 32) * 257; i += kThreads) (&s_wcount[0][0])[i] = 0;
    __syncthreads();

    // warp w owns keys [w*512, (w+1)*512) of the block; iteration i, lane l -> key w*512 + i*32 + l
    const uint64_t wbase = (uint64_t)blockIdx.x * kSortBlock + (uint64_t)warp * (32 * kSortItems);
    K key[kSortItems];
    uint32_t val[kSortItems];
    uint16_t rank[kSortItems];
    uint16_t dig[kSortItems];
#pragma unroll
    for (int i = 0; i < kSortItems; i++) {
        uint64_t j = wbase + (uint64_t)i * 32 + lane;
        bool valid = j < n;
        if (valid) {
            key[i] = keys_in[j];
            if (HAS_VALS) val[i] = vals_in[j];
        }
        uint32_t d = valid ? digit_of(key[i], bit) : 256u;
        uint32_t peers = __match_any_sync(0xFFFFFFFFu, d);
        uint32_t leader = (uint32_t)__ffs((int)peers) - 1u;
        uint32_t pre = 0;
        if ((uint32_t)lane == leader) pre = s_wcount[warp][d];
        pre = __shfl_sync(0xFFFFFFFFu, pre, (int)leader);
        __syncwarp();
        if ((uint32_t)lane == leader) s_wcount[warp][d] = pre + (uint32_t)__popc(peers);
        __syncwarp();
        rank[i] = (uint16_t)(pre + (uint32_t)__popc(peers & ((1u << lane) - 1u)));
        dig[i] = (uint16_t)d;
    }
    __syncthreads();
    {
        // exclusive scan over warps for digit `tid`, plus the global base of (digit, block)
        uint32_t acc = 0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; w++) {
            uint32_t c = s_wcount[w][tid];
            s_wcount[w][tid] = acc;
            acc += c;
        }
        s_base[tid] = offsets[(size_t)tid * n_blocks + blockIdx.x];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kSortItems; i++) {
        uint32_t d = dig[i];
        if (d < 256u) {
            uint64_t dst = (uint64_t)s_base[d] + s_wcount[warp][d] + rank[i];
            keys_out[dst] = key[i];
            if (HAS_VALS) vals_out[dst] = val[i];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// unique compaction of a sorted array + per-cell distinct counts
// ---------------------------------------------------------------------------------------------
struct HeadFlag64 {
    const uint64_t* k;
    __device__ __forceinline__ uint32_t operator()(uint64_t i) const { return i == 0 || k[i] != k[i - 1] ? 1u : 0u; }
};

struct HeadFlagIrr {  // irregular keys are sorted through an index permutation
    const IrrKey* k;
    const uint32_t* perm;
    __device__ __forceinline__ uint32_t operator()(uint64_t i) const
    {
        if (i == 0) return 1u;
        IrrKey a = k[perm[i]], b = k[perm[i - 1]];
        return a.hi != b.hi || a.lo != b.lo ? 1u : 0u;
    }
};

__global__ void k_compact64(const uint64_t* in, const uint32_t* pos, uint64_t n, uint64_t* out)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i == 0 || in[i] != in[i - 1]) out[pos[i]] = in[i];
}

__global__ void k_compact_irr(const IrrKey* in, const uint32_t* perm, const uint32_t* pos, uint64_t n, IrrKey* out)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    IrrKey a = in[perm[i]];
    bool head = i == 0;
    if (!head) {
        IrrKey b = in[perm[i - 1]];
        head = a.hi != b.hi || a.lo != b.lo;
    }
    if (head) out[pos[i]] = a;
}

// distinct[cell] += (index of the cell's last unique key) + 1 - (index of its first): two atomics per cell run
__global__ void k_count_cells64(const uint64_t* uniq, uint64_t n, int shift, uint32_t* distinct)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t c = (uint32_t)(uniq[i] >> shift);
    if (i == 0 || (uint32_t)(uniq[i - 1] >> shift) != c) atomicSub(&distinct[c], (uint32_t)i);
    if (i == n - 1 || (uint32_t)(uniq[i + 1] >> shift) != c) atomicAdd(&distinct[c], (uint32_t)i + 1u);
}

__global__ void k_count_cells_irr(const IrrKey* uniq, uint64_t n, uint32_t* distinct)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t c = (uint32_t)(uniq[i].hi >> kIrrCellShift);
    if (i == 0 || (uint32_t)(uniq[i - 1].hi >> kIrrCellShift) != c) atomicSub(&distinct[c], (uint32_t)i);
    if (i == n - 1 || (uint32_t)(uniq[i + 1].hi >> kIrrCellShift) != c) atomicAdd(&distinct[c], (uint32_t)i + 1u);
}

__global__ void k_iota(uint32_t* p, uint64_t n)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (uint32_t)i;
}

__global__ void k_gather_irr_word(const IrrKey* in, const uint32_t* perm, uint64_t n, int which, uint64_t* out)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = which ? in[perm[i]].hi : in[perm[i]].lo;
}

// ---------------------------------------------------------------------------------------------
// the -m decision (V2:409-412) and the collapse map (Collapse_RanHex_Odt.py:44-71)
// ---------------------------------------------------------------------------------------------
__global__ void k_filter(const uint32_t* hist, const uint32_t* distinct, uint64_t n_cells, long long min_count, int mode, uint8_t* pass,
                         DevCounters* cnt)
{
    uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool present = false, ok = false;
    if (c < n_cells) {
        uint32_t reads = hist[c];
        present = reads != 0;
        long long v = mode == 1 ? (long long)reads : (long long)distinct[c];
        ok = present && v >= min_count;
        pass[c] = ok ? 1 : 0;
    }
    uint32_t b1 = __ballot_sync(0xFFFFFFFFu, present), b2 = __ballot_sync(0xFFFFFFFFu, ok);
    if ((threadIdx.x & 31) == 0) {
        if (b1) atomicAdd(&cnt->n_cells, (unsigned long long)__popc(b1));
        if (b2) atomicAdd(&cnt->n_passing_cells, (unsigned long long)__popc(b2));
    }
}

__global__ void k_collapse_map(const uint8_t* pass, int n, int pairs, uint32_t* remap)
{
    uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t total = (uint64_t)n * n * n;
    if (c >= total) return;
    uint32_t i1 = (uint32_t)(c / ((uint64_t)n * n));
    uint32_t out = (uint32_t)c;
    if (i1 >= (uint32_t)pairs && i1 < 2u * (uint32_t)pairs && pass[c]) {
        uint64_t odt = c - (uint64_t)pairs * n * n;
        if (pass[odt]) out = (uint32_t)odt;
    }
    remap[c] = out;
}

// ---------------------------------------------------------------------------------------------
// output planning: length of each record's output (0 when not emitted)
// ---------------------------------------------------------------------------------------------
struct OutLen {
    const uint64_t* ann;
    const uint32_t* r1_base;
    const uint8_t* pass;   // NULL = every matched record (MergedCells_1)
    const uint32_t* remap; // NULL = no collapse
    const uint8_t* wl_len;  // device copy of per-entry lengths
    int n, umi_len;
    __device__ __forceinline__ uint32_t operator()(uint64_t r) const
    {
        uint64_t a = ann[r];
        if (ann_state(a) == 0) return 0u;
        uint32_t cell = ann_cell(a);
        if (pass && !pass[cell]) return 0u;
        if (remap) cell = remap[cell];
        uint32_t i3 = cell % (uint32_t)n, i2 = (cell / (uint32_t)n) % (uint32_t)n, i1 = cell / ((uint32_t)n * (uint32_t)n);
        return r1_base[r] + wl_len[i1] + wl_len[i2] + wl_len[i3] + ann_umi_len(a, umi_len);
    }
};

__global__ void k_plan_totals(OutLen matched, OutLen passing, uint64_t n, DevCounters* cnt)
{
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long a = 0, b = 0, k = 0, mt = 0;
    if (r < n) {
        a = matched(r);
        b = passing(r);
        k = b ? 1 : 0;
        mt = a ? 1 : 0;
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        a += __shfl_xor_sync(0xFFFFFFFFu, a, d);
        b += __shfl_xor_sync(0xFFFFFFFFu, b, d);
        k += __shfl_xor_sync(0xFFFFFFFFu, k, d);
        mt += __shfl_xor_sync(0xFFFFFFFFu, mt, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (a) atomicAdd(&cnt->bytes_matched, a);
        if (b) atomicAdd(&cnt->bytes_passing, b);
        if (k) atomicAdd(&cnt->n_retained, k);
        if (mt) atomicAdd(&cnt->n_matched, mt);
    }
}

// ---------------------------------------------------------------------------------------------
// the writer
// ---------------------------------------------------------------------------------------------
constexpr int kEmitIn = 20480;   // staged input bytes per block
constexpr int kEmitOut = 20480;  // staged output bytes per block
constexpr int kEmitInWords = (kEmitIn + 32) / 32 + 1;

template <typename OffT>
struct EmitArgs {
    const uint8_t* r1buf;
    size_t r1_len;
    const uint8_t* r2buf;
    size_t r2_len;
    const OffT* r1_start;
    const uint64_t* ann;
    const uint64_t* out_off;  // n_rec + 1 entries
    const uint32_t* remap;    // NULL = none
    uint8_t* out;
    uint64_t n_rec;           // records of this batch (min of the two files)
    uint64_t n_rec_r1;        // records in read 1 (its table has this many entries)
    uint32_t rpb;             // records per block
    DevCounters* cnt;
};

// One warp writes one output record (V2:134-141).  `dst` is a generic pointer (shared staging buffer or
// global memory) to the first byte of the record's output; returns the number of bytes written.
template <class Acc, typename OffT>
__device__ __forceinline__ uint32_t emit_record(const Acc& a, const EmitArgs<OffT>& args, const DevCfg& cfg, uint64_t r, uint8_t* dst)
{
    const int lane = threadIdx.x & 31;
    const size_t start = (size_t)args.r1_start[r];
    Lines L;
    find_lines(a, start, args.r1_len, true, L);
    uint32_t o = 0;

    // header: bytes before the first '/', blanks removed, then strip()
    size_t cut = L.e[0];
    for (size_t p = L.s[0]; p < L.e[0]; p += 32) {
        uint32_t c = p + lane < L.e[0] ? a.byte(p + lane) : 0u;
        uint32_t sl = __ballot_sync(0xFFFFFFFFu, c == '/');
        if (sl) {
            cut = p + (uint32_t)__ffs((int)sl) - 1u;
            break;
        }
    }
    size_t end2 = L.s[0];
    for (size_t hi = cut; hi > L.s[0];) {
        size_t lo = hi - L.s[0] > 32 ? hi - 32 : L.s[0];
        uint32_t c = lo + lane < hi ? a.byte(lo + lane) : (uint32_t)' ';
        uint32_t ns = __ballot_sync(0xFFFFFFFFu, !py_space(c));
        if (ns) {
            end2 = lo + (32u - (uint32_t)__clz((int)ns));
            break;
        }
        hi = lo;
    }
    for (size_t p = L.s[0]; p < end2; p += 32) {
        uint32_t c = p + lane < end2 ? a.byte(p + lane) : (uint32_t)' ';
        uint32_t keep = __ballot_sync(0xFFFFFFFFu, c != ' ');
        if (c != ' ') dst[o + (uint32_t)__popc(keep & ((1u << lane) - 1u))] = (uint8_t)c;
        o += (uint32_t)__popc(keep);
    }

    // '_' + barcode1 + barcode2 + barcode3 + '_' + umi + '\n'
    const uint64_t an = args.ann[r];
    uint32_t cell = ann_cell(an);
    if (args.remap) cell = args.remap[cell];
    const uint32_t n = (uint32_t)cfg.n_wl;
    const uint32_t idx[3] = {cell / (n * n), (cell / n) % n, cell % n};
    if (lane == 0) dst[o] = '_';
    o += 1;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        uint32_t wl = cfg.wl_len[idx[k]];
        if ((uint32_t)lane < wl) dst[o + lane] = cfg.wl[idx[k] * 8 + lane];
        o += wl;
    }
    if (lane == 0) dst[o] = '_';
    o += 1;
    const uint32_t ulen = ann_umi_len(an, cfg.umi_len);
    if (ann_state(an) == 1) {
        if ((uint32_t)lane < ulen) dst[o + lane] = (uint8_t)base2_char((uint32_t)(ann_payload(an) >> (2 * lane)));
    } else {
        const size_t uoff = (size_t)(ann_payload(an) >> 5);
        if ((uint32_t)lane < ulen) dst[o + lane] = args.r2buf[uoff + lane];
    }
    o += ulen;
    if (lane == 0) dst[o] = '\n';
    o += 1;

    // read, "+", quality — each rstrip()ped (V2:246, 249)
    const size_t seq_e = rstrip_end(a, L.s[1], L.e[1]);
    const uint32_t seq_len = (uint32_t)(seq_e - L.s[1]);
    for (uint32_t i = lane; i < seq_len; i += 32) dst[o + i] = (uint8_t)a.byte(L.s[1] + i);
    o += seq_len;
    if (lane < 3) dst[o + lane] = (uint8_t)"\n+\n"[lane];
    o += 3;
    const size_t qual_e = rstrip_end(a, L.s[3], L.e[3]);
    const uint32_t qual_len = (uint32_t)(qual_e - L.s[3]);
    for (uint32_t i = lane; i < qual_len; i += 32) dst[o + i] = (uint8_t)a.byte(L.s[3] + i);
    o += qual_len;
    if (lane == 0) dst[o] = '\n';
    o += 1;
    return o;
}

template <typename OffT>
__global__ void __launch_bounds__(kThreads) k_emit(const __grid_constant__ EmitArgs<OffT> args, const __grid_constant__ DevCfg cfg)
{
    __shared__ __align__(128) uint8_t s_in[kEmitIn + 32];
    __shared__ __align__(128) uint8_t s_out[kEmitOut + 32];
    __shared__ __align__(16) uint32_t s_mask[kEmitInWords + 1];
    __shared__ __align__(8) uint64_t s_bar;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t r0 = (uint64_t)blockIdx.x * args.rpb;
    if (r0 >= args.n_rec) return;
    const uint64_t r1 = r0 + args.rpb < args.n_rec ? r0 + args.rpb : args.n_rec;
    const uint64_t out_lo = args.out_off[r0], out_hi = args.out_off[r1];
    if (out_hi == out_lo) return;  // nothing of this block survives the filter: no loads at all

    const size_t in_lo = (size_t)args.r1_start[r0];
    const size_t in_hi = r1 < args.n_rec_r1 ? (size_t)args.r1_start[r1] : args.r1_len;
    const size_t in_lo16 = in_lo & ~(size_t)15;
    const uint32_t out_skew = (uint32_t)(((uintptr_t)args.out + out_lo) & 15u);
    const bool staged = in_hi - in_lo16 <= (size_t)kEmitIn && out_hi - out_lo + out_skew <= (uint64_t)kEmitOut;

    if (!staged) {
        // records too long for the staging buffers: warps read global memory and write global memory directly
        GlobAcc ga{args.r1buf, args.r1_len};
        for (uint64_t r = r0 + warp; r < r1; r += kThreads / 32) {
            uint64_t o0 = args.out_off[r], o1 = args.out_off[r + 1];
            if (o1 == o0) continue;
            uint32_t wrote = emit_record(ga, args, cfg, r, args.out + o0);
            if (lane == 0 && wrote != (uint32_t)(o1 - o0)) atomicOr(&args.cnt->err_flags, kErrEmitLen);
        }
        return;
    }

    if (tid == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    const uint32_t in_bytes = (uint32_t)(in_hi - in_lo16);
    load_window(s_in, args.r1buf + in_lo16, in_bytes, &s_bar, 0);
    build_nl_mask(s_in, in_bytes, reinterpret_cast<uint16_t*>(s_mask), (in_bytes + 15u) / 16u + 2u);
    __syncthreads();

    // the window "ends the file" for line finding when it reaches the real end of read 1
    WinAcc wa{s_in, s_mask, in_lo16, in_bytes, args.r1_len};
    for (uint64_t r = r0 + warp; r < r1; r += kThreads / 32) {
        uint64_t o0 = args.out_off[r], o1 = args.out_off[r + 1];
        if (o1 == o0) continue;
        uint32_t wrote = emit_record(wa, args, cfg, r, s_out + out_skew + (uint32_t)(o0 - out_lo));
        if (lane == 0 && wrote != (uint32_t)(o1 - o0)) atomicOr(&args.cnt->err_flags, kErrEmitLen);
    }
    __syncthreads();

    // flush: s_out[out_skew + k] -> out[out_lo + k]; shared and global addresses agree modulo 16
    const uint32_t nout = (uint32_t)(out_hi - out_lo);
    uint8_t* gdst = args.out + out_lo;
    const uint32_t head = out_skew ? (16u - out_skew < nout ? 16u - out_skew : nout) : 0u;
    const uint32_t body = (nout - head) & ~15u;
    if ((uint32_t)tid < head) gdst[tid] = s_out[out_skew + tid];
    {
        const uint4* src = reinterpret_cast<const uint4*>(s_out + out_skew + head);
        uint4* dst = reinterpret_cast<uint4*>(gdst + head);
        for (uint32_t i = tid; i < body / 16u; i += kThreads) dst[i] = src[i];
    }
    for (uint32_t i = head + body + tid; i < nout; i += kThreads) gdst[i] = s_out[out_skew + i];
}

// ---------------------------------------------------------------------------------------------
// misc
// ---------------------------------------------------------------------------------------------
__global__ void k_record_cells(const uint64_t* ann, uint64_t n, int32_t* cells)
{
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) cells[r] = ann_state(ann[r]) ? (int32_t)ann_cell(ann[r]) : -1;
}

}  // namespace ssd
