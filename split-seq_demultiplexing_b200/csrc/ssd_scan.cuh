// ssd_scan.cuh — single-pass FASTQ record-boundary scan fused with the per-record work (sm_100a).
//
// One block owns a 32 KiB tile of the file.  The tile plus a 2 KiB halo is bulk-copied (TMA, one
// cp.async.bulk per block) into shared memory, newlines become a bitmap through 16-byte vector loads,
// the per-tile newline counts are chained with a decoupled look-back (warp-parallel) so the block learns
// the global line index of its first byte, and every record that starts in the tile is then processed
// from shared memory by one thread:
//   read 1: record offset + a 16-byte descriptor the writer needs (V2:134-141, 246, 249)
//   read 2: UMI / barcode extraction (V2:293-301), first-hit Hamming matching (V2:302-304, 327-329),
//           mate-id check (V2:350), per-cell histogram, (cell, UMI) keys
// The per-record code is written to stay convergent: line boundaries come from a per-tile list of line
// starts (no searching), headers and barcode windows are handled four bytes at a time with
// SIMD-in-register tests, loops have warp-uniform trip counts.  Whatever falls outside the fast path's
// preconditions (control characters in headers, truncated read 2, records longer than the window, more
// than two trailing blanks, ...) is routed to an exact generic path, first against the window and then
// against global memory.
#pragma once
#include "ssd_device.cuh"

namespace ssd {

constexpr int kTile = 32768;   // bytes of the file owned by one scan block
constexpr int kHalo = 2048;    // extra bytes loaded so records that start in the tile are whole
constexpr int kWin = kTile + kHalo;
constexpr int kWinWords = kWin / 32;
constexpr int kTileWords = kTile / 32;
constexpr int kHaloWords = kHalo / 32;
constexpr int kMaxRecTile = 1024;                 // records starting in one tile (>= 32 bytes each on average)
constexpr int kMaxLines = 4 * kMaxRecTile + 72;   // line-start list entries (tile + halo)

constexpr unsigned long long kFlagAgg = 1ull << 62;
constexpr unsigned long long kFlagPrefix = 2ull << 62;
constexpr unsigned long long kStatusValue = (1ull << 62) - 1;

// ---- read-1 record descriptor (16 bytes) ------------------------------------------------------------
//   x = cut_len | kept << 16      cut_len: header bytes before the first '/' (or the whole line),
//                                 kept:    how many of them survive .replace(" ", "").strip()
//   y = seq_rel | seq_len << 16   offsets are relative to the record start, lengths are rstrip()ped
//   z = qual_rel | qual_len << 16
//   w = flags; kMetaLong: the fields did not fit, the writer re-derives the record from its bytes
// next to it r1_base[r] = kept + seq_len + qual_len + 7: the bytes the record contributes to its output
// record besides barcodes and UMI (header' _ bc _ umi \n seq \n+\n qual \n has 7 fixed bytes)
constexpr uint32_t kMetaLong = 1u;


template <typename OffT>
struct ScanArgs {
    const uint8_t* buf;  // file being scanned
    size_t len;
    unsigned long long* status;  // one word per tile: flag << 62 | newline count (aggregate or inclusive prefix)
    uint32_t n_tiles;
    DevCounters* cnt;
    unsigned long long* err_off;  // [2] smallest offending byte offset per file
    uint64_t rec_cap;
    // read 1 outputs
    OffT* r1_start;
    uint4* r1_meta;
    uint32_t* r1_base;
    // read 2 inputs / outputs
    const uint8_t* r1buf;
    size_t r1_len;
    uint64_t* ann;
    uint32_t* hist;
    uint64_t* keys;
    IrrKey* irr;
    const uint8_t* lut;  // (e + 1) levels of 65536 bytes
};

__device__ __forceinline__ void raise(DevCounters* cnt, unsigned long long* err_off, int file, uint32_t flag, size_t off)
{
    atomicOr(&cnt->err_flags, flag);
    atomicMin(&err_off[file], (unsigned long long)off);
}

// ---- SIMD-in-register byte tests (flags land in bit 7 of each byte) ----------------------------------
__device__ __forceinline__ uint32_t zero_flags(uint32_t t) { return ~(((t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | t | 0x7F7F7F7Fu); }
// bytes whose low 7 bits are < n (n <= 0x80); bytes >= 0x80 may be flagged too (callers treat a flag as "take the exact path")
__device__ __forceinline__ uint32_t lt_flags(uint32_t x, uint32_t n4) { return ~((x | 0x80808080u) - n4) & 0x80808080u; }
// flags of the bytes [lo, hi) of a word
__device__ __forceinline__ uint32_t byte_range_flags(uint32_t lo, uint32_t hi) { return (0x80808080u << (8u * lo)) & (0x80808080u >> (8u * (4u - hi))); }

// four bases -> 8-bit 2-bit code (position j at bits 2j); `bad` collects x ^ (the bases those codes stand for)
__device__ __forceinline__ uint32_t pack4(uint32_t x, uint32_t& diff)
{
    uint32_t y = (x >> 1) & 0x03030303u;
    uint32_t sel = (y & 0x3u) | ((y >> 4) & 0x30u) | ((y >> 8) & 0x300u) | ((y >> 12) & 0x3000u);
    diff = x ^ __byte_perm(0x47544341u, 0u, sel);  // A C T G by code
    return (y * 0x01041040u) >> 24;
}

__device__ __forceinline__ void load8(const uint32_t* words, uint32_t p, uint32_t& a, uint32_t& b)
{
    uint32_t i = p >> 2, sh = (p & 3u) * 8u;
    uint32_t w0 = words[i], w1 = words[i + 1], w2 = words[i + 2];
    a = __funnelshift_r(w0, w1, sh);
    b = __funnelshift_r(w1, w2, sh);
}

// line.rstrip() for up to two trailing blanks; false when more would have to go (exact path)
__device__ __forceinline__ bool rstrip2(const uint8_t* win, uint32_t s, uint32_t& e)
{
    if (e > s && py_space(win[e - 1])) {
        e--;
        if (e > s && py_space(win[e - 1])) {
            e--;
            if (e > s && py_space(win[e - 1])) return false;
        }
    }
    return true;
}

// aligned 32-bit load of file bytes [off, off + 4) with zero fill past the end (off is a multiple of 4)
__device__ __forceinline__ uint32_t ldg_word_safe(const uint8_t* buf, size_t off, size_t len)
{
    if (off + 4 <= len) return __ldg(reinterpret_cast<const uint32_t*>(buf + off));
    uint32_t v = 0;
    for (int b = 0; b < 4; b++)
        if (off + b < len) v |= (uint32_t)buf[off + b] << (8 * b);
    return v;
}

// ======================================================================================================
// exact generic path (any accessor)
// ======================================================================================================
template <class Acc, typename OffT>
__device__ __noinline__ bool r1_record_generic(const Acc& a, const ScanArgs<OffT>& args, uint64_t r, size_t start)
{
    Lines L;
    if (!find_lines(a, start, args.len, true, L)) return false;
    if (!L.complete) return true;
    if (r >= args.rec_cap) {
        atomicOr(&args.cnt->err_flags, kErrOverflow);
        return true;
    }
    if (a.byte(L.s[0]) != '@') raise(args.cnt, args.err_off, 0, kErrNoAt, start);
    // name.split("/")[0].replace(" ", "").strip()  (V2:135-138)
    uint32_t kept = 0, kept_at_last_nonspace = 0;
    size_t end2 = L.s[0];
    for (size_t p = L.s[0]; p < L.e[0]; p++) {
        uint32_t c = a.byte(p);
        if (c == '/') break;
        if (c != ' ') kept++;
        if (!py_space(c)) {
            kept_at_last_nonspace = kept;
            end2 = p + 1;
        }
    }
    const size_t seq_e = rstrip_end(a, L.s[1], L.e[1]);
    const size_t qual_e = rstrip_end(a, L.s[3], L.e[3]);
    const size_t cut_len = end2 - L.s[0], seq_rel = L.s[1] - start, seq_len = seq_e - L.s[1], qual_rel = L.s[3] - start,
                 qual_len = qual_e - L.s[3];
    uint4 m;
    if (cut_len < 65536 && seq_rel < 65536 && seq_len < 65536 && qual_rel < 65536 && qual_len < 65536) {
        m.x = (uint32_t)cut_len | (kept_at_last_nonspace << 16);
        m.y = (uint32_t)seq_rel | ((uint32_t)seq_len << 16);
        m.z = (uint32_t)qual_rel | ((uint32_t)qual_len << 16);
        m.w = 0;
    } else {
        m.x = m.y = m.z = 0;
        m.w = kMetaLong;
    }
    args.r1_start[r] = (OffT)start;
    args.r1_meta[r] = m;
    args.r1_base[r] = kept_at_last_nonspace + (uint32_t)seq_len + (uint32_t)qual_len + 7u;
    return true;
}

struct R2Out {
    uint64_t ann;  // 0 = rejected
    uint64_t key;  // regular key
    IrrKey irr;
};

// Mate join by read id (V2:349-350), exact and bytewise: token [tok_s, tok_e) of read 2 against the first
// whitespace-delimited token of read-1 record r.
template <class Acc, typename OffT>
__device__ __forceinline__ bool mate_ok_generic(const Acc& a, const ScanArgs<OffT>& args, uint64_t r, size_t tok_s, size_t tok_e)
{
    if (r >= args.cnt->n_records[0]) return false;
    const size_t o1 = (size_t)args.r1_start[r];
    for (size_t j = 0;; j++) {
        bool e2 = tok_s + j >= tok_e;
        uint32_t c1 = o1 + j < args.r1_len ? args.r1buf[o1 + j] : (uint32_t)'\n';
        bool e1 = py_space(c1);
        if (e1 || e2) return e1 && e2;
        if (c1 != a.byte(tok_s + j)) return false;
    }
}

template <class Acc, typename OffT>
__device__ __noinline__ bool r2_record_generic(const Acc& a, const ScanArgs<OffT>& args, const DevCfg& cfg, uint64_t r, size_t start,
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
    const uint8_t* lut = args.lut + (size_t)cfg.e * 65536u;
    Packed p1 = pack_window(a, seq_s, seq_len, cfg.b1s, cfg.b1e, raw);
    int i1 = match_window(cfg, p1, raw, lut);
    if (i1 < 0) return true;  // V2:310-313
    Packed p2 = pack_window(a, seq_s, seq_len, cfg.b2s, cfg.b2e, raw);
    int i2 = match_window(cfg, p2, raw, lut);
    if (i2 < 0) return true;  // V2:314-317
    Packed p3 = pack_window(a, seq_s, seq_len, cfg.b3s, cfg.b3e, raw);
    int i3 = match_window(cfg, p3, raw, lut);
    if (i3 < 0) return true;  // V2:318-321
    const uint32_t cell = (uint32_t)((i1 * cfg.n_wl + i2) * cfg.n_wl + i3);

    size_t tok_e = L.s[0];
    while (tok_e < L.e[0] && !py_space(a.byte(tok_e))) tok_e++;
    if (!mate_ok_generic(a, args, r, L.s[0], tok_e)) {
        raise(args.cnt, args.err_off, 1, kErrKey, start);
        return true;
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

// ======================================================================================================
// fast paths (one thread per record, whole warps in lock step)
// ======================================================================================================

// First-hit index of an 8-base window given its 2-bit code and the XOR-against-expected words (zero = all ACGT).
__device__ __forceinline__ int match8(const DevCfg& cfg, const uint8_t* __restrict__ lut, uint32_t code, uint32_t d0, uint32_t d1)
{
    if ((d0 | d1) == 0u) {
        uint32_t v = __ldg(lut + (size_t)cfg.e * 65536u + code);
        return v == 0xFFu ? -1 : (int)v;
    }
    // some bytes are not A/C/G/T: they mismatch every whitelist base
    const uint32_t inv0 = ~zero_flags(d0) & 0x80808080u, inv1 = ~zero_flags(d1) & 0x80808080u;
    const int k = __popc(inv0) + __popc(inv1);
    if (k > cfg.e) return -1;
    if (k == 1) {
        // first i with dist over the 7 good positions <= e-1  ==  min over the 4 substitutions of LUT[e-1]
        const uint32_t bit = inv0 ? (uint32_t)__ffs((int)inv0) - 1u : 32u + (uint32_t)__ffs((int)inv1) - 1u;  // 7,15,...,63
        const uint32_t pos = bit >> 3;
        const uint8_t* l = lut + (size_t)(cfg.e - 1) * 65536u;
        const uint32_t base = code & ~(3u << (2u * pos));
        uint32_t best = 0xFFu;
#pragma unroll
        for (uint32_t b = 0; b < 4; b++) {
            uint32_t v = __ldg(l + (base | (b << (2u * pos))));
            best = v < best ? v : best;
        }
        return best == 0xFFu ? -1 : (int)best;
    }
    // several invalid positions: popcount loop over the whitelist
    uint32_t valid = 0;
#pragma unroll
    for (uint32_t j = 0; j < 4; j++) {
        if (!((inv0 >> (8u * j + 7u)) & 1u)) valid |= 1u << (2u * j);
        if (!((inv1 >> (8u * j + 7u)) & 1u)) valid |= 1u << (2u * (j + 4u));
    }
    for (int i = 0; i < cfg.n_wl; i++) {
        uint32_t x = code ^ cfg.wl_code[i];
        if (k + __popc((x | (x >> 1)) & valid) <= cfg.e) return i;
    }
    return -1;
}

constexpr int kScanR1 = 0;
constexpr int kScanR2 = 1;

template <int MODE, typename OffT>
__global__ void __launch_bounds__(kThreads) k_scan(const __grid_constant__ ScanArgs<OffT> args, const __grid_constant__ DevCfg cfg)
{
    __shared__ __align__(128) uint8_t s_win[kWin + 16];
    __shared__ __align__(16) uint32_t s_mask[kWinWords];
    __shared__ uint16_t s_lstart[kMaxLines];
    __shared__ uint32_t s_warp[kThreads / 32];
    __shared__ uint32_t s_halo[2];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ unsigned int s_tile;
    __shared__ unsigned long long s_line_base;
    __shared__ unsigned int s_nk, s_ni;
    __shared__ unsigned long long s_kbase, s_ibase;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        s_tile = atomicAdd(&args.cnt->ticket[MODE], 1u);  // tiles are claimed in order: look-back never waits on an unstarted tile
        mbar_init(&s_bar, 1);
        s_nk = 0;
        s_ni = 0;
    }
    __syncthreads();
    const uint32_t tile = s_tile;
    const size_t tile_lo = (size_t)tile * kTile;
    const uint32_t win_bytes = (uint32_t)(args.len - tile_lo < (size_t)kWin ? args.len - tile_lo : (size_t)kWin);

    load_window(s_win, args.buf + tile_lo, win_bytes, &s_bar, 0);
    build_nl_mask(s_win, win_bytes, reinterpret_cast<uint16_t*>(s_mask), kWinWords * 2);
    __syncthreads();

    // ---- newlines owned by this tile: mask words [0, kTileWords), 4 per thread; halo words: 1 per thread of warps 0-1
    uint32_t m[4], cnt = 0;
#pragma unroll
    for (int w = 0; w < 4; w++) {
        m[w] = s_mask[4 * tid + w];
        cnt += __popc(m[w]);
    }
    const uint32_t hm = tid < kHaloWords ? s_mask[kTileWords + tid] : 0u;
    uint32_t incl = cnt, hincl = __popc(hm);
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        uint32_t hv = __shfl_up_sync(0xFFFFFFFFu, hincl, d);
        if (lane >= d) {
            incl += v;
            hincl += hv;
        }
    }
    if (lane == 31) {
        s_warp[warp] = incl;
        if (warp < 2) s_halo[warp] = hincl;
    }
    __syncthreads();
    uint32_t warp_base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; w++) {
        uint32_t v = s_warp[w];
        if (w < warp) warp_base += v;
        total += v;
    }
    const uint32_t excl = warp_base + incl - cnt;
    const uint32_t halo_total = s_halo[0] + s_halo[1];
    const uint32_t hexcl = total + (warp == 1 ? s_halo[0] : 0u) + hincl - (uint32_t)__popc(hm);

    // ---- decoupled look-back over the tile newline counts, 32 predecessors per probe (warp 0) ----
    if (warp == 0) {
        volatile unsigned long long* st = args.status;
        unsigned long long excl_lines = 0;
        if (tile == 0) {
            if (lane == 0) st[0] = kFlagPrefix | total;
        } else {
            if (lane == 0) st[tile] = kFlagAgg | total;
            long long base = (long long)tile - 1;
            for (;;) {
                const long long idx = base - lane;
                unsigned long long v = 2ull << 62;  // before tile 0: an inclusive prefix of 0
                if (idx >= 0) v = st[idx];
                const uint32_t flag = (uint32_t)(v >> 62);
                const uint32_t ready = __ballot_sync(0xFFFFFFFFu, flag != 0u);
                const uint32_t pref = __ballot_sync(0xFFFFFFFFu, flag == 2u);
                const uint32_t first_pref = pref ? (uint32_t)__ffs((int)pref) - 1u : 31u;
                const uint32_t need = first_pref >= 31u ? 0xFFFFFFFFu : ((2u << first_pref) - 1u);  // lanes 0..first_pref
                if ((ready & need) != need) continue;  // a closer tile has not published yet: probe again
                unsigned long long contrib = ((need >> lane) & 1u) ? (v & kStatusValue) : 0ull;
#pragma unroll
                for (int d = 16; d; d >>= 1) contrib += __shfl_xor_sync(0xFFFFFFFFu, contrib, d);
                excl_lines += contrib;
                if (pref) break;
                base -= 32;
            }
            if (lane == 0) st[tile] = kFlagPrefix | (excl_lines + total);
        }
        if (lane == 0) {
            s_line_base = excl_lines;
            if (tile == args.n_tiles - 1) {
                unsigned long long n_lines = excl_lines + total + ((args.len && s_win[win_bytes - 1] != '\n') ? 1ull : 0ull);
                args.cnt->n_lines[MODE] = n_lines;
                args.cnt->n_records[MODE] = n_lines >> 2;
            }
        }
    }

    // ---- line-start list: s_lstart[1 + k] = byte after the k-th newline of the window (tile first, then halo) ----
    const uint32_t n_listed = total + halo_total;  // newlines listed when everything fits
    {
        uint32_t rank = excl;
#pragma unroll
        for (int w = 0; w < 4; w++) {
            uint32_t mm = m[w];
            while (mm) {
                uint32_t b = (uint32_t)__ffs((int)mm) - 1u;
                mm &= mm - 1u;
                if (rank + 1u < (uint32_t)kMaxLines) s_lstart[rank + 1u] = (uint16_t)((4 * tid + w) * 32 + b + 1);
                rank++;
            }
        }
        uint32_t hrank = hexcl, mm = hm;
        while (mm) {
            uint32_t b = (uint32_t)__ffs((int)mm) - 1u;
            mm &= mm - 1u;
            if (hrank + 1u < (uint32_t)kMaxLines) s_lstart[hrank + 1u] = (uint16_t)((kTileWords + tid) * 32 + b + 1);
            hrank++;
        }
        if (tid == 0) s_lstart[0] = 0;
    }
    __syncthreads();
    const unsigned long long line_base = s_line_base;

    // records that start in this tile: the line after a newline whose next-line index is 0 mod 4
    const unsigned long long r_first = tile == 0 ? 0ull : (line_base + 4) >> 2;
    const unsigned long long r_end = ((line_base + total) >> 2) + 1;  // exclusive
    uint32_t n_rec = r_end > r_first ? (uint32_t)(r_end - r_first) : 0u;
    if (tile == 0 && args.len == 0) n_rec = 0;
    if (n_rec > (uint32_t)kMaxRecTile || n_listed + 1u > (uint32_t)kMaxLines) {
        if (tid == 0) atomicOr(&args.cnt->err_flags, kErrDense);
        return;
    }
    const uint32_t j_first = (uint32_t)(4ull * r_first - line_base);  // list index of the first record's start

    const uint32_t* words = reinterpret_cast<const uint32_t*>(s_win);
    WinAcc wa{s_win, s_mask, tile_lo, win_bytes, args.len};
    GlobAcc ga{args.buf, args.len};

    for (uint32_t base = 0; base < n_rec; base += kThreads) {
        const uint32_t i = base + tid;
        const bool have = i < n_rec;
        const uint64_t r = r_first + i;
        const uint32_t j0 = j_first + 4u * i;
        uint32_t s0 = 0, s1 = 1, s2 = 1, s3 = 1, s4 = 1;
        bool fast = false;
        if (have) {
            s0 = s_lstart[j0];
            if (j0 + (MODE == kScanR1 ? 4u : 3u) <= n_listed) {
                s1 = s_lstart[j0 + 1];
                s2 = s_lstart[j0 + 2];
                s3 = s_lstart[j0 + 3];
                if (MODE == kScanR1) s4 = s_lstart[j0 + 4];
                fast = MODE == kScanR1 ? true : tile_lo + s3 < args.len;  // line 3 must exist
                if (r >= args.rec_cap) {
                    atomicOr(&args.cnt->err_flags, kErrOverflow);
                    fast = false;
                }
            }
        }
        // header geometry shared by both modes: bytes [s0, s0 + hdr_len), one trailing '\r' dropped
        uint32_t hdr_len = fast ? s1 - 1u - s0 : 0u;
        if (fast && hdr_len && s_win[s0 + hdr_len - 1u] == '\r') hdr_len--;
        const uint32_t a = s0 & 3u, nbytes = a + hdr_len, nwords = fast ? (nbytes + 3u) >> 2 : 0u;
        const uint32_t trip = __reduce_max_sync(0xFFFFFFFFu, nwords);

        if (MODE == kScanR1) {
            // ---- name.split("/")[0].replace(" ", "").strip() as counts (V2:135-138) ----
            uint32_t cut = hdr_len, spaces = 0, bad = 0;
            bool found = false;
            for (uint32_t w = 0; w < trip; w++) {
                if (w < nwords) {
                    const uint32_t x = words[(s0 >> 2) + w];
                    const uint32_t rem = nbytes - 4u * w;
                    const uint32_t V = byte_range_flags(w == 0 ? a : 0u, rem < 4u ? rem : 4u);
                    bad |= lt_flags(x, 0x20202020u) & V;
                    if (!found) {
                        const uint32_t zs = zero_flags(x ^ 0x2F2F2F2Fu) & V, zp = zero_flags(x ^ 0x20202020u) & V;
                        if (zs) {
                            const uint32_t bit = (uint32_t)__ffs((int)zs) - 1u;
                            cut = 4u * w + (bit >> 3) - a;
                            spaces += (uint32_t)__popc(zp & ((1u << bit) - 1u));
                            found = true;
                        } else {
                            spaces += (uint32_t)__popc(zp);
                        }
                    }
                }
            }
            uint32_t seq_e = s2 - 1u, qual_e = s4 - 1u;
            bool ok = fast && bad == 0u;
            if (ok) ok = rstrip2(s_win, s1, seq_e) && rstrip2(s_win, s3, qual_e);
            if (ok) {
                if (s_win[s0] != '@') raise(args.cnt, args.err_off, 0, kErrNoAt, tile_lo + s0);
                uint4 mt;
                mt.x = cut | ((cut - spaces) << 16);
                mt.y = (s1 - s0) | ((seq_e - s1) << 16);
                mt.z = (s3 - s0) | ((qual_e - s3) << 16);
                mt.w = 0;
                args.r1_start[r] = (OffT)(tile_lo + s0);
                args.r1_meta[r] = mt;
                args.r1_base[r] = (cut - spaces) + (seq_e - s1) + (qual_e - s3) + 7u;
            }
            __syncwarp();
            if (have && !ok) {
                const size_t start = tile_lo + s0;
                if (!r1_record_generic(wa, args, r, start)) r1_record_generic(ga, args, r, start);
            }
            __syncwarp();
        } else {
            // ---- read id = first whitespace-delimited token of the header (V2:281-283) ----
            uint32_t tok_len = hdr_len, ctl = 0;
            {
                bool found = false;
                for (uint32_t w = 0; w < trip; w++) {
                    if (w < nwords && !found) {
                        const uint32_t x = words[(s0 >> 2) + w];
                        const uint32_t rem = nbytes - 4u * w;
                        const uint32_t V = byte_range_flags(w == 0 ? a : 0u, rem < 4u ? rem : 4u);
                        const uint32_t le = lt_flags(x, 0x21212121u) & V;
                        if (le) {
                            const uint32_t bit = (uint32_t)__ffs((int)le) - 1u;
                            tok_len = 4u * w + (bit >> 3) - a;
                            ctl = lt_flags(x, 0x20202020u) & V & ((2u << bit) - 1u);  // the delimiter must be a plain blank
                            found = true;
                        }
                    }
                }
            }
            // ---- lineRead = line.rstrip(); the four windows (V2:293-301) ----
            uint32_t seq_e = s2 - 1u;
            bool ok = fast && ctl == 0u && cfg.fast_wl != 0 && cfg.fast_geom != 0;
            if (ok) ok = rstrip2(s_win, s1, seq_e);
            if (ok) ok = seq_e - s1 >= (uint32_t)cfg.need_len;  // truncated read 2 goes the exact way
            R2Out o;
            o.ann = 0;
            bool complete = false;
            uint32_t cell = 0, nw_tok = 0;
            bool matched = false;
            uint32_t u0 = 0, u1 = 0, u2 = 0;
            if (ok) {
                complete = true;
                if (s_win[s0] != '@') raise(args.cnt, args.err_off, 1, kErrNoAt, tile_lo + s0);
                uint32_t x0, x1, d0, d1, code;
                load8(words, s1 + (uint32_t)cfg.b1s, x0, x1);
                code = pack4(x0, d0) | (pack4(x1, d1) << 8);
                const int i1 = match8(cfg, args.lut, code, d0, d1);
                load8(words, s1 + (uint32_t)cfg.b2s, x0, x1);
                code = pack4(x0, d0) | (pack4(x1, d1) << 8);
                const int i2 = i1 < 0 ? -1 : match8(cfg, args.lut, code, d0, d1);
                load8(words, s1 + (uint32_t)cfg.b3s, x0, x1);
                code = pack4(x0, d0) | (pack4(x1, d1) << 8);
                const int i3 = i2 < 0 ? -1 : match8(cfg, args.lut, code, d0, d1);
                matched = i3 >= 0;  // V2:310-321: all three lists non-empty
                cell = matched ? (uint32_t)((i1 * cfg.n_wl + i2) * cfg.n_wl + i3) : 0u;
                nw_tok = matched ? (tok_len + 3u) >> 2 : 0u;
            }
            // ---- mate join by read id (V2:349-350): compare the token with read 1's, four bytes at a time ----
            const uint32_t trip2 = __reduce_max_sync(0xFFFFFFFFu, nw_tok);
            if (trip2) {
                bool same = matched && r < args.cnt->n_records[0];
                size_t o1 = 0, g = 0;
                uint32_t sh1 = 0, prev1 = 0, prev2 = 0;
                const uint32_t sh2 = a * 8u;
                if (same) {
                    o1 = (size_t)args.r1_start[r];
                    g = o1 & ~(size_t)3;
                    sh1 = (uint32_t)(o1 & 3) * 8u;
                    prev1 = ldg_word_safe(args.r1buf, g, args.r1_len);
                    prev2 = words[s0 >> 2];
                }
                for (uint32_t w = 0; w < trip2; w++) {
                    if (same && w < nw_tok) {
                        const uint32_t n1 = ldg_word_safe(args.r1buf, g + 4u * (w + 1u), args.r1_len);
                        const uint32_t n2 = words[(s0 >> 2) + w + 1u];
                        const uint32_t rem = tok_len - 4u * w;
                        const uint32_t mask = rem >= 4u ? 0xFFFFFFFFu : (1u << (8u * rem)) - 1u;
                        if ((__funnelshift_r(prev1, n1, sh1) ^ __funnelshift_r(prev2, n2, sh2)) & mask) same = false;
                        prev1 = n1;
                        prev2 = n2;
                    }
                }
                if (same) {  // read 1's token must end where read 2's does
                    const size_t p = o1 + tok_len;
                    same = p >= args.r1_len || py_space(args.r1buf[p]);
                }
                if (matched && !same) {
                    raise(args.cnt, args.err_off, 1, kErrKey, tile_lo + s0);
                    matched = false;
                }
            }
            // ---- UMI = lineRead[umi_start:umi_end] verbatim (V2:295) ----
            if (matched) {
                const uint32_t up = s1 + (uint32_t)cfg.us;
                const uint32_t i0 = up >> 2, sh = (up & 3u) * 8u;
                const uint32_t w0 = words[i0], w1 = words[i0 + 1], w2 = words[i0 + 2], w3 = words[i0 + 3];
                const uint32_t ul = (uint32_t)cfg.umi_len;
                const uint32_t m0 = ul >= 4u ? 0xFFFFFFFFu : (1u << (8u * ul)) - 1u;
                const uint32_t m1 = ul >= 8u ? 0xFFFFFFFFu : (ul > 4u ? (1u << (8u * (ul - 4u))) - 1u : 0u);
                const uint32_t m2 = ul > 8u ? (1u << (8u * (ul - 8u))) - 1u : 0u;
                u0 = __funnelshift_r(w0, w1, sh) & m0;
                u1 = __funnelshift_r(w1, w2, sh) & m1;
                u2 = __funnelshift_r(w2, w3, sh) & m2;
                uint32_t d0, d1, d2;
                const uint64_t code = ((uint64_t)pack4(u0, d0) | ((uint64_t)pack4(u1, d1) << 8) | ((uint64_t)pack4(u2, d2) << 16)) &
                                      ((1ull << cfg.umi_bits) - 1ull);
                if (((d0 & m0) | (d1 & m1) | (d2 & m2)) == 0u) {
                    o.ann = kAnnRegular | ((uint64_t)cell << kAnnCellShift) | code;
                    o.key = ((uint64_t)cell << cfg.umi_bits) | code;
                } else {
                    o.ann = kAnnIrregular | ((uint64_t)cell << kAnnCellShift) | ((uint64_t)(tile_lo + up) << 5) | (uint64_t)ul;
                    o.irr.hi = ((uint64_t)cell << kIrrCellShift) | ((uint64_t)ul << 32) | (uint64_t)__byte_perm(u0, 0u, 0x0123u);
                    o.irr.lo = ((uint64_t)__byte_perm(u1, 0u, 0x0123u) << 16) | (uint64_t)(((u2 & 0xFFu) << 8) | ((u2 >> 8) & 0xFFu));
                }
            }
            __syncwarp();
            if (have && !ok) {
                const size_t start = tile_lo + s0;
                if (!r2_record_generic(wa, args, cfg, r, start, o, complete)) r2_record_generic(ga, args, cfg, r, start, o, complete);
            }
            __syncwarp();

            // ---- results: annotation, histogram, block-aggregated key append ----
            uint32_t slot = 0;
            if (complete) args.ann[r] = o.ann;
            if (ann_state(o.ann) == 1) slot = atomicAdd(&s_nk, 1u);
            else if (ann_state(o.ann) == 2) slot = atomicAdd(&s_ni, 1u);
            if (o.ann) atomicAdd(&args.hist[ann_cell(o.ann)], 1u);
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

}  // namespace ssd
