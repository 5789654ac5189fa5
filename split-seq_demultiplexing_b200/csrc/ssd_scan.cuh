// ssd_scan.cuh — single-pass FASTQ record-boundary scan fused with the per-record work (sm_100a).
//
// Persistent, warp-specialised blocks (one wave, two per SM, 16 warps each) take 24 KiB tiles of the file by ticket.  A copy
// warp keeps the tiles (plus a 2 KiB halo) arriving in shared memory with bulk copies (TMA); stage A (8 warps) turns newlines
// into a bitmap through 16-byte vector loads, publishes the tile's newline count (a prefix warp chains the counts into global
// line indices: decoupled look-back) and lists the line starts; a guess warp guesses the record phase of the tile from the
// list; stage B (4 warps) processes every record that starts in the tile, one thread per record, from shared memory, on that
// guess, so that it never waits for the line index, and parks the results; a commit warp checks the guess and writes the parked
// results out under their record numbers once the index has arrived:
//   read 1: record offset + a 16-byte descriptor the writer needs (V2:134-141, 246, 249), read-id fingerprint
//   read 2: UMI / barcode extraction (V2:293-301), first-hit Hamming matching (V2:302-304, 327-329),
//           mate-id check (V2:350), per-cell histogram, (cell, UMI) keys
//   annotated input (MergedCells_1.fastq, V2:389-404): cell / UMI parsed back from the header
// The per-record code is written to stay convergent: line boundaries come from a per-tile list of line
// starts (no searching), headers and barcode windows are handled four bytes at a time with
// SIMD-in-register tests, loops have warp-uniform trip counts.  Whatever falls outside the fast path's
// preconditions (control characters in headers, records longer than the window, more than two trailing
// blanks, ... and, in the plain read-2 scan, a sequence that ends before the last slice does) is routed
// to an exact generic path against global memory, one thread per record; the SHORT variant of the read-2
// scan keeps short sequences on the fast path (the whole warp searches the whitelist for their windows).
#pragma once
#include "ssd_device.cuh"
#include "ssd_match.cuh"

namespace ssd {

#ifndef SSD_SCAN_TILE
#define SSD_SCAN_TILE 24576
#endif
#ifndef SSD_SCAN_MINBLOCKS
#define SSD_SCAN_MINBLOCKS 2
#endif
#ifndef SSD_SCAN_SLOTS
#define SSD_SCAN_SLOTS 3
#endif
constexpr int kTile = SSD_SCAN_TILE;  // bytes of the file owned by one tile
constexpr int kHalo = 2048;           // extra bytes loaded so records that start in the tile are whole
constexpr int kWinSlots = SSD_SCAN_SLOTS;  // window buffers per block: 1 being filled, kDepth waiting for their line index, 1 in use
constexpr int kDepth = kWinSlots - 2;      // how many tiles the bitmap/list stage runs ahead of the record stage
constexpr int kListSlots = kDepth + 1;
constexpr int kWin = kTile + kHalo;
constexpr int kWinWords = kWin / 32;
constexpr int kTileWords = kTile / 32;
constexpr int kHaloWords = kHalo / 32;
constexpr int kMaxRecTile = kTile / 32;            // records starting in one tile (>= 32 bytes each on average)
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
// read 1 only: the bytes from the first base to the end of the record are exactly read "\n+\n" quality "\n"
// (bare '+' line, nothing to strip), so the writer copies them in one piece
constexpr uint32_t kMetaTail = 2u;


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
    unsigned long long* r1_tok;  // read id fingerprint: 64-bit hash of the first header token, low 16 bits = its length
    // read 2 inputs / outputs
    const uint8_t* r1buf;
    size_t r1_len;
    uint64_t* ann;
    uint32_t* hist;
    IrrKey* irr;
    const uint8_t* lut;  // per distinct list: (e + 1) levels of 65536 bytes (lut_of)
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

// pack4 (four bases -> 2-bit codes + the bytes that are not bases): ssd_match.cuh

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

// ---- read-id fingerprint -------------------------------------------------------------------------------
// The mate join of the reference is by read id (V2:349-350).  Read 1's scan leaves, per record, a 64-bit
// fingerprint of the id (first whitespace-delimited header token): a hash over the token taken as little-endian
// 32-bit words from its first byte (last word zero-padded), with the token length in the low 16 bits.  Read 2's
// scan compares fingerprints instead of chasing read 1's bytes through global memory.
__device__ __forceinline__ unsigned long long tok_hash_step(unsigned long long h, uint32_t word)
{
    h = (h ^ (unsigned long long)word) * 0xFF51AFD7ED558CCDull;
    return h ^ (h >> 29);
}
__device__ __forceinline__ unsigned long long tok_hash_finish(unsigned long long h, uint32_t len)
{
    h = (h ^ (unsigned long long)len) * 0xC4CEB9FE1A85EC53ull;
    h ^= h >> 32;
    return (h & ~0xFFFFull) | (unsigned long long)(len & 0xFFFFu);
}
constexpr unsigned long long kTokSeed = 0x9E3779B97F4A7C15ull;

template <class Acc>
__device__ __forceinline__ unsigned long long tok_hash_generic(const Acc& a, size_t tok_s, size_t tok_e)
{
    unsigned long long h = kTokSeed;
    for (size_t p = tok_s; p < tok_e; p += 4) {
        uint32_t w = 0;
        for (int b = 0; b < 4; b++)
            if (p + b < tok_e) w |= a.byte(p + b) << (8 * b);
        h = tok_hash_step(h, w);
    }
    return tok_hash_finish(h, (uint32_t)(tok_e - tok_s));
}

// ======================================================================================================
// exact generic path (any accessor)
// ======================================================================================================
template <class Acc, typename OffT>
__device__ __noinline__ bool r1_record_generic(const Acc& a, const ScanArgs<OffT>& args, uint64_t r, size_t start, int header_style = 0)
{
    Lines L;
    if (!find_lines(a, start, args.len, true, L)) return false;
    if (!L.complete) return true;
    if (r >= args.rec_cap) {
        atomicOr(&args.cnt->err_flags, kErrOverflow);
        return true;
    }
    if (a.byte(L.s[0]) != '@') raise(args.cnt, args.err_off, 0, kErrNoAt, start);
    const size_t seq_e = rstrip_end(a, L.s[1], L.e[1]);
    const size_t qual_e = rstrip_end(a, L.s[3], L.e[3]);
    const size_t seq_rel = L.s[1] - start, seq_len = seq_e - L.s[1], qual_rel = L.s[3] - start, qual_len = qual_e - L.s[3];
    uint32_t kept_at_last_nonspace = 0;
    uint4 m;
    if (header_style == 1) {
        // Extract_BC_UMI.py:78-85: ReadName.split()[0] + "_" + cell + "_" + umi + " " + ReadName.split()[1]; the writer takes the
        // record from its bytes (kMetaLong), the scan only has to say how long the header gets
        size_t p = L.s[0];
        uint32_t tl[2] = {0, 0};
        int nt = 0;
        while (p < L.e[0] && nt < 2) {
            while (p < L.e[0] && py_space(a.byte(p))) p++;
            const size_t t0 = p;
            while (p < L.e[0] && !py_space(a.byte(p))) p++;
            if (p > t0) tl[nt++] = (uint32_t)(p - t0);
        }
        if (nt < 2) raise(args.cnt, args.err_off, 0, kErrIndex, start);  // ReadName_parts[1]: IndexError
        kept_at_last_nonspace = tl[0] + 1u + tl[1];
        m.x = m.y = m.z = 0;
        m.w = kMetaLong;
    } else {
        // name.split("/")[0].replace(" ", "").strip()  (V2:135-138)
        uint32_t kept = 0;
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
        const size_t cut_len = end2 - L.s[0];
        if (cut_len < 65536 && seq_rel < 65536 && seq_len < 65536 && qual_rel < 65536 && qual_len < 65536) {
            m.x = (uint32_t)cut_len | (kept_at_last_nonspace << 16);
            m.y = (uint32_t)seq_rel | ((uint32_t)seq_len << 16);
            m.z = (uint32_t)qual_rel | ((uint32_t)qual_len << 16);
            m.w = 0;
        } else {
            m.x = m.y = m.z = 0;
            m.w = kMetaLong;
        }
    }
    size_t tok_e = L.s[0];
    while (tok_e < L.e[0] && !py_space(a.byte(tok_e))) tok_e++;
    args.r1_start[r] = (OffT)start;
    args.r1_meta[r] = m;
    args.r1_base[r] = kept_at_last_nonspace + (uint32_t)seq_len + (uint32_t)qual_len + 7u;
    args.r1_tok[r] = tok_hash_generic(a, L.s[0], tok_e);
    return true;
}

struct R2Out {
    uint64_t ann;  // 0 = rejected
    IrrKey irr;
};

// Mate join by read id (V2:349-350): read-2 token [tok_s, tok_e) against the fingerprint read 1 left for record r.
template <class Acc, typename OffT>
__device__ __forceinline__ bool mate_ok_generic(const Acc& a, const ScanArgs<OffT>& args, uint64_t r, size_t tok_s, size_t tok_e, uint64_t n_rec_r1)
{
    if (r >= n_rec_r1) return false;
    return args.r1_tok[r] == tok_hash_generic(a, tok_s, tok_e);
}

template <class Acc, typename OffT>
__device__ __noinline__ bool r2_record_generic(const Acc& a, const ScanArgs<OffT>& args, const DevCfg& cfg, uint64_t r, size_t start,
                                               R2Out& out, bool& complete, uint64_t n_rec_r1)
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
    int i1 = match_window(cfg, 0, p1, raw, lut_of(cfg, args.lut, 0, cfg.e));
    if (i1 < 0) return true;  // V2:310-313
    Packed p2 = pack_window(a, seq_s, seq_len, cfg.b2s, cfg.b2e, raw);
    int i2 = match_window(cfg, 1, p2, raw, lut_of(cfg, args.lut, 1, cfg.e));
    if (i2 < 0) return true;  // V2:314-317
    Packed p3 = pack_window(a, seq_s, seq_len, cfg.b3s, cfg.b3e, raw);
    int i3 = match_window(cfg, 2, p3, raw, lut_of(cfg, args.lut, 2, cfg.e));
    if (i3 < 0) return true;  // V2:318-321
    const uint32_t cell = cell_of(cfg, i1, i2, i3);

    size_t tok_e = L.s[0];
    while (tok_e < L.e[0] && !py_space(a.byte(tok_e))) tok_e++;
    if (!mate_ok_generic(a, args, r, L.s[0], tok_e, n_rec_r1)) {
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
    } else {
        out.ann = kAnnIrregular | ((uint64_t)cell << kAnnCellShift) | ((uint64_t)(seq_s + ulo) << 5) | (uint64_t)ulen;
        out.irr.hi = ((uint64_t)cell << kIrrCellShift) | ((uint64_t)ulen << 32) | hi;
        out.irr.lo = lo;
    }
    return true;
}

// An annotated record (V2.calc_demux_results, V2:389-454), exact and bytewise.  Returns the annotation word.
template <class Acc, typename OffT>
__device__ __noinline__ uint64_t ann_record_generic(const Acc& a, const ScanArgs<OffT>& args, const DevCfg& cfg, uint64_t r, size_t start)
{
    Lines L;
    if (!find_lines(a, start, args.len, true, L) || !L.complete) return 0;
    if (r >= args.rec_cap) {
        atomicOr(&args.cnt->err_flags, kErrOverflow);
        return 0;
    }
    // every line is written strip()ped (V2:435-447); the fields come from the rstrip()ped header (V2:395-398)
    size_t ls[4], le[4];
    for (int k = 0; k < 4; k++) {
        le[k] = rstrip_end(a, L.s[k], L.e[k]);
        ls[k] = L.s[k];
        while (ls[k] < le[k] && py_space(a.byte(ls[k]))) ls[k]++;
    }
    size_t us[3] = {0, 0, 0};
    int nus = 0;
    for (size_t p = L.s[0]; p < le[0] && nus < 3; p++)
        if (a.byte(p) == '_') us[nus++] = p;
    uint64_t an = 0;
    if (nus < 2) {
        raise(args.cnt, args.err_off, 0, kErrIndex, start);
    } else {
        const size_t cell_s = us[0] + 1, cell_n = us[1] - us[0] - 1, umi_s = us[1] + 1, umi_e = nus == 3 ? us[2] : le[0];
        const size_t ulen = umi_e - umi_s;
        int idx[3] = {-1, -1, -1};
        // the cell must be three whitelist entries (of rounds 1, 2, 3) back to back (lengths as configured), matched exactly
        size_t p = cell_s;
        bool shape = true;
        for (int k = 0; k < 3 && shape; k++) {
            idx[k] = -1;
            for (int i = 0; i < cfg.n_wl[k] && idx[k] < 0; i++) {
                const size_t n = cfg.wl_len[k][i];
                if (p + n > cell_s + cell_n) continue;
                bool eq = true;
                for (size_t j = 0; j < n && eq; j++) eq = a.byte(p + j) == cfg.wl[k][i * 8 + j];
                if (eq && (k < 2 || p + n == cell_s + cell_n)) idx[k] = i;
            }
            if (idx[k] < 0) shape = false;
            else p += cfg.wl_len[k][idx[k]];
        }
        if (!shape || ulen > 10) {
            raise(args.cnt, args.err_off, 0, kErrForeign, start);
        } else {
            const uint32_t cell = cell_of(cfg, idx[0], idx[1], idx[2]);
            uint64_t code = 0;
            bool regular = (int)ulen == cfg.umi_len;
            for (size_t j = 0; j < ulen; j++) {
                const uint32_t c = a.byte(umi_s + j);
                regular = regular && is_acgt(c);
                code |= (uint64_t)base2(c) << (2 * j);
            }
            an = regular ? (kAnnRegular | ((uint64_t)cell << kAnnCellShift) | code)
                         : (kAnnIrregular | ((uint64_t)cell << kAnnCellShift) | ((uint64_t)umi_s << 5) | (uint64_t)ulen);
        }
    }
    const size_t hl = le[0] - ls[0], sl = le[1] - ls[1], pl = le[2] - ls[2], ql = le[3] - ls[3];
    uint4 m;
    if (ls[0] == start && hl < 65536 && ls[1] - start < 65536 && sl < 65536 && ls[2] - start < 65536 && pl < 32768 && ls[3] - start < 65536 &&
        ql < 65536) {
        m.x = (uint32_t)hl | ((uint32_t)hl << 16);
        m.y = (uint32_t)(ls[1] - start) | ((uint32_t)sl << 16);
        m.z = (uint32_t)(ls[3] - start) | ((uint32_t)ql << 16);
        m.w = ((uint32_t)pl << 1) | ((uint32_t)(ls[2] - start) << 16);
    } else {
        m.x = m.y = m.z = 0;
        m.w = kMetaLong;
    }
    args.r1_start[r] = (OffT)start;
    args.r1_meta[r] = m;
    args.r1_base[r] = (uint32_t)(hl + sl + pl + ql) + 4u;
    args.ann[r] = an;
    return an;
}

// ======================================================================================================
// fast paths (one thread per record, whole warps in lock step)
// ======================================================================================================

// First-hit index of an 8-base window given its 2-bit code and the XOR-against-expected words (zero = all ACGT).
__device__ __forceinline__ int match8(const DevCfg& cfg, int round, const uint8_t* __restrict__ lut, uint32_t code, uint32_t d0, uint32_t d1)
{
    if ((d0 | d1) == 0u) {
        uint32_t v = __ldg(lut_of(cfg, lut, round, cfg.e) + code);
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
        const uint8_t* l = lut_of(cfg, lut, round, cfg.e - 1);
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
    for (int i = 0; i < cfg.n_wl[round]; i++) {
        uint32_t x = code ^ cfg.wl_code[round][i];
        if (k + __popc((x | (x >> 1)) & valid) <= cfg.e) return i;
    }
    return -1;
}

// First-hit index for the windows the lookup tables do not answer (short read 2: zip() truncation, V2:72-73; several bytes
// outside ACGT), searched by the whole warp: the lanes that ask (`need`) take turns, 32 whitelist entries are tested at a time
// (codes = the round's 2-bit whitelist codes in shared memory), the lowest hit wins.  All 32 lanes call; -1 = no entry within e.
__device__ __forceinline__ int coop_first_hit(const uint16_t* codes, int n_wl, uint32_t e, bool need, uint32_t code, uint32_t valid, uint32_t k, int lane)
{
    int res = -1;
    uint32_t todo = __ballot_sync(0xFFFFFFFFu, need);
    while (todo) {
        const int src = __ffs((int)todo) - 1;
        todo &= todo - 1u;
        const uint32_t c = __shfl_sync(0xFFFFFFFFu, code, src), v = __shfl_sync(0xFFFFFFFFu, valid, src), kk = __shfl_sync(0xFFFFFFFFu, k, src);
        int hit = -1;
        if (kk <= e) {
            for (int base = 0; base < n_wl; base += 32) {
                const int i = base + lane;
                const bool h = i < n_wl && masked_hit(c, codes[i], v, kk, e);
                const uint32_t b = __ballot_sync(0xFFFFFFFFu, h);
                if (b) {
                    hit = base + __ffs((int)b) - 1;
                    break;
                }
            }
        }
        if (lane == src) res = hit;
    }
    return res;
}

constexpr int kScanR1 = 0;
constexpr int kScanR2 = 1;
constexpr int kScanAnn = 2;  // an already annotated FASTQ (MergedCells_1.fastq): V2.calc_demux_results


#ifndef SSD_SCAN_AWARPS
#define SSD_SCAN_AWARPS 8
#endif
#ifndef SSD_SCAN_R2_SPREAD
#define SSD_SCAN_R2_SPREAD 0
#endif
#ifndef SSD_SCAN_BWARPS
#define SSD_SCAN_BWARPS 4
#endif
#ifndef SSD_SCAN_BGROUPS
#define SSD_SCAN_BGROUPS 1
#endif
constexpr int kAWarps = SSD_SCAN_AWARPS, kBWarps = SSD_SCAN_BWARPS;  // first the stage A warps (bitmap, count, list), then the stage B warps (records)
// Stage B is a chain of dependent shared-memory reads and integer operations per record (one thread per record), so the time it
// takes per tile hardly depends on how many records the tile holds.  kBGroups groups of stage-B warps therefore take the tiles
// in turn (group g: the block's tiles g, g + kBGroups, ...): kBGroups tiles are in stage B at any time.
constexpr int kBGroups = SSD_SCAN_BGROUPS;
static_assert(kBWarps % kBGroups == 0, "stage-B warps per group");
constexpr int kBGWarps = kBWarps / kBGroups;
constexpr int kComputeWarps = kAWarps + kBWarps;
constexpr int kComputeThreads = kComputeWarps * 32;
constexpr int kAThreads = kAWarps * 32, kBThreads = kBGWarps * 32;  // kBThreads: threads of ONE stage-B group
constexpr uint32_t kNoTile = 0xFFFFFFFFu;
constexpr int kScanThreads = kComputeThreads + 128;  // the compute warps + the copy warp + the prefix warp (works in block 0) + the commit warp + the guess warp
#ifndef SSD_SCAN_STAGE_SLOTS
#define SSD_SCAN_STAGE_SLOTS (SSD_SCAN_BGROUPS > 1 ? SSD_SCAN_BGROUPS : 2)
#endif
constexpr int kStageSlots = SSD_SCAN_STAGE_SLOTS;  // tiles whose records wait (in shared memory) for their record numbers
static_assert(kStageSlots % kBGroups == 0, "a stage slot belongs to one stage-B group");

// Shared memory of one persistent scan block.
constexpr int kWinChunks = kWin / 16;                    // 16-byte chunks of a window

// What stage B found for the records of one tile, parked until the tile's first record number is known.
struct StageSlot {
    uint4 mt[kBThreads];                // read 1 / annotated: the record descriptor
    unsigned long long fp[kBThreads];   // read-id fingerprint
    unsigned long long ann[kBThreads];  // read 2 / annotated: the annotation word
    uint32_t s0f[kBThreads];            // offset of the record in the window | flags << 16
    uint32_t base[kBThreads];           // read 1 / annotated: output length contribution
};
constexpr uint32_t kStHave = 1u, kStOk = 2u, kStNoAt = 4u, kStMatched = 8u, kStComplete = 16u;  // flags; the error code of annotated input sits above

struct __align__(128) ScanSmem {
    uint8_t win[kWinSlots][kWin + 128];       // tile windows (bulk-copy destinations)
    StageSlot stage[kStageSlots];
    unsigned long long stage_full[kStageSlots], stage_empty[kStageSlots];  // mbarriers: stage B parked a tile / the commit warp wrote it out
    uint32_t stage_tile[kStageSlots], stage_total[kStageSlots], stage_guess[kStageSlots], stage_flags[kStageSlots];
    uint16_t lstart[kListSlots][kMaxLines];   // line-start lists of the tiles between stage A and stage B
    unsigned long long full[kWinSlots], empty[kWinSlots];  // mbarriers: bytes landed / window may be refilled
    unsigned long long list_full[kListSlots], list_empty[kListSlots];  // mbarriers: stage A done with a tile (and its phase guessed) / stage B done with it
    unsigned long long list_done[kListSlots];                          // mbarrier: stage A listed the tile's line starts (over to the guess warp)
    unsigned long long line_base[kBGroups];
    uint32_t tile[kWinSlots];
    uint32_t total[kListSlots], n_listed[kListSlots], list_tile[kListSlots];
    uint32_t guess_j[kListSlots], guess_n[kListSlots];  // the guess warp's guess of the record phase of the tile (stage B takes it from here)
    uint32_t warp_tile[kAWarps], halo_sum[2];
    alignas(16) uint32_t mask[kWinWords];     // newline bitmap of the tile in stage A (bit b of word w = byte 32 w + b)
    uint16_t wl_code[3][kMaxWl];              // read 2, SHORT variant: the 2-bit whitelist codes of the three rounds (coop_first_hit)
};
// SSD_SCAN_MINBLOCKS blocks per SM: 227 KB of shared memory per SM, 1 KB per block reserved by the system
static_assert(SSD_SCAN_MINBLOCKS * (sizeof(ScanSmem) + 1024) <= 227 * 1024, "the scan blocks of one SM do not fit its shared memory");

// which scans run the per-record work ahead of the line index (read 2 is issue-bound: nothing to hide there)
#ifndef SSD_SCAN_SPECULATE
#define SSD_SCAN_SPECULATE(mode) ((mode) != kScanR2)
#endif

// -DSSD_ASSERT: invariants of the warp-specialised scan checked on the device (a violation traps: every later call fails with a
// CUDA error).  compute-sanitizer is not available on the GPU pool, so the parity and fuzz suites are run once per round against
// a library built this way (profiles/run_assert.sh); off in the shipped build.
#ifdef SSD_ASSERT
#define SSD_CHECK(cond)                                                                                   \
    do {                                                                                                  \
        if (!(cond)) {                                                                                    \
            printf("SSD_ASSERT %s:%d block %d thread %d: %s\n", __FILE__, __LINE__, blockIdx.x, threadIdx.x, #cond); \
            __trap();                                                                                     \
        }                                                                                                 \
    } while (0)
#else
#define SSD_CHECK(cond)
#endif

#ifdef SSD_SCAN_PROFILE
__device__ unsigned long long g_scan_prof[2][16];
#define PROF_T(var) const long long var = clock64()
#define PROF_ADD(slot, a, b) if ((threadIdx.x & (kAThreads - 1)) == 0) prof[slot] += (unsigned long long)((b) - (a))
#else
#define PROF_T(var)
#define PROF_ADD(slot, a, b)
#endif

__device__ __forceinline__ void sync_a() { asm volatile("bar.sync 1, %0;" ::"n"(kAThreads) : "memory"); }
__device__ __forceinline__ void sync_b(int group) { asm volatile("bar.sync %0, %1;" ::"r"(2 + group), "n"(kBThreads) : "memory"); }

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Persistent, warp-specialised scan: one wave of blocks, tiles are claimed in order through a ticket.
//   copy warp     claims tiles and keeps kWinSlots bulk copies (TMA) in flight / in use
//   prefix warp   (block 0) walks the per-tile newline counts in tile order and turns each into the inclusive
//                 prefix, so a tile learns its global line index with one poll of its own entry
//   compute warps stage A(i + kDepth): newline bitmap (kept in registers), count (published), line-start list
//                 stage B(i)         : poll the line index, process the records that start in the tile
// SHORT (read 2 only): a read 2 whose sequence ends before the last slice does, and any window with several bytes outside ACGT,
// stays on the fast path -- its windows are matched by the whole warp against the whitelist (coop_first_hit) -- instead of the
// record going through the exact path, one thread against global memory.  The host picks the variant per context: it starts
// without (the common path is 5 % shorter) and switches when a batch sent more than 1 record in 1 000 down the exact path.
template <int MODE, typename OffT, bool DEFER, bool SHORT = false>
__global__ void __launch_bounds__(kScanThreads, SSD_SCAN_MINBLOCKS) k_scan(const __grid_constant__ ScanArgs<OffT> args, const __grid_constant__ DevCfg cfg)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    ScanSmem& S = *reinterpret_cast<ScanSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < kWinSlots; s++) {
            mbar_init(reinterpret_cast<uint64_t*>(&S.full[s]), 1);
            mbar_init(reinterpret_cast<uint64_t*>(&S.empty[s]), 1);
        }
        for (int s = 0; s < kListSlots; s++) {
            mbar_init(reinterpret_cast<uint64_t*>(&S.list_full[s]), 1);
            mbar_init(reinterpret_cast<uint64_t*>(&S.list_empty[s]), 1);
            mbar_init(reinterpret_cast<uint64_t*>(&S.list_done[s]), 1);
        }
        for (int s = 0; s < kStageSlots; s++) {
            mbar_init(reinterpret_cast<uint64_t*>(&S.stage_full[s]), 1);
            mbar_init(reinterpret_cast<uint64_t*>(&S.stage_empty[s]), 1);
        }
    }
    if constexpr (SHORT && MODE == kScanR2) {
        for (int i = tid; i < 3 * kMaxWl; i += kScanThreads) S.wl_code[i / kMaxWl][i % kMaxWl] = cfg.wl_code[i / kMaxWl][i % kMaxWl];
    }
    __syncthreads();
    volatile unsigned long long* st = args.status;

    if (warp == kComputeWarps + 3) {
        // ================================ guess warp ================================
        // Between stage A and stage B: guesses the record phase of every listed tile -- a record starts at the first line (of
        // the four candidates, one per lane) that begins with '@' and whose second next line begins with '+' -- so that stage B
        // can work on the records before the tile's line index is known.  A wrong guess costs a second pass, never a wrong
        // result.  A warp of its own: neither stage pays for the four dependent shared-memory reads.
        uint32_t ended = 0;
        for (uint32_t it = 0;; it++) {
            const uint32_t ws = it % kWinSlots, ls = it % kListSlots;
            mbar_wait(reinterpret_cast<uint64_t*>(&S.list_done[ls]), (it / kListSlots) & 1u);
            const uint32_t tile = S.list_tile[ls];
            if (tile == kNoTile) {  // the end: passed on to the stage-B group whose turn it is
                if (lane == 0) mbar_arrive(&S.list_full[ls]);
                if (++ended == (uint32_t)kBGroups) break;
                continue;
            }
            const uint8_t* s_win = S.win[ws];
            const uint16_t* s_lstart = S.lstart[ls];
            const uint32_t total = S.total[ls], n_listed = S.n_listed[ls];
            uint32_t j_guess = 0xFFFFFFFFu, n_guess = 0;
            const uint32_t j = (uint32_t)lane + 1u;
            bool cand = false;
            if (tile != 0 && n_listed + 1u <= (uint32_t)kMaxLines && j <= 4u && j <= total && j + 2u <= n_listed)
                cand = s_win[s_lstart[j]] == '@' && s_win[s_lstart[j + 2u]] == '+';
            const uint32_t cands = __ballot_sync(0xFFFFFFFFu, cand);
            if (lane == 0) {
                if (n_listed + 1u <= (uint32_t)kMaxLines) {
                    if (tile == 0) {
                        j_guess = 0;
                        n_guess = args.len ? (total >> 2) + 1u : 0u;
                    } else if (cands) {
                        j_guess = (uint32_t)__ffs((int)cands);  // lane + 1
                        n_guess = ((total - j_guess) >> 2) + 1u;
                    }
                }
                S.guess_j[ls] = j_guess;
                S.guess_n[ls] = n_guess;
                mbar_arrive(&S.list_full[ls]);  // hand the tile to stage B
            }
            __syncwarp();
        }
        return;
    }
    if (warp == kComputeWarps + 2) {
        // ================================ commit warp ================================
        // Stage B parks what it found for a tile's records; this warp waits for the tile's line index, checks that stage B
        // guessed the record phase right, and writes the records out under their numbers (coalesced).  Nothing else in the
        // block ever waits for the line index.
        if (!DEFER) return;
        GlobAcc ga{args.buf, args.len};
        const uint64_t n_rec_r1 = MODE == kScanR2 ? args.cnt->n_records[0] : 0;
        for (uint32_t it = 0;; it++) {
            const uint32_t ss = it % kStageSlots;
            mbar_wait(reinterpret_cast<uint64_t*>(&S.stage_full[ss]), (it / kStageSlots) & 1u);
            const uint32_t tile = S.stage_tile[ss];
            if (tile == kNoTile) break;
            const uint32_t total = S.stage_total[ss], guess = S.stage_guess[ss], sflags = S.stage_flags[ss];
            const size_t tile_lo = (size_t)tile * kTile;
            unsigned long long v = 0;
            if (lane == 0) {
                uint32_t ns = 32;
                while (((v = st[tile]) >> 62) != 2ull) {
                    __nanosleep(ns);
                    if (ns < 256u) ns += ns;
                }
            }
            v = __shfl_sync(0xFFFFFFFFu, v, 0);
            const unsigned long long line_base = (v & kStatusValue) - total;
            if (tile == args.n_tiles - 1 && lane == 0) {
                const unsigned long long n_lines = line_base + total + ((sflags & 2u) ? 1ull : 0ull);
                args.cnt->n_lines[MODE == kScanR2 ? 1 : 0] = n_lines;
                args.cnt->n_records[MODE == kScanR2 ? 1 : 0] = n_lines >> 2;
            }
            const unsigned long long r_first = tile == 0 ? 0ull : (line_base + 4) >> 2;
            const unsigned long long r_end = ((line_base + total) >> 2) + 1;  // exclusive
            uint32_t n_rec = r_end > r_first ? (uint32_t)(r_end - r_first) : 0u;
            if (tile == 0 && args.len == 0) n_rec = 0;
            if (n_rec > (uint32_t)kMaxRecTile || (sflags & 1u)) {
                if (lane == 0) atomicOr(&args.cnt->err_flags, kErrDense);
                n_rec = 0;
            }
            const uint32_t j_first = (uint32_t)(4ull * r_first - line_base);
            const bool good = guess != 0xFFFFFFFFu ? (n_rec == (guess >> 8) && (n_rec == 0u || j_first == (guess & 0xFFu))) : n_rec == 0u;
            if (!good) {
                if (lane == 0) atomicOr(&args.cnt->err_flags, kErrRespec << (MODE == kScanR2 ? 1 : 0));  // the host repeats this scan without guessing
                n_rec = 0;
            }
            const StageSlot& G = S.stage[ss];
            for (uint32_t i = (uint32_t)lane; i < n_rec; i += 32u) {
                const uint64_t r = r_first + i;
                SSD_CHECK(i < (uint32_t)kBThreads);
                const uint32_t sf = G.s0f[i], s0 = sf & 0xFFFFu, fl = sf >> 16;
                SSD_CHECK(!(fl & kStOk) || r >= args.rec_cap || tile_lo + s0 < args.len);
                const bool have = fl & kStHave, ok2 = (fl & kStOk) && r < args.rec_cap;  // beyond the capacity: the exact path raises the overflow flag
                if (MODE == kScanR1) {
                    if (ok2) {
                        if (fl & kStNoAt) raise(args.cnt, args.err_off, 0, kErrNoAt, tile_lo + s0);
                        args.r1_start[r] = (OffT)(tile_lo + s0);
                        args.r1_meta[r] = G.mt[i];
                        args.r1_base[r] = G.base[i];
                        args.r1_tok[r] = G.fp[i];
                    } else if (have) {
                        r1_record_generic(ga, args, r, tile_lo + s0, cfg.header_style);
                    }
                } else if (MODE == kScanAnn) {
                    if (ok2) {
                        if (fl >> 8) raise(args.cnt, args.err_off, 0, fl >> 8, tile_lo + s0);
                        args.r1_start[r] = (OffT)(tile_lo + s0);
                        args.r1_meta[r] = G.mt[i];
                        args.r1_base[r] = G.base[i];
                        args.ann[r] = G.ann[i];
                    } else if (have) {
                        const uint64_t an = ann_record_generic(ga, args, cfg, r, tile_lo + s0);
                        const uint32_t st8 = ann_state(an);
                        if (st8) atomicAdd(&args.hist[ann_cell(an)], 1u);
                        if (st8 == 1u) atomicAdd(&args.cnt->n_keys, 1ull);
                        if (st8 == 2u) atomicAdd(&args.cnt->n_irr, 1ull);
                    }
                } else {
                    if (ok2) {
                        if (fl & kStNoAt) raise(args.cnt, args.err_off, 1, kErrNoAt, tile_lo + s0);
                        // V2:349-350: the read-1 mate must carry the same id (fingerprint compare)
                        if ((fl & kStMatched) && !(r + 1u <= n_rec_r1 && args.r1_tok[r] == G.fp[i])) raise(args.cnt, args.err_off, 1, kErrKey, tile_lo + s0);
                        if (fl & kStComplete) args.ann[r] = G.ann[i];
                    } else if (have) {
                        R2Out o;
                        bool complete = false;
                        atomicAdd(&args.cnt->n_exact, 1u);
                        r2_record_generic(ga, args, cfg, r, tile_lo + s0, o, complete, n_rec_r1);
                        if (complete) args.ann[r] = o.ann;
                        const uint32_t st8 = ann_state(o.ann);
                        if (st8) atomicAdd(&args.hist[ann_cell(o.ann)], 1u);
                        if (st8 == 1u) atomicAdd(&args.cnt->n_keys, 1ull);
                        if (st8 == 2u) atomicAdd(&args.cnt->n_irr, 1ull);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.stage_empty[ss]);
        }
        return;
    }
    if (warp == kComputeWarps + 1) {
        // ================================ prefix warp (block 0 only) ================================
        if (blockIdx.x != 0) return;
        // every lane takes 8 consecutive entries (256 tiles per round trip to L2)
        unsigned long long running = 0;
        uint32_t next = 0, ns = 20;
        while (next < args.n_tiles) {
            unsigned long long v[8];
            const uint32_t first = next + 8u * (uint32_t)lane;
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = first + k < args.n_tiles ? st[first + k] : 0ull;
            // leading published entries of this lane, then of the whole warp
            uint32_t mine = 0;
#pragma unroll
            for (int k = 0; k < 8; k++)
                if (mine == (uint32_t)k && (v[k] >> 62) != 0ull) mine = k + 1u;
            const uint32_t full = __ballot_sync(0xFFFFFFFFu, mine == 8u);
            const uint32_t lead = full == 0xFFFFFFFFu ? 32u : (uint32_t)__ffs((int)~full) - 1u;  // lanes whose 8 entries are all published
            const uint32_t partial = __shfl_sync(0xFFFFFFFFu, mine, lead < 32u ? (int)lead : 31);
            const uint32_t take = (uint32_t)lane < lead ? 8u : ((uint32_t)lane == lead ? partial : 0u);  // entries this lane finalises
            unsigned long long sum = 0, pre[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if ((uint32_t)k < take) sum += v[k] & kStatusValue;
                pre[k] = sum;
            }
            unsigned long long incl = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                unsigned long long u = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= d) incl += u;
            }
            const unsigned long long lane_base = running + incl - sum;
#pragma unroll
            for (int k = 0; k < 8; k++)
                if ((uint32_t)k < take) st[first + k] = kFlagPrefix | (lane_base + pre[k]);
            running += __shfl_sync(0xFFFFFFFFu, incl, 31);
            const uint32_t advanced = 8u * lead + (lead < 32u ? partial : 0u);
            next += advanced;
            if (advanced == 0) {
                __nanosleep(ns);
                if (ns < 200u) ns += ns;
            } else {
                ns = 20;
            }
        }
        return;
    }
    if (warp == kComputeWarps) {
        // ================================ copy warp ================================
        for (uint32_t it = 0;; it++) {
            const uint32_t ws = it % kWinSlots;
            if (it >= (uint32_t)kWinSlots) mbar_wait(reinterpret_cast<uint64_t*>(&S.empty[ws]), ((it / kWinSlots) - 1u) & 1u);
#ifndef SSD_SCAN_ROUNDROBIN
            uint32_t t = 0;
            if (lane == 0) t = atomicAdd(&args.cnt->ticket[MODE == kScanR2 ? 1 : 0], 1u);  // tiles are claimed in order
            t = __shfl_sync(0xFFFFFFFFu, t, 0);
#else
            // round robin: block b owns tiles b, b + G, b + 2G, ...  All blocks of the (single-wave) grid are resident, and the
            // tiles whose counts a tile's line index depends on belong to the same or earlier rounds of the other blocks.
            const uint32_t t = it * gridDim.x + blockIdx.x;
#endif
            if (t >= args.n_tiles) {
                if (lane == 0) {
                    S.tile[ws] = kNoTile;
                    mbar_arrive(&S.full[ws]);
                }
                break;
            }
            const size_t tile_lo = (size_t)t * kTile;
            SSD_CHECK(tile_lo < args.len);
            const uint32_t wb = (uint32_t)(args.len - tile_lo < (size_t)kWin ? args.len - tile_lo : (size_t)kWin);
            const uint32_t bulk = wb & ~15u;
            for (uint32_t i = bulk + lane; i < wb; i += 32u) S.win[ws][i] = args.buf[tile_lo + i];
            __syncwarp();
            if (lane == 0) {
                S.tile[ws] = t;
                if (bulk) {
                    mbar_expect_tx(reinterpret_cast<uint64_t*>(&S.full[ws]), bulk);
                    bulk_g2s(S.win[ws], args.buf + tile_lo, bulk, reinterpret_cast<uint64_t*>(&S.full[ws]));
                } else {
                    mbar_arrive(&S.full[ws]);
                }
            }
            __syncwarp();
        }
        return;
    }

    // ================================ compute warps ================================
#ifdef SSD_SCAN_PROFILE
    unsigned long long prof[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#endif
    const uint64_t n_rec_r1 = MODE == kScanR2 ? args.cnt->n_records[0] : 0;  // read 1 was scanned by the previous launch

    // stage A (warps 0-3): newline bitmap, count, line-start list of the block's it-th tile; false when the tiles are exhausted
    auto stage_a = [&](uint32_t it) -> bool {
        const uint32_t ws = it % kWinSlots, ls = it % kListSlots;
        PROF_T(t9);
        if (it >= (uint32_t)kListSlots) mbar_wait(reinterpret_cast<uint64_t*>(&S.list_empty[ls]), ((it / kListSlots) - 1u) & 1u);
        PROF_T(t0);
        PROF_ADD(9, t9, t0);
        mbar_wait(reinterpret_cast<uint64_t*>(&S.full[ws]), (it / kWinSlots) & 1u);
        PROF_T(t1);
        PROF_ADD(0, t0, t1);
        const uint32_t tile = S.tile[ws];
        if (tile == kNoTile) {
            // the end, told to every stage-B group: the next kBGroups list slots in turn
            for (uint32_t k = 0; k < (uint32_t)kBGroups; k++) {
                const uint32_t itk = it + k, lk = itk % kListSlots;
                if (k && itk >= (uint32_t)kListSlots) mbar_wait(reinterpret_cast<uint64_t*>(&S.list_empty[lk]), ((itk / kListSlots) - 1u) & 1u);
                if (tid == 0) {
                    S.list_tile[lk] = kNoTile;
                    mbar_arrive(&S.list_done[lk]);  // the guess warp passes the end on to stage B
                }
            }
            return false;
        }
        const size_t tile_lo = (size_t)tile * kTile;
        const uint32_t win_bytes = (uint32_t)(args.len - tile_lo < (size_t)kWin ? args.len - tile_lo : (size_t)kWin);
        const uint8_t* s_win = S.win[ws];
        uint32_t* s_mask = S.mask;
        uint16_t* s_lstart = S.lstart[ls];
        // newlines owned by this tile: mask words [0, kTileWords), WPT consecutive ones per thread; halo words: 1 per thread of warps 0-1
        constexpr int WPT = kTileWords / kAThreads;  // mask words (32 bytes each) per thread
        static_assert(WPT >= 1 && WPT <= 8 && WPT * kAThreads == kTileWords, "tile size");
        static_assert(kHaloWords <= 64 && kAThreads >= 64, "halo size");
        uint32_t mw[WPT];
        uint32_t cnt = 0;
        build_nl_mask<kAThreads, (kWinWords * 2 + kAThreads - 1) / kAThreads>(s_win, win_bytes, reinterpret_cast<uint16_t*>(s_mask), kWinWords * 2);
        sync_a();
        PROF_T(t2);
        PROF_ADD(1, t1, t2);
#pragma unroll
        for (int w = 0; w < WPT; w++) {
            mw[w] = s_mask[WPT * tid + w];
            cnt += (uint32_t)__popc(mw[w]);
        }
        const uint32_t hm = tid < kHaloWords ? s_mask[kTileWords + tid] : 0u;
        // both prefix sums in one: a warp's tile count (<= 32 x 32 WPT bits) in the low half, its halo count in the high half
        static_assert(32 * 32 * WPT < 65536, "packed scan");
        uint32_t both = cnt | ((uint32_t)__popc(hm) << 16);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, both, d);
            if (lane >= d) both += v;
        }
        const uint32_t incl = both & 0xFFFFu, hincl = both >> 16;
        if (lane == 31) {
            S.warp_tile[warp] = incl;
            if (warp < 2) S.halo_sum[warp] = hincl;
        }
        PROF_T(t2a);
        sync_a();
        PROF_T(t2b);
        PROF_ADD(10, t2, t2a);
        PROF_ADD(11, t2a, t2b);
        uint32_t warp_base = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kAWarps; w++) {
            uint32_t v = S.warp_tile[w];
            if (w < warp) warp_base += v;
            total += v;
        }
        const uint32_t excl = warp_base + incl - cnt;
        const uint32_t halo0 = S.halo_sum[0], halo_total = halo0 + S.halo_sum[1];
        const uint32_t hexcl = total + (warp == 1 ? halo0 : 0u) + hincl - (uint32_t)__popc(hm);
        if (tid == 0) {
            S.total[ls] = total;
            S.n_listed[ls] = total + halo_total;
            S.list_tile[ls] = tile;
            st[tile] = kFlagAgg | total;  // published: the prefix warp takes it from here
            s_lstart[0] = 0;
        }
        // line-start list: s_lstart[1 + k] = byte after the k-th newline of the window (tile first, then halo);
        // one newline per iteration for every lane that still has one: warp-uniform trip count
        {
            // word by word (32 bytes each), without loops in the common case: a word holds at most two newlines unless lines
            // are shorter than 16 bytes ("\n+\n" gives the two); its first is the lowest set bit, its second the highest
            uint32_t idx = excl + 1u;
#pragma unroll
            for (int w = 0; w < WPT; w++) {
                uint32_t m = mw[w];
                const uint32_t p = (uint32_t)__popc(m);
                const uint32_t pos0 = (uint32_t)tid * (32 * WPT) + 32u * w + 1u;
                if (p > 2u) {  // rare: every newline of the word, one by one
                    while (m) {
                        const uint32_t b = (uint32_t)__ffs((int)m) - 1u;
                        m &= m - 1u;
                        if (idx < (uint32_t)kMaxLines) s_lstart[idx] = (uint16_t)(pos0 + b);
                        idx++;
                    }
                } else {
                    if (p >= 1u && idx < (uint32_t)kMaxLines) s_lstart[idx] = (uint16_t)(pos0 + (uint32_t)__ffs((int)m) - 1u);
                    if (p == 2u && idx + 1u < (uint32_t)kMaxLines) s_lstart[idx + 1u] = (uint16_t)(pos0 + 31u - (uint32_t)__clz((int)m));
                    idx += p;
                }
            }
            uint32_t hrank = hexcl, mm = hm;
            while (mm) {
                uint32_t b = (uint32_t)__ffs((int)mm) - 1u;
                mm &= mm - 1u;
                if (hrank + 1u < (uint32_t)kMaxLines) s_lstart[hrank + 1u] = (uint16_t)((kTileWords + tid) * 32 + b + 1);
                hrank++;
            }
        }
        PROF_T(t2c);
        PROF_ADD(12, t2b, t2c);
        sync_a();  // list complete; warp_tile / halo_sum / mask may be reused
        PROF_T(t2d);
        PROF_ADD(13, t2c, t2d);
        SSD_CHECK(total <= (uint32_t)kTile && halo_total <= (uint32_t)kHalo + 64u && S.list_tile[ls] == tile);
        if (tid == 0) mbar_arrive(&S.list_done[ls]);  // over to the guess warp, then to stage B
        PROF_T(t3);
        PROF_ADD(2, t2, t3);
        return true;
    };

    // stage B: the records that start in the block's it-th tile
    const int bgroup = warp >= kAWarps ? (warp - kAWarps) / kBGWarps : 0;
    const int tb = tid - kAThreads - bgroup * kBThreads, wb = warp - kAWarps - bgroup * kBGWarps;  // thread / warp index inside its stage-B group
    // read 2: the mate check still in flight (fingerprint of read 1 requested, not yet compared)
    bool pend = false;
    unsigned long long pend_fp1 = 0, pend_fp = 0;
    size_t pend_off = 0;
    auto settle = [&]() {
        if (pend && pend_fp1 != pend_fp) raise(args.cnt, args.err_off, 1, kErrKey, pend_off);
        pend = false;
    };
    auto stage_b = [&](uint32_t it) -> bool {
        const uint32_t ws = it % kWinSlots, ls = it % kListSlots;
        PROF_T(t7);
        mbar_wait(reinterpret_cast<uint64_t*>(&S.list_full[ls]), (it / kListSlots) & 1u);
        PROF_T(t8);
        PROF_ADD(5, t7, t8);
        const uint32_t tile = S.list_tile[ls];
        const uint32_t ss = it % kStageSlots;
        if (tile == kNoTile) {
            if (DEFER && it >= (uint32_t)kStageSlots) mbar_wait(reinterpret_cast<uint64_t*>(&S.stage_empty[ss]), ((it / kStageSlots) - 1u) & 1u);
            if (DEFER && tb == 0) {
                S.stage_tile[ss] = kNoTile;  // the commit warp stops at the first of these
                mbar_arrive(&S.stage_full[ss]);
            }
            if (tb == 0) mbar_arrive(&S.list_empty[ls]);  // stage A may be waiting for this slot to tell the next group
            return false;
        }
        const size_t tile_lo = (size_t)tile * kTile;
        const uint32_t win_bytes = (uint32_t)(args.len - tile_lo < (size_t)kWin ? args.len - tile_lo : (size_t)kWin);
        const uint8_t* s_win = S.win[ws];
        const uint16_t* s_lstart = S.lstart[ls];
        const uint32_t total = S.total[ls];
        const uint32_t n_listed = S.n_listed[ls];
        const uint32_t* words = reinterpret_cast<const uint32_t*>(s_win);
        GlobAcc ga{args.buf, args.len};

        // ---- per-record work, split in two: compute() needs only the tile (which list entries start records is all it takes
        //      from the global line index), commit() needs the record number.  When the phase can be guessed from the
        //      tile, compute() runs before the line index has arrived and hides the wait for it. ----
        bool have = false, ok = false, complete = false, noat = false, matched = false;
        uint32_t s0 = 0, base_len = 0, err_code = 0;
        uint4 mt = make_uint4(0, 0, 0, 0);
        unsigned long long fp = 0;
        uint64_t an = 0;
        R2Out o;
        o.ann = 0;

        auto compute = [&](uint32_t i, uint32_t j_first, uint32_t n_rec) {
            have = i < n_rec;
            ok = complete = noat = matched = false;
            err_code = 0;
            an = 0;
            o.ann = 0;
            const uint32_t j0 = j_first + 4u * i;
            uint32_t s1 = 1, s2 = 1, s3 = 1, s4 = 1;
            s0 = 0;
            bool fast = false;
            if (have) {
                SSD_CHECK(j0 <= n_listed && j0 < (uint32_t)kMaxLines);
                s0 = s_lstart[j0];
                SSD_CHECK(s0 <= (uint32_t)kTile && s0 <= win_bytes);  // a record of the tile starts inside the tile
                if (j0 + (MODE == kScanR2 ? 3u : 4u) <= n_listed) {
                    s1 = s_lstart[j0 + 1];
                    s2 = s_lstart[j0 + 2];
                    s3 = s_lstart[j0 + 3];
                    if (MODE != kScanR2) s4 = s_lstart[j0 + 4];
                    SSD_CHECK(s0 < s1 && s1 < s2 && s2 < s3 && s3 <= win_bytes && (MODE == kScanR2 || (s3 < s4 && s4 <= win_bytes)));
                    // read 2 only needs line 3 to exist; the generic path raises the overflow flag
                    fast = MODE != kScanR2 || tile_lo + s3 < args.len;
                }
            }
            // ---- read 2: the three windows are packed and ALL their table lookups requested before the header walk, which hides
            //      the trip to the L2 (the tables do not fit what the shared-memory carve-out leaves of the L1).  A window with one
            //      byte outside ACGT needs four lookups in the level e - 1 table (match8): they go out in the same batch, for one
            //      such window per record; records with more take the lookups of match8 afterwards (rare). ----
            uint32_t code1 = 0, code2 = 0, code3 = 0, dirty = 0, w_seq_e = s2 - 1u;
            uint32_t v1 = 0xFFu, v2 = 0xFFu, v3 = 0xFFu, x0 = 0xFFu, x1 = 0xFFu, x2 = 0xFFu, x3 = 0xFFu;  // x*: the four lookups of the one dirty window
            uint32_t da0 = 0, da1 = 0, db0 = 0, db1 = 0, dc0 = 0, dc1 = 0;
            bool win_ok = false, shortseq = false;
            if (MODE == kScanR2) {
                win_ok = fast && cfg.fast_wl != 0 && cfg.fast_geom != 0;
                if (win_ok) win_ok = rstrip2(s_win, s1, w_seq_e);
                if constexpr (SHORT) {
                    shortseq = win_ok && w_seq_e - s1 < (uint32_t)cfg.need_len;  // truncated read 2: its windows are matched by the warp below
                } else {
                    if (win_ok) win_ok = w_seq_e - s1 >= (uint32_t)cfg.need_len;  // truncated read 2 goes the exact way
                }
                if (win_ok && !shortseq) {
                    uint32_t y0, y1;
                    load8(words, s1 + (uint32_t)cfg.b1s, y0, y1);
                    code1 = pack4(y0, da0) | (pack4(y1, da1) << 8);
                    load8(words, s1 + (uint32_t)cfg.b2s, y0, y1);
                    code2 = pack4(y0, db0) | (pack4(y1, db1) << 8);
                    load8(words, s1 + (uint32_t)cfg.b3s, y0, y1);
                    code3 = pack4(y0, dc0) | (pack4(y1, dc1) << 8);
                    dirty = ((da0 | da1) ? 1u : 0u) | ((db0 | db1) ? 2u : 0u) | ((dc0 | dc1) ? 4u : 0u);
                    if (!(dirty & 1u)) v1 = __ldg(lut_of(cfg, args.lut, 0, cfg.e) + code1);
                    if (!(dirty & 2u)) v2 = __ldg(lut_of(cfg, args.lut, 1, cfg.e) + code2);
                    if (!(dirty & 4u)) v3 = __ldg(lut_of(cfg, args.lut, 2, cfg.e) + code3);
                    if (dirty && (dirty & (dirty - 1u)) == 0u && cfg.e >= 1) {  // exactly one dirty window
                        const int rd = dirty == 1u ? 0 : (dirty == 2u ? 1 : 2);
                        const uint32_t dc = rd == 0 ? code1 : (rd == 1 ? code2 : code3);
                        const uint32_t q0 = rd == 0 ? da0 : (rd == 1 ? db0 : dc0), q1 = rd == 0 ? da1 : (rd == 1 ? db1 : dc1);
                        const uint32_t inv0 = ~zero_flags(q0) & 0x80808080u, inv1 = ~zero_flags(q1) & 0x80808080u;
                        if (__popc(inv0) + __popc(inv1) == 1) {
                            const uint32_t bit = inv0 ? (uint32_t)__ffs((int)inv0) - 1u : 32u + (uint32_t)__ffs((int)inv1) - 1u;
                            const uint32_t sh = 2u * (bit >> 3), base = dc & ~(3u << sh);
                            const uint8_t* l = lut_of(cfg, args.lut, rd, cfg.e - 1);
                            x0 = __ldg(l + base);
                            x1 = __ldg(l + (base | (1u << sh)));
                            x2 = __ldg(l + (base | (2u << sh)));
                            x3 = __ldg(l + (base | (3u << sh)));
                            dirty |= 8u;  // answered by x0..x3
                        }
                    }
                }
            }
            // header bytes [s0, s0 + hdr_len), one trailing '\r' dropped; token = bytes before the first byte <= 0x20
            uint32_t hdr_len = fast ? s1 - 1u - s0 : 0u;
            if (fast && hdr_len && s_win[s0 + hdr_len - 1u] == '\r') hdr_len--;
            const uint32_t a = s0 & 3u, nwords = fast ? (hdr_len + 3u) >> 2 : 0u;
            const uint32_t trip = __reduce_max_sync(0xFFFFFFFFu, nwords);
            // one pass over the header, four bytes at a time from its first byte (word w = bytes 4w .. 4w+3): id fingerprint
            // (both mates), and for read 1 the pieces of name.split("/")[0].replace(" ", "").strip() (V2:135-138)
            uint32_t cut = hdr_len, spaces = 0, bad = 0, tok_len = hdr_len;
            uint32_t us1 = 0xFFFFu, us2 = 0xFFFFu, us3 = 0xFFFFu;  // annotated mode: the first three '_' of the header
            bool found_slash = false, found_tok = false;
            unsigned long long h = kTokSeed;
            {
                uint32_t prev = nwords ? words[s0 >> 2] : 0u;
                for (uint32_t w = 0; w < trip; w++) {
                    if (w < nwords) {
                        const uint32_t nxt = words[(s0 >> 2) + w + 1u];
                        const uint32_t x = __funnelshift_r(prev, nxt, a * 8u);
                        prev = nxt;
                        const uint32_t rem = hdr_len - 4u * w;
                        const uint32_t V = byte_range_flags(0u, rem < 4u ? rem : 4u);
                        bad |= lt_flags(x, 0x20202020u) & V;
                        if (!found_tok) {
                            const uint32_t le = lt_flags(x, 0x21212121u) & V;
                            uint32_t tw = x;
                            if (le) {
                                const uint32_t bit = (uint32_t)__ffs((int)le) - 1u;  // 7, 15, 23 or 31
                                tok_len = 4u * w + (bit >> 3);
                                found_tok = true;
                                tw &= (1u << (bit - 7u)) - 1u;  // bytes before the delimiter
                            } else if (rem < 4u) {
                                tw &= (1u << (8u * rem)) - 1u;  // the line ends inside this word
                            }
                            if (4u * w < tok_len) h = tok_hash_step(h, tw);
                        }
                        if (MODE == kScanAnn && us3 == 0xFFFFu) {
                            uint32_t zu = zero_flags(x ^ 0x5F5F5F5Fu) & V;
                            while (zu && us3 == 0xFFFFu) {
                                const uint32_t p = 4u * w + (((uint32_t)__ffs((int)zu) - 1u) >> 3);
                                zu &= zu - 1u;
                                if (us1 == 0xFFFFu) us1 = p;
                                else if (us2 == 0xFFFFu) us2 = p;
                                else us3 = p;
                            }
                        }
                        if (MODE == kScanR1 && !found_slash) {
                            const uint32_t zs = zero_flags(x ^ 0x2F2F2F2Fu) & V, zp = zero_flags(x ^ 0x20202020u) & V;
                            if (zs) {
                                const uint32_t bit = (uint32_t)__ffs((int)zs) - 1u;
                                cut = 4u * w + (bit >> 3);
                                spaces += (uint32_t)__popc(zp & ((1u << bit) - 1u));
                                found_slash = true;
                            } else {
                                spaces += (uint32_t)__popc(zp);
                            }
                        }
                    }
                    // read 2 only needs the id token: stop as soon as every lane has seen its end (control bytes behind it do
                    // not matter, a control byte before or at it has been flagged above)
                    if (MODE == kScanR2 && __all_sync(0xFFFFFFFFu, found_tok || w + 1u >= nwords)) break;
                }
            }
            fp = tok_hash_finish(h, tok_len);

            if (MODE == kScanR1) {
                uint32_t seq_e = s2 - 1u, qual_e = s4 - 1u;
                ok = fast && bad == 0u && cfg.header_style == 0;  // the other header style goes the exact way, record by record
                if (ok) ok = rstrip2(s_win, s1, seq_e) && rstrip2(s_win, s3, qual_e);
                if (ok) {
                    noat = s_win[s0] != '@';
                    mt.x = cut | ((cut - spaces) << 16);
                    mt.y = (s1 - s0) | ((seq_e - s1) << 16);
                    mt.z = (s3 - s0) | ((qual_e - s3) << 16);
                    mt.w = (seq_e == s2 - 1u && s3 - s2 == 2u && s_win[s2] == '+' && qual_e == s4 - 1u) ? kMetaTail : 0u;
                    base_len = (cut - spaces) + (seq_e - s1) + (qual_e - s3) + 7u;
                }
            } else if (MODE == kScanAnn) {
                // ---- an annotated record: header = id _ cell _ umi [_ ...] (V2:395-398); every line is copied strip()ped ----
                uint32_t seq_e = s2 - 1u, plus_e = s3 - 1u, qual_e = s4 - 1u;
                ok = fast && bad == 0u && cfg.fast_wl != 0 && cfg.bc8 != 0 && hdr_len != 0u && !py_space(s_win[s0]);
                if (ok) ok = rstrip2(s_win, s1, seq_e) && rstrip2(s_win, s2, plus_e) && rstrip2(s_win, s3, qual_e);
                // leading blanks of the other lines would have to be stripped too: exact path
                if (ok) ok = !(seq_e > s1 && py_space(s_win[s1])) && !(plus_e > s2 && py_space(s_win[s2])) && !(qual_e > s3 && py_space(s_win[s3]));
                if (ok) {
                    if (us2 == 0xFFFFu) {
                        err_code = kErrIndex;
                    } else {
                        const uint32_t cell_s = us1 + 1u, cell_n = us2 - us1 - 1u;
                        const uint32_t umi_s = us2 + 1u, ulen = (us3 == 0xFFFFu ? hdr_len : us3) - umi_s;
                        int idx[3] = {-1, -1, -1};
                        if (cell_n == 24u) {
#pragma unroll
                            for (int k = 0; k < 3; k++) {
                                uint32_t x0, x1, d0, d1;
                                load8(words, s0 + cell_s + 8u * k, x0, x1);
                                const uint32_t code = pack4(x0, d0) | (pack4(x1, d1) << 8);
                                if ((d0 | d1) == 0u) {
                                    const uint32_t v = __ldg(lut_of(cfg, args.lut, k, 0) + code);  // level 0: exact match
                                    idx[k] = v == 0xFFu ? -1 : (int)v;
                                }
                            }
                        }
                        if (idx[0] < 0 || idx[1] < 0 || idx[2] < 0 || ulen > 10u) {
                            err_code = kErrForeign;
                        } else {
                            const uint32_t cell = cell_of(cfg, idx[0], idx[1], idx[2]);
                            const uint32_t up = s0 + umi_s;
                            const uint32_t i0 = up >> 2, sh = (up & 3u) * 8u;
                            const uint32_t w0 = words[i0], w1 = words[i0 + 1], w2 = words[i0 + 2], w3 = words[i0 + 3];
                            const uint32_t m0 = ulen >= 4u ? 0xFFFFFFFFu : (1u << (8u * ulen)) - 1u;
                            const uint32_t m1 = ulen >= 8u ? 0xFFFFFFFFu : (ulen > 4u ? (1u << (8u * (ulen - 4u))) - 1u : 0u);
                            const uint32_t m2 = ulen > 8u ? (1u << (8u * (ulen - 8u))) - 1u : 0u;
                            const uint32_t u0 = __funnelshift_r(w0, w1, sh) & m0, u1 = __funnelshift_r(w1, w2, sh) & m1,
                                           u2 = __funnelshift_r(w2, w3, sh) & m2;
                            uint32_t e0, e1, e2;
                            const uint64_t ucode = ((uint64_t)pack4(u0, e0) | ((uint64_t)pack4(u1, e1) << 8) | ((uint64_t)pack4(u2, e2) << 16)) &
                                                   ((1ull << cfg.umi_bits) - 1ull);
                            if (ulen == (uint32_t)cfg.umi_len && ((e0 & m0) | (e1 & m1) | (e2 & m2)) == 0u)
                                an = kAnnRegular | ((uint64_t)cell << kAnnCellShift) | ucode;
                            else
                                an = kAnnIrregular | ((uint64_t)cell << kAnnCellShift) | ((uint64_t)(tile_lo + up) << 5) | (uint64_t)ulen;
                        }
                    }
                    mt.x = hdr_len | (hdr_len << 16);
                    mt.y = (s1 - s0) | ((seq_e - s1) << 16);
                    mt.z = (s3 - s0) | ((qual_e - s3) << 16);
                    mt.w = ((plus_e - s2) << 1) | ((s2 - s0) << 16);
                    base_len = hdr_len + (seq_e - s1) + (plus_e - s2) + (qual_e - s3) + 4u;
                }
            } else {
                // ---- lineRead = line.rstrip(); the four windows (V2:293-301): requested above, answered by now ----
                ok = win_ok && bad == 0u;
                // the windows the tables did not answer -- all three of a short sequence (lineRead[s:e] gives what is there, hamming()
                // zips to the shorter operand: V2:72-73, 296-301), and dirty windows beyond the one-N case -- go to the whole warp
                const uint32_t seq_len = w_seq_e - s1;
                int c1 = -1, c2 = -1, c3 = -1;
                if constexpr (SHORT) {
                    const uint32_t ask = !ok ? 0u : (shortseq ? 7u : ((dirty & 8u) ? 0u : (dirty & 7u)));
                    if (__any_sync(0xFFFFFFFFu, ask != 0u)) {
#pragma unroll
                        for (int rd = 0; rd < 3; rd++) {
                            const bool need = ((ask >> rd) & 1u) != 0u;
                            uint32_t c = 0, v = 0, kk = 0;
                            if (need) {
                                if (shortseq) {
                                    const uint32_t bs = (uint32_t)(rd == 0 ? cfg.b1s : (rd == 1 ? cfg.b2s : cfg.b3s));
                                    const uint32_t n = seq_len > bs ? min(seq_len - bs, 8u) : 0u;  // bytes of the window that exist
                                    if (n) {
                                        uint32_t y0, y1, q0, q1;
                                        load8(words, s1 + bs, y0, y1);
                                        c = pack4(y0, q0) | (pack4(y1, q1) << 8);
                                        window_masks(q0, q1, n, v, kk);
                                    }
                                } else {
                                    c = rd == 0 ? code1 : (rd == 1 ? code2 : code3);
                                    window_masks(rd == 0 ? da0 : (rd == 1 ? db0 : dc0), rd == 0 ? da1 : (rd == 1 ? db1 : dc1), 8u, v, kk);
                                }
                            }
                            const int hit = coop_first_hit(S.wl_code[rd], cfg.n_wl[rd], (uint32_t)cfg.e, need, c, v, kk, lane);
                            if (rd == 0) c1 = hit;
                            else if (rd == 1) c2 = hit;
                            else c3 = hit;
                        }
                    }
                }
                if (ok) {
                    complete = true;
                    noat = s_win[s0] != '@';
                    int i1 = v1 == 0xFFu ? -1 : (int)v1, i2 = v2 == 0xFFu ? -1 : (int)v2, i3 = v3 == 0xFFu ? -1 : (int)v3;
                    if (SHORT && shortseq) {
                        i1 = c1;
                        i2 = c2;
                        i3 = c3;
                    } else if (dirty) {  // some window holds a byte that is not A/C/G/T
                        if (dirty & 8u) {
                            const uint32_t best = min(min(x0, x1), min(x2, x3));
                            const int ix = best == 0xFFu ? -1 : (int)best;
                            if (dirty & 1u) i1 = ix;
                            if (dirty & 2u) i2 = ix;
                            if (dirty & 4u) i3 = ix;
                        } else {
                            if constexpr (SHORT) {
                                if (dirty & 1u) i1 = c1;
                                if (dirty & 2u) i2 = c2;
                                if (dirty & 4u) i3 = c3;
                            } else {
                                if (dirty & 1u) i1 = match8(cfg, 0, args.lut, code1, da0, da1);
                                if (dirty & 2u) i2 = match8(cfg, 1, args.lut, code2, db0, db1);
                                if (dirty & 4u) i3 = match8(cfg, 2, args.lut, code3, dc0, dc1);
                            }
                        }
                    }
                    matched = i1 >= 0 && i2 >= 0 && i3 >= 0;  // V2:310-321: all three lists non-empty (the mate check follows in commit())
                    if (matched) {
                        const uint32_t cell = cell_of(cfg, i1, i2, i3);
                        // UMI = lineRead[umi_start:umi_end] verbatim (V2:295): what the sequence holds of it
                        uint32_t up, ul;
                        if constexpr (SHORT) {
                            const uint32_t ulo = min((uint32_t)cfg.us, seq_len), uhi = min((uint32_t)cfg.ue, seq_len);
                            up = s1 + ulo;
                            ul = uhi - ulo;
                        } else {
                            up = s1 + (uint32_t)cfg.us;
                            ul = (uint32_t)cfg.umi_len;
                        }
                        const uint32_t i0 = up >> 2, sh = (up & 3u) * 8u;
                        const uint32_t w0 = words[i0], w1 = words[i0 + 1], w2 = words[i0 + 2], w3 = words[i0 + 3];
                        const uint32_t m0 = ul >= 4u ? 0xFFFFFFFFu : (1u << (8u * ul)) - 1u;
                        const uint32_t m1 = ul >= 8u ? 0xFFFFFFFFu : (ul > 4u ? (1u << (8u * (ul - 4u))) - 1u : 0u);
                        const uint32_t m2 = ul > 8u ? (1u << (8u * (ul - 8u))) - 1u : 0u;
                        const uint32_t u0 = __funnelshift_r(w0, w1, sh) & m0, u1 = __funnelshift_r(w1, w2, sh) & m1,
                                       u2 = __funnelshift_r(w2, w3, sh) & m2;
                        uint32_t e0, e1, e2;
                        const uint64_t ucode = ((uint64_t)pack4(u0, e0) | ((uint64_t)pack4(u1, e1) << 8) | ((uint64_t)pack4(u2, e2) << 16)) &
                                               ((1ull << cfg.umi_bits) - 1ull);
                        if (ul == (uint32_t)cfg.umi_len && ((e0 & m0) | (e1 & m1) | (e2 & m2)) == 0u) {
                            o.ann = kAnnRegular | ((uint64_t)cell << kAnnCellShift) | ucode;
                        } else {
                            o.ann = kAnnIrregular | ((uint64_t)cell << kAnnCellShift) | ((uint64_t)(tile_lo + up) << 5) | (uint64_t)ul;
                            o.irr.hi = ((uint64_t)cell << kIrrCellShift) | ((uint64_t)ul << 32) | (uint64_t)__byte_perm(u0, 0u, 0x0123u);
                            o.irr.lo = ((uint64_t)__byte_perm(u1, 0u, 0x0123u) << 16) | (uint64_t)(((u2 & 0xFFu) << 8) | ((u2 >> 8) & 0xFFu));
                        }
                    }
                }
            }
        };

        auto commit = [&](uint64_t r) {
            const bool ok2 = ok && r < args.rec_cap;  // beyond the capacity: the exact path raises the overflow flag
            if (MODE == kScanR2) settle();
            if (MODE == kScanR1) {
                if (ok2) {
                    if (noat) raise(args.cnt, args.err_off, 0, kErrNoAt, tile_lo + s0);
                    args.r1_start[r] = (OffT)(tile_lo + s0);
                    args.r1_meta[r] = mt;
                    args.r1_base[r] = base_len;
                    args.r1_tok[r] = fp;
                }
                __syncwarp();
                if (have && !ok2) r1_record_generic(ga, args, r, tile_lo + s0, cfg.header_style);
                __syncwarp();
            } else if (MODE == kScanAnn) {
                if (ok2) {
                    if (err_code) raise(args.cnt, args.err_off, 0, err_code, tile_lo + s0);
                    args.r1_start[r] = (OffT)(tile_lo + s0);
                    args.r1_meta[r] = mt;
                    args.r1_base[r] = base_len;
                    args.ann[r] = an;
                } else {
                    an = 0;
                }
                __syncwarp();
                if (have && !ok2) an = ann_record_generic(ga, args, cfg, r, tile_lo + s0);
                __syncwarp();
                const uint32_t st8 = ann_state(an);
                if (st8) atomicAdd(&args.hist[ann_cell(an)], 1u);
                const uint32_t nreg = __popc(__ballot_sync(0xFFFFFFFFu, st8 == 1u)), nirr = __popc(__ballot_sync(0xFFFFFFFFu, st8 == 2u));
                if (lane == 0 && nreg) atomicAdd(&args.cnt->n_keys, (unsigned long long)nreg);
                if (lane == 0 && nirr) atomicAdd(&args.cnt->n_irr, (unsigned long long)nirr);
            } else {
                if (ok2) {
                    if (noat) raise(args.cnt, args.err_off, 1, kErrNoAt, tile_lo + s0);
                    if (matched) {
                        // V2:349-350: the read-1 mate must carry the same id (fingerprint compare).  A mismatch fails the whole
                        // call, so nothing below waits for the fingerprint: it is requested here and compared one tile later.
                        if (r + 1u <= n_rec_r1) {
                            pend_fp1 = args.r1_tok[r];
                            pend_fp = fp;
                            pend_off = tile_lo + s0;
                            pend = true;
                        } else {
                            raise(args.cnt, args.err_off, 1, kErrKey, tile_lo + s0);
                        }
                    }
                } else {
                    o.ann = 0;
                    complete = false;
                }
                __syncwarp();
                if (have && !ok2) {
                    atomicAdd(&args.cnt->n_exact, 1u);
                    r2_record_generic(ga, args, cfg, r, tile_lo + s0, o, complete, n_rec_r1);
                }
                __syncwarp();
                // ---- results: annotation, histogram, key counts; the regular (cell, UMI) key is the annotation itself
                //      (the sort reads it from there), irregular keys are collected afterwards by k_collect_irr ----
                if (complete) args.ann[r] = o.ann;
                const uint32_t st8 = ann_state(o.ann);
                if (st8) atomicAdd(&args.hist[ann_cell(o.ann)], 1u);
                const uint32_t nreg = __popc(__ballot_sync(0xFFFFFFFFu, st8 == 1u)), nirr = __popc(__ballot_sync(0xFFFFFFFFu, st8 == 2u));
                if (lane == 0 && nreg) atomicAdd(&args.cnt->n_keys, (unsigned long long)nreg);
                if (lane == 0 && nirr) atomicAdd(&args.cnt->n_irr, (unsigned long long)nirr);
            }
        };

        // records are packed into as few warps as possible (fewer warp instructions); SSD_SCAN_R2_SPREAD=1 deals read 2's records
        // round robin to the warps instead, which ends the stage sooner but costs a third more instructions
        const uint32_t my_i = (MODE == kScanR2 && SSD_SCAN_R2_SPREAD) ? (uint32_t)lane * kBGWarps + (uint32_t)wb : (uint32_t)tb;

        // the guess of the record phase: made by the guess warp
        const uint32_t j_guess = S.guess_j[ls], n_guess = S.guess_n[ls];
        if (DEFER) {
            // ---- park the results; the commit warp writes them out once the tile's first record number is known ----
            const bool guessed = j_guess != 0xFFFFFFFFu && n_guess <= (uint32_t)kBThreads;
            if (guessed && n_guess != 0u) {
                compute(my_i, j_guess, n_guess);
            } else {
                have = ok = complete = noat = matched = false;
                err_code = 0;
                an = 0;
                o.ann = 0;
            }
            PROF_T(t6a);
            PROF_ADD(6, t8, t6a);
            if (it >= (uint32_t)kStageSlots) mbar_wait(reinterpret_cast<uint64_t*>(&S.stage_empty[ss]), ((it / kStageSlots) - 1u) & 1u);
            PROF_T(t6b);
            PROF_ADD(3, t6a, t6b);
            StageSlot& G = S.stage[ss];
            SSD_CHECK(my_i < (uint32_t)kBThreads && ss < (uint32_t)kStageSlots && (!have || n_guess <= (uint32_t)kBThreads));
            const uint32_t fl = (have ? kStHave : 0u) | (ok ? kStOk : 0u) | (noat ? kStNoAt : 0u) | (matched ? kStMatched : 0u) |
                                (complete ? kStComplete : 0u) | (err_code << 8);
            G.s0f[my_i] = s0 | (fl << 16);
            G.fp[my_i] = fp;
            if (MODE != kScanR2) {
                G.mt[my_i] = mt;
                G.base[my_i] = base_len;
            }
            if (MODE != kScanR1) {
                // what does not depend on the record number happens here: histogram, key counts
                const uint64_t a = !ok ? 0ull : (MODE == kScanR2 ? o.ann : an);
                G.ann[my_i] = a;
                const uint32_t st8 = ann_state(a);
                if (st8) atomicAdd(&args.hist[ann_cell(a)], 1u);
                const uint32_t nreg = __popc(__ballot_sync(0xFFFFFFFFu, st8 == 1u)), nirr = __popc(__ballot_sync(0xFFFFFFFFu, st8 == 2u));
                if (lane == 0 && nreg) atomicAdd(&args.cnt->n_keys, (unsigned long long)nreg);
                if (lane == 0 && nirr) atomicAdd(&args.cnt->n_irr, (unsigned long long)nirr);
            }
            if (tb == 0) {
                S.stage_tile[ss] = tile;
                S.stage_total[ss] = total;
                S.stage_guess[ss] = guessed ? (j_guess | (n_guess << 8)) : 0xFFFFFFFFu;
                S.stage_flags[ss] = (n_listed + 1u > (uint32_t)kMaxLines ? 1u : 0u) |
                                    ((tile == args.n_tiles - 1 && args.len && s_win[win_bytes - 1] != '\n') ? 2u : 0u);
            }
            sync_b(bgroup);
            PROF_T(t6d);
            PROF_ADD(4, t8, t6d);
            if (tb == 0) {
                mbar_arrive(&S.stage_full[ss]);  // over to the commit warp
                mbar_arrive(&S.empty[ws]);       // the window may be refilled
                mbar_arrive(&S.list_empty[ls]);  // and so may the list slot
            }
            return true;
        }
        const bool speculate = SSD_SCAN_SPECULATE(MODE) && j_guess != 0xFFFFFFFFu && n_guess != 0u && n_guess <= (uint32_t)kBThreads;
        if (speculate) compute(my_i, j_guess, n_guess);

        PROF_T(t4);
        PROF_ADD(6, t8, t4);
        if (tb == 0) {
            // the inclusive newline prefix of this tile, written by the prefix warp
            unsigned long long v;
            uint32_t ns = 20;
            while (((v = st[tile]) >> 62) != 2ull) {
                __nanosleep(ns);
                if (ns < 200u) ns += ns;
            }
            PROF_T(t10);
            PROF_ADD(8, t4, t10);
            const unsigned long long excl_lines = (v & kStatusValue) - total;
            S.line_base[bgroup] = excl_lines;
            if (tile == args.n_tiles - 1) {
                unsigned long long n_lines = excl_lines + total + ((args.len && s_win[win_bytes - 1] != '\n') ? 1ull : 0ull);
                args.cnt->n_lines[MODE == kScanR2 ? 1 : 0] = n_lines;
                args.cnt->n_records[MODE == kScanR2 ? 1 : 0] = n_lines >> 2;
            }
        }
        sync_b(bgroup);
        PROF_T(t5);
        PROF_ADD(3, t4, t5);
        const unsigned long long line_base = S.line_base[bgroup];

        // records that start in this tile: the line after a newline whose next-line index is 0 mod 4
        const unsigned long long r_first = tile == 0 ? 0ull : (line_base + 4) >> 2;
        const unsigned long long r_end = ((line_base + total) >> 2) + 1;  // exclusive
        uint32_t n_rec = r_end > r_first ? (uint32_t)(r_end - r_first) : 0u;
        if (tile == 0 && args.len == 0) n_rec = 0;
        if (n_rec > (uint32_t)kMaxRecTile || n_listed + 1u > (uint32_t)kMaxLines) {
            if (tb == 0) atomicOr(&args.cnt->err_flags, kErrDense);
            n_rec = 0;
        }
        const uint32_t j_first = (uint32_t)(4ull * r_first - line_base);  // list index of the first record's start

        if (speculate && j_first == j_guess && n_rec == n_guess) {
            commit(r_first + my_i);
        } else {
            for (uint32_t base = 0; base < n_rec; base += kBThreads) {
                compute(base + my_i, j_first, n_rec);
                commit(r_first + base + my_i);
            }
        }
        sync_b(bgroup);
        PROF_T(t6);
        PROF_ADD(4, t5, t6);
        if (tb == 0) {
            mbar_arrive(&S.empty[ws]);       // the window may be refilled
            mbar_arrive(&S.list_empty[ls]);  // and so may the list slot
        }
        return true;
    };

    uint32_t n_a = 0;
    if (warp < kAWarps) {
        while (stage_a(n_a)) n_a++;
    } else {
        n_a = (uint32_t)bgroup;
        while (stage_b(n_a)) n_a += (uint32_t)kBGroups;
        if (MODE == kScanR2) settle();
    }
#ifdef SSD_SCAN_PROFILE
    if (tid == 0 || tb == 0) {
        for (int k = 0; k < 16; k++)
            if (k != 7) atomicAdd(&g_scan_prof[MODE == kScanR2 ? 1 : 0][k], prof[k]);
        if (tid == 0) atomicAdd(&g_scan_prof[MODE == kScanR2 ? 1 : 0][7], (unsigned long long)n_a);
    }
#endif
}

}  // namespace ssd
