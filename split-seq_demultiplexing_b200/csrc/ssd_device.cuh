// ssd_device.cuh — device-side building blocks shared by the kernels of libssd_b200.so:
// configuration block, annotation word layout, byte accessors over a shared-memory window (with a
// newline bitmap) or over global memory, FASTQ line finding, barcode packing and matching.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ssd {

constexpr int kMaxWl = 128;
constexpr int kThreads = 256;

// ---------------------------------------------------------------------------------------------
// configuration passed to kernels by value (__grid_constant__)
// ---------------------------------------------------------------------------------------------
struct DevCfg {
    // Round k (0: the BC1 window, 1: BC2, 2: BC3) is matched against its own list.  The reference's fast path uses ONE list for
    // all three (V2:60, 302-304: the 8-mers of Round1_barcodes_new5.txt), which is three equal lists here; rounds with equal
    // lists share their lookup tables (lut_set).
    int n_wl[3], lut_set[3], e;
    int us, ue, b1s, b1e, b2s, b2e, b3s, b3e;  // Python slice bounds into the read-2 sequence (V2:155-162)
    int umi_len;      // ue - us (<= 10)
    int umi_bits;     // 2 * umi_len
    int cell_bits;    // ceil(log2(n1 n2 n3))
    int fast_wl;      // every whitelist entry (of every round) is 8 x ACGT -> 2-bit path + LUT
    int fast_geom;    // all three barcode windows are exactly 8 wide
    int need_len;     // largest slice end: shorter read-2 sequences take the exact generic path
    int bc8;          // every whitelist entry is exactly 8 bytes long
    int header_style; // 0: V2:134-141 (id up to the first '/', blanks removed, _cell_umi); 1: Extract_BC_UMI.py:78-85 (first token _cell_umi, blank, second token)
    unsigned long long magic_n, magic_nn;  // ceil(2^43 / n3), ceil(2^43 / (n2 n3)): exact division of cell ids < 2^21
    uint16_t wl_code[3][kMaxWl];
    uint8_t wl_len[3][kMaxWl];
    uint8_t wl[3][kMaxWl * 8];
};

// cell id of the three list indices: (i1 * n2 + i2) * n3 + i3
__device__ __forceinline__ uint32_t cell_of(const DevCfg& c, int i1, int i2, int i3) { return (uint32_t)((i1 * c.n_wl[1] + i2) * c.n_wl[2] + i3); }
// the (e + 1) x 65536 first-hit tables of a round; level `lvl` of them
__device__ __forceinline__ const uint8_t* lut_of(const DevCfg& c, const uint8_t* lut, int round, int lvl)
{
    return lut + ((size_t)c.lut_set[round] * (size_t)(c.e + 1) + (size_t)lvl) * 65536u;
}

// ---------------------------------------------------------------------------------------------
// per-record annotation word (one uint64 per read pair, written by the read-2 scan)
//   [63:62] state   0 = rejected, 1 = matched with a regular UMI, 2 = matched with an irregular UMI
//   [61:41] cell    (i1 * n2 + i2) * n3 + i3
//   [40:0]  payload regular:   2-bit packed UMI (umi_bits wide)
//                   irregular: (global offset of the UMI's first byte in read 2) << 5 | UMI length
// ---------------------------------------------------------------------------------------------
constexpr int kAnnCellShift = 41;
constexpr uint64_t kAnnPayloadMask = (1ull << kAnnCellShift) - 1;
constexpr uint64_t kAnnCellMask = (1ull << 21) - 1;
constexpr uint64_t kAnnRegular = 1ull << 62;
constexpr uint64_t kAnnIrregular = 2ull << 62;

__device__ __forceinline__ uint32_t ann_state(uint64_t a) { return (uint32_t)(a >> 62); }
__device__ __forceinline__ uint32_t ann_cell(uint64_t a) { return (uint32_t)((a >> kAnnCellShift) & kAnnCellMask); }
__device__ __forceinline__ uint64_t ann_payload(uint64_t a) { return a & kAnnPayloadMask; }
__device__ __forceinline__ uint32_t ann_umi_len(uint64_t a, int umi_len)
{
    return ann_state(a) == 1 ? (uint32_t)umi_len : (uint32_t)(ann_payload(a) & 31u);
}

// irregular (cell, UMI) key: hi = cell << 36 | len << 32 | bytes[0..4), lo = bytes[4..10)
struct IrrKey {
    uint64_t hi, lo;
};
constexpr int kIrrCellShift = 36;
constexpr int kIrrHiBits = 57;  // 21 + 4 + 32
constexpr int kIrrLoBits = 48;

// error flags raised by kernels (DevCounters::err_flags)
constexpr uint32_t kErrNoAt = 1u;         // first line of a complete record does not start with '@'
constexpr uint32_t kErrKey = 2u;          // matched read-2 record whose read-1 mate has another id / is missing
constexpr uint32_t kErrOverflow = 4u;     // more records than the tables were sized for (host retries)
constexpr uint32_t kErrDense = 8u;        // too many records in one tile (records shorter than 16 bytes on average)
constexpr uint32_t kErrEmitLen = 16u;     // emitted record length differs from the planned one (internal)
constexpr uint32_t kErrHeaderLong = 32u;  // header line longer than the supported maximum
constexpr uint32_t kErrIndex = 64u;       // annotated header without two '_' (the reference's IndexError, V2:397-398)
constexpr uint32_t kErrRespec = 256u;     // read-1 (<< 1: read-2) scan guessed the record phase of a tile wrong or met a tile with more
                                          // records than it can park: the host repeats that scan with the variant that waits instead
constexpr uint32_t kErrForeign = 128u;    // annotated header whose cell / UMI fields cannot come from this whitelist

struct DevCounters {
    unsigned long long n_lines[2];     // total lines of read 1 / read 2
    unsigned long long n_records[2];   // complete records
    unsigned long long n_keys;         // regular keys appended
    unsigned long long n_irr;          // irregular keys appended
    unsigned int err_flags;
    unsigned int ticket[2];            // dynamic tile tickets of the two scans
    unsigned int n_exact;              // read-2 records that left the fast path for the exact one (what the host picks the scan variant on)
    // produced by ssd_build_filter; zeroed as one block starting at n_matched
    unsigned long long n_matched, bytes_matched, bytes_passing, n_retained, n_cells, n_passing_cells;
};

// ---------------------------------------------------------------------------------------------
// characters
// ---------------------------------------------------------------------------------------------
// str.isspace() over ASCII: 0x09-0x0D, 0x1C-0x1F, 0x20 (what rstrip()/strip()/split() act on)
__device__ __forceinline__ bool py_space(uint32_t c) { return (c - 9u) <= 4u || (c - 28u) <= 4u; }

// 2-bit base code used by the matcher and the UMI key: A=0 C=1 T=2 G=3 ((c >> 1) & 3)
__device__ __forceinline__ bool is_acgt(uint32_t c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }
__device__ __forceinline__ uint32_t base2(uint32_t c) { return (c >> 1) & 3u; }
__device__ __forceinline__ char base2_char(uint32_t code) { return "ACTG"[code & 3u]; }

// ---------------------------------------------------------------------------------------------
// byte accessors.  Positions are global byte offsets into the file text.
// ---------------------------------------------------------------------------------------------
constexpr size_t kUnresolved = ~(size_t)0;

struct WinAcc {  // shared-memory window [g0, g0 + nbytes) with a newline bitmap (bit b of word w = byte 32w+b)
    const uint8_t* win;
    const uint32_t* mask;
    size_t g0;
    uint32_t nbytes;
    size_t len;  // file length
    __device__ __forceinline__ bool covers(size_t g) const { return g >= g0 && g - g0 < nbytes; }
    __device__ __forceinline__ uint32_t byte(size_t g) const { return win[g - g0]; }
    // position of the first '\n' at or after g; len when the file ends first; kUnresolved when the window does
    __device__ __forceinline__ size_t next_nl(size_t g) const
    {
        const bool at_eof = g0 + nbytes >= len;
        uint32_t p = (uint32_t)(g - g0);
        if (p >= nbytes) return at_eof ? len : kUnresolved;
        uint32_t w = p >> 5, nw = (nbytes + 31u) >> 5;
        uint32_t m = mask[w] & (0xFFFFFFFFu << (p & 31u));
        while (!m) {
            if (++w >= nw) return at_eof ? len : kUnresolved;
            m = mask[w];
        }
        return g0 + (size_t)(w << 5) + (uint32_t)(__ffs((int)m) - 1);
    }
};

struct GlobAcc {  // the whole file in global memory (slow path for records that outgrow a window)
    const uint8_t* buf;
    size_t len;
    __device__ __forceinline__ bool covers(size_t g) const { return g < len; }
    __device__ __forceinline__ uint32_t byte(size_t g) const { return buf[g]; }
    // 16 bytes per load once g is aligned (the buffer is 16-byte aligned, ssd_b200.h): the exact path of a record is a chain of
    // dependent loads, one per byte costs tens of microseconds per record
    __device__ __forceinline__ size_t next_nl(size_t g) const
    {
        while (g < len && (g & 15u)) {
            if (buf[g] == '\n') return g;
            g++;
        }
        while (g + 16 <= len) {
            const uint4 v = *reinterpret_cast<const uint4*>(buf + g);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t t = w[k] ^ 0x0A0A0A0Au;
                const uint32_t f = ~(((t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | t | 0x7F7F7F7Fu);  // 0x80 in every '\n' byte, exact
                if (f) return g + 4u * (uint32_t)k + (((uint32_t)__ffs((int)f) - 1u) >> 3);
            }
            g += 16;
        }
        while (g < len && buf[g] != '\n') g++;
        return g;
    }
};

// The four lines of a record: s[k] = first byte, e[k] = position of the terminating '\n' (or len).
struct Lines {
    size_t s[4], e[4];
    bool complete;  // line 3 exists  <=>  record index < n_lines / 4
};

// Returns false when the accessor cannot resolve the record (window ended).  need_e3: also find the
// end of the quality line.
template <class Acc>
__device__ __forceinline__ bool find_lines(const Acc& a, size_t start, size_t len, bool need_e3, Lines& L)
{
    size_t p = start;
    L.complete = false;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (p >= len) return true;  // line k does not exist: trailing partial record
        L.s[k] = p;
        if (k == 3) {
            L.complete = true;
            if (!need_e3) return true;
        }
        size_t nl = a.next_nl(p);
        if (nl == kUnresolved) return false;
        L.e[k] = nl;
        p = nl + 1;
    }
    return true;
}

template <class Acc>
__device__ __forceinline__ size_t rstrip_end(const Acc& a, size_t s, size_t e)
{
    while (e > s && py_space(a.byte(e - 1))) e--;
    return e;
}

// ---------------------------------------------------------------------------------------------
// barcode windows
// ---------------------------------------------------------------------------------------------
struct Packed {
    uint32_t code;   // 2 bits per position
    uint32_t valid;  // bit 2j set when position j < n holds A/C/G/T
    int n;           // bytes available (0..8)
};

template <class Acc>
__device__ __forceinline__ Packed pack_window(const Acc& a, size_t seq_s, size_t seq_len, int ws, int we, uint8_t* raw)
{
    Packed p;
    size_t lo = (size_t)ws < seq_len ? (size_t)ws : seq_len;
    size_t hi = (size_t)we < seq_len ? (size_t)we : seq_len;
    p.n = hi > lo ? (int)(hi - lo) : 0;
    if (p.n > 8) p.n = 8;
    p.code = 0;
    p.valid = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        uint32_t c = j < p.n ? a.byte(seq_s + lo + j) : 0u;
        raw[j] = (uint8_t)c;
        if (j < p.n && is_acgt(c)) {
            p.code |= base2(c) << (2 * j);
            p.valid |= 1u << (2 * j);
        }
    }
    return p;
}

// V2:302-304 + 327-329: index of the FIRST whitelist entry (file order) whose Hamming distance to the
// window over the overlapping prefix (zip() truncation, V2:72-73) is <= e; -1 when there is none.
// lut[code] holds that index for full-length all-ACGT windows (0xFF = none).
// `lut` = the round's level-e table (lut_of(c, base, round, c.e)).
__device__ __forceinline__ int match_window(const DevCfg& c, int round, const Packed& p, const uint8_t* raw, const uint8_t* __restrict__ lut)
{
    if (c.fast_wl) {
        if (p.n == 8 && p.valid == 0x5555u) {
            uint32_t v = __ldg(lut + p.code);
            return v == 0xFFu ? -1 : (int)v;
        }
        const uint32_t lenmask = p.n >= 8 ? 0x5555u : ((1u << (2 * p.n)) - 1u) & 0x5555u;
        const int base_d = __popc(lenmask & ~p.valid);  // non-ACGT bytes never equal a whitelist base
        if (base_d > c.e) return -1;
        for (int i = 0; i < c.n_wl[round]; i++) {
            uint32_t x = p.code ^ c.wl_code[round][i];
            int d = base_d + __popc((x | (x >> 1)) & p.valid);
            if (d <= c.e) return i;
        }
        return -1;
    }
    for (int i = 0; i < c.n_wl[round]; i++) {
        int n = p.n < (int)c.wl_len[round][i] ? p.n : (int)c.wl_len[round][i];
        int d = 0;
        for (int j = 0; j < n; j++) d += raw[j] != c.wl[round][i * 8 + j];
        if (d <= c.e) return i;
    }
    return -1;
}

// ---------------------------------------------------------------------------------------------
// async bulk copy (TMA, 1-D) global -> shared with an mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

constexpr uint32_t kMbarSuspendNs = 20000u;
#ifndef SSD_MBAR_SLEEP_NS
#define SSD_MBAR_SLEEP_NS 0
#endif
// SLEEP: nanoseconds to stay away from the issue slots after a failed try (0 = poll again at once).  Waiting warps that poll
// back to back take issue slots from the working ones: the polls were 12-19 % of the scans' warp instructions.
template <uint32_t SLEEP = SSD_MBAR_SLEEP_NS>
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity), "r"(kMbarSuspendNs)  // the thread may stay suspended this long before it polls again
            : "memory");
        if (ok) break;
        if (SLEEP) __nanosleep(SLEEP);
    }
}

// Load [src, src + nbytes) into shared memory at dst (both 16-byte aligned): one bulk copy for the
// 16-byte multiple, plain byte loads for the tail.  All threads call; returns after the data landed.
// `phase` is the mbarrier parity to wait for (0 for the first use of a freshly initialised barrier).
__device__ __forceinline__ void load_window(uint8_t* dst, const uint8_t* src, uint32_t nbytes, uint64_t* bar, uint32_t phase)
{
    const uint32_t bulk = nbytes & ~15u;
    if (threadIdx.x == 0 && bulk) {
        mbar_expect_tx(bar, bulk);
        bulk_g2s(dst, src, bulk, bar);
    }
    for (uint32_t i = bulk + threadIdx.x; i < nbytes; i += blockDim.x) dst[i] = src[i];
    if (bulk) mbar_wait(bar, phase);
    __syncthreads();
}

// 0x80 in every byte of x that is '\n' (exact for any byte values)
__device__ __forceinline__ uint32_t nl_flags(uint32_t x)
{
    uint32_t t = x ^ 0x0A0A0A0Au;
    return ~(((t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | t | 0x7F7F7F7Fu);
}

// Newline bitmap of win[0, nbytes): mask16[c] gets one bit per byte of 16-byte chunk c, for all `nchunks`
// chunks (zeros beyond the data).  The four flag words of a chunk are gathered with two dp4a each.
template <int NT, int ITERS>
__device__ __forceinline__ void build_nl_mask(const uint8_t* win, uint32_t nbytes, uint16_t* mask16, uint32_t nchunks)
{
    const uint4* w4 = reinterpret_cast<const uint4*>(win);
    const uint32_t nfull = nbytes >> 4;
#pragma unroll
    for (int it = 0; it < ITERS; it++) {
        const uint32_t c = (uint32_t)it * NT + threadIdx.x;
        if (c < nfull) {
            const uint4 v = w4[c];
            const uint32_t lo = __dp4a(nl_flags(v.x), 0x08040201u, __dp4a(nl_flags(v.y), 0x80402010u, 0u));
            const uint32_t hi = __dp4a(nl_flags(v.z), 0x08040201u, __dp4a(nl_flags(v.w), 0x80402010u, 0u));
            mask16[c] = (uint16_t)((lo >> 7) | ((hi >> 7) << 8));
        } else if (c < nchunks) {
            uint32_t m = 0;
            if (c == nfull && (nbytes & 15u)) {  // the chunk the data ends in
                const uint4 v = w4[c];
                const uint32_t lo = __dp4a(nl_flags(v.x), 0x08040201u, __dp4a(nl_flags(v.y), 0x80402010u, 0u));
                const uint32_t hi = __dp4a(nl_flags(v.z), 0x08040201u, __dp4a(nl_flags(v.w), 0x80402010u, 0u));
                m = ((lo >> 7) | ((hi >> 7) << 8)) & ((1u << (nbytes & 15u)) - 1u);
            }
            mask16[c] = (uint16_t)m;
        }
    }
}

// exact division of x < 2^21 by the divisor behind `magic` (= ceil(2^43 / d), d <= 2^14)
__device__ __forceinline__ uint32_t div_magic(uint32_t x, unsigned long long magic) { return (uint32_t)(((unsigned long long)x * magic) >> 43); }

__device__ __forceinline__ void cell_indices(const DevCfg& c, uint32_t cell, uint32_t& i1, uint32_t& i2, uint32_t& i3)
{
    const uint32_t n2 = (uint32_t)c.n_wl[1], n3 = (uint32_t)c.n_wl[2];
    i1 = div_magic(cell, c.magic_nn);
    const uint32_t rem = cell - i1 * n2 * n3;
    i2 = div_magic(rem, c.magic_n);
    i3 = rem - i2 * n3;
}

}  // namespace ssd
