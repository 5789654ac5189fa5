// ssd_sort.cuh — first-hit LUT builder, device-wide exclusive scan, LSD radix sort, unique compaction and
// per-cell distinct counting (the len(set(UMIs)) of V2:411).
#pragma once
#include "ssd_device.cuh"

namespace ssd {

// ---------------------------------------------------------------------------------------------
// first-hit LUT: lut[e'][code] for e' = 0..e
// ---------------------------------------------------------------------------------------------
// `round`: whose list; `lut`: that list's set of tables
__global__ void k_build_lut(const __grid_constant__ DevCfg cfg, int round, uint8_t* lut, int n_levels)
{
    uint32_t code = blockIdx.x * blockDim.x + threadIdx.x;
    if (code >= 65536u) return;
    for (int lev = 0; lev < n_levels; lev++) {
        uint32_t hit = 0xFFu;
        for (int i = 0; i < cfg.n_wl[round]; i++) {
            uint32_t x = code ^ cfg.wl_code[round][i];
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

// Small inputs (the digit counts of a short sort): the whole exclusive scan in one block, one launch.
constexpr int kSmallScanThreads = 1024;
constexpr uint32_t kSmallScanMax = 1u << 16;
__global__ void __launch_bounds__(kSmallScanThreads) k_scan_small_u32(const uint32_t* in, uint32_t n, uint32_t* out)
{
    __shared__ uint32_t s_warp[kSmallScanThreads / 32];
    const uint32_t per = (n + kSmallScanThreads - 1) / kSmallScanThreads;
    const uint32_t lo = threadIdx.x * per, hi = lo + per < n ? lo + per : n;
    uint32_t sum = 0;
    for (uint32_t i = lo; i < hi; i++) sum += in[i];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += u;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t v = s_warp[lane], w = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t u = __shfl_up_sync(0xFFFFFFFFu, w, d);
            if (lane >= d) w += u;
        }
        s_warp[lane] = w - v;  // exclusive prefix of the warp totals
    }
    __syncthreads();
    uint32_t run = s_warp[warp] + incl - sum;
    for (uint32_t i = lo; i < hi; i++) {
        const uint32_t v = in[i];
        out[i] = run;
        run += v;
    }
    if (threadIdx.x == kSmallScanThreads - 1) out[n] = run;
}

// ---------------------------------------------------------------------------------------------
// LSD radix sort, kRadixBits-bit digits, stable.  Pass = per-block digit histogram -> scan -> ranked scatter.
// ---------------------------------------------------------------------------------------------
#ifndef SSD_RADIX_BITS
#define SSD_RADIX_BITS 10
#endif
#ifndef SSD_SORT_ITEMS
#define SSD_SORT_ITEMS 8
#endif
constexpr int kRadixBits = SSD_RADIX_BITS;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kSortItems = SSD_SORT_ITEMS;
constexpr int kSortBlock = kThreads * kSortItems;  // keys per block

template <typename K>
__device__ __forceinline__ uint32_t digit_of(const K& k, int bit);
template <>
__device__ __forceinline__ uint32_t digit_of<uint64_t>(const uint64_t& k, int bit)
{
    return (uint32_t)(k >> bit) & (uint32_t)(kRadix - 1);
}

// Key sources of a sort pass: a plain array, or the annotation words (regular records only; the rest is skipped,
// which makes the first pass double as the compaction of the keys).
constexpr uint64_t kNoKey = ~0ull;
struct KeysArray {
    const uint64_t* p;
    __device__ __forceinline__ uint64_t operator()(uint64_t i) const { return p[i]; }
};
struct KeysFromAnn {
    const uint64_t* ann;
    int umi_bits;
    __device__ __forceinline__ uint64_t operator()(uint64_t i) const
    {
        const uint64_t a = ann[i];
        return ann_state(a) == 1 ? (((uint64_t)ann_cell(a) << umi_bits) | ann_payload(a)) : kNoKey;
    }
};

// the same, read with the streaming hint: the words are used once and should not push other data out of the L2
struct KeysFromAnnStream {
    const uint64_t* ann;
    int umi_bits;
    __device__ __forceinline__ uint64_t operator()(uint64_t i) const
    {
#ifdef SSD_UMI_NOSTREAM
        const uint64_t a = ann[i];
#else
        const uint64_t a = __ldcs(reinterpret_cast<const unsigned long long*>(ann) + i);
#endif
        return ann_state(a) == 1 ? (((uint64_t)ann_cell(a) << umi_bits) | ann_payload(a)) : kNoKey;
    }
};

template <class Src>
__global__ void __launch_bounds__(kThreads) k_sort_hist(Src keys, uint64_t n, int bit, uint32_t* counts, uint32_t n_blocks)
{
    __shared__ uint32_t s_hist[kRadix];
    for (int d = threadIdx.x; d < kRadix; d += kThreads) s_hist[d] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * kSortBlock;
#pragma unroll
    for (int i = 0; i < kSortItems; i++) {
        uint64_t j = base + (uint64_t)i * kThreads + threadIdx.x;
        if (j < n) {
            const uint64_t k = keys(j);
            if (k != kNoKey) atomicAdd(&s_hist[digit_of(k, bit)], 1u);
        }
    }
    __syncthreads();
    for (int d = threadIdx.x; d < kRadix; d += kThreads) counts[(size_t)d * n_blocks + blockIdx.x] = s_hist[d];
}

template <class Src, typename K, bool HAS_VALS>
__global__ void __launch_bounds__(kThreads) k_sort_scatter(Src keys_in, K* keys_out, const uint32_t* vals_in, uint32_t* vals_out,
                                                           uint64_t n, int bit, const uint32_t* offsets, uint32_t n_blocks)
{
    __shared__ uint32_t s_wcount[kThreads / 32][kRadix + 1];
    __shared__ uint32_t s_base[kRadix];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < (kThreads / 32) * (kRadix + 1); i += kThreads) (&s_wcount[0][0])[i] = 0;
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
            key[i] = keys_in(j);
            valid = key[i] != kNoKey;
            if (HAS_VALS) val[i] = vals_in[j];
        }
        uint32_t d = valid ? digit_of(key[i], bit) : (uint32_t)kRadix;
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
    for (int d = tid; d < kRadix; d += kThreads) {
        // exclusive scan over warps for digit d, plus the global base of (digit, block)
        uint32_t acc = 0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; w++) {
            uint32_t c = s_wcount[w][d];
            s_wcount[w][d] = acc;
            acc += c;
        }
        s_base[d] = offsets[(size_t)d * n_blocks + blockIdx.x];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kSortItems; i++) {
        uint32_t d = dig[i];
        if (d < (uint32_t)kRadix) {
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

// first (optional): first[u] = position of unique key u in the sorted list (the run lengths follow from it: k_run_lengths)
__global__ void k_compact64(const uint64_t* in, const uint32_t* pos, uint64_t n, uint64_t* out, uint32_t* first = nullptr)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i == 0 || in[i] != in[i - 1]) {
        out[pos[i]] = in[i];
        if (first) first[pos[i]] = (uint32_t)i;
    }
}

// reads[u] = length of the run of unique key u in a sorted list of n keys (n_uniq runs, first[u] = where run u starts)
__global__ void k_run_lengths(const uint32_t* first, uint64_t n_uniq, uint64_t n, uint32_t* reads)
{
    uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_uniq) return;
    reads[u] = (u + 1 < n_uniq ? first[u + 1] : (uint32_t)n) - first[u];
}

__global__ void k_compact_irr(const IrrKey* in, const uint32_t* perm, const uint32_t* pos, uint64_t n, IrrKey* out, uint32_t* first = nullptr)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    IrrKey a = in[perm[i]];
    bool head = i == 0;
    if (!head) {
        IrrKey b = in[perm[i - 1]];
        head = a.hi != b.hi || a.lo != b.lo;
    }
    if (head) {
        out[pos[i]] = a;
        if (first) first[pos[i]] = (uint32_t)i;
    }
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

// Irregular (cell, UMI) keys are rare: they are gathered after the scan from the annotation words, the UMI bytes
// re-read from read 2 (the annotation holds their offset and length).
__global__ void k_collect_irr(const uint64_t* ann, uint64_t n, const uint8_t* r2buf, IrrKey* out, unsigned long long* cursor)
{
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint64_t a = ann[r];
    if (ann_state(a) != 2) return;
    const uint64_t payload = ann_payload(a);
    const uint32_t ulen = (uint32_t)(payload & 31u);
    const uint8_t* p = r2buf + (size_t)(payload >> 5);
    uint64_t hi = 0, lo = 0;
    for (uint32_t j = 0; j < ulen; j++) {
        const uint64_t c = p[j];
        if (j < 4) hi |= c << (8 * (3 - j));
        else lo |= c << (8 * (9 - j));
    }
    IrrKey k;
    k.hi = ((uint64_t)ann_cell(a) << kIrrCellShift) | ((uint64_t)ulen << 32) | hi;
    k.lo = lo;
    out[atomicAdd(cursor, 1ull)] = k;
}

// Multi-GPU -m decision without shipping every key.  Per cell one 64-bit word travels through ONE all-reduce(SUM):
//   bits 0..39  reads of the cell in this batch            -> global reads
//   bits 40..63 1 when this rank alone counts >= m distinct -> number of ranks that decide "pass" by themselves
// A cell is UNDECIDED when no rank decides alone although all ranks together hold >= m reads of it; only the (cell, UMI)
// keys of such cells have to be exchanged, and every rank holds fewer than m distinct ones per cell.
constexpr int kPackPassShift = 40;
constexpr unsigned long long kPackReadsMask = (1ull << kPackPassShift) - 1ull;

__global__ void k_pack_decision(const uint32_t* __restrict__ hist, const uint32_t* __restrict__ distinct, uint64_t n_cells, uint32_t m, bool never,
                                unsigned long long* __restrict__ packed)
{
    const uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    packed[c] = (unsigned long long)hist[c] | ((!never && distinct[c] >= m) ? 1ull << kPackPassShift : 0ull);
}

// one bit per cell: undecided
__global__ void k_mark_undecided(const unsigned long long* __restrict__ packed, uint64_t n_cells, uint32_t m, uint32_t* __restrict__ bits)
{
    const uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool u = false;
    if (c < n_cells) {
        const unsigned long long p = packed[c];
        u = (p >> kPackPassShift) == 0ull && (p & kPackReadsMask) >= (unsigned long long)m;
    }
    const uint32_t w = __ballot_sync(0xFFFFFFFFu, u);
    if ((threadIdx.x & 31) == 0 && c < n_cells) bits[c >> 5] = w;
}

// the tables after the exchange: global reads; distinct = exact count for the undecided cells (already there), m for the cells
// some rank decided alone, untouched (zero) for the cells that fail on reads -- a lower bound that decides -m like the exact table
__global__ void k_adopt_decision(const unsigned long long* __restrict__ packed, uint64_t n_cells, uint32_t m, uint32_t* __restrict__ hist,
                                 uint32_t* __restrict__ distinct)
{
    const uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    const unsigned long long p = packed[c], reads = p & kPackReadsMask;
    hist[c] = reads > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)reads;
    if (p >> kPackPassShift) distinct[c] = max(distinct[c], m);
}

// This batch's keys of the undecided cells (duplicates kept) as ONE list of 128-bit records in the irregular-key format
// (cell, length, UMI bytes) -- a regular UMI is spelled out from its 2-bit code, so equal UMIs are equal records whichever
// way a rank happened to hold them.  A block takes 1024 consecutive records per round (four per thread: their annotation words
// and bitmap lookups are in flight together) and claims room for all its keys with ONE atomic on the cursor.  out[0] is the
// header (k_undecided_header), keys go to out[1 ..= cap]; the cursor keeps counting past cap so that the receivers see the
// overflow.
__global__ void __launch_bounds__(256) k_undecided_packed(const uint64_t* __restrict__ ann, uint64_t n, const uint8_t* r2buf, const uint32_t* __restrict__ und_bits,
                                                          int umi_len, IrrKey* out, unsigned long long cap, unsigned long long* cursor)
{
    __shared__ uint32_t s_warp[8];
    __shared__ unsigned long long s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint64_t chunk = blockIdx.x; chunk * 1024 < n; chunk += gridDim.x) {
        uint64_t a[4];
        uint32_t st[4], any[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint64_t r = chunk * 1024 + (uint64_t)u * 256 + threadIdx.x;
            a[u] = r < n ? __ldcs(reinterpret_cast<const unsigned long long*>(ann) + r) : 0ull;
        }
        uint32_t mine = 0;
#pragma unroll
        for (int u = 0; u < 4; u++) {
            st[u] = ann_state(a[u]);
            if (st[u]) {
                const uint32_t cell = ann_cell(a[u]);
                if (!((__ldg(und_bits + (cell >> 5)) >> (cell & 31u)) & 1u)) st[u] = 0;
            }
            any[u] = __ballot_sync(0xFFFFFFFFu, st[u] != 0u);
            mine += (uint32_t)__popc(any[u]);
        }
        if (lane == 0) s_warp[warp] = mine;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t total = 0;
#pragma unroll
            for (int w = 0; w < 8; w++) total += s_warp[w];
            s_base = total ? atomicAdd(cursor, (unsigned long long)total) : 0ull;
        }
        __syncthreads();
        unsigned long long b0 = s_base;
        for (int w = 0; w < warp; w++) b0 += s_warp[w];
        __syncthreads();  // s_warp / s_base are rewritten in the next round
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const unsigned long long slot = b0 + (unsigned long long)__popc(any[u] & ((1u << lane) - 1u));
            b0 += (unsigned long long)__popc(any[u]);
            if (!st[u] || slot >= cap) continue;
            uint64_t hi = 0, lo = 0;
            uint32_t ulen;
            if (st[u] == 1u) {
                const uint64_t code = ann_payload(a[u]);
                ulen = (uint32_t)umi_len;
                for (uint32_t j = 0; j < ulen; j++) {
                    const uint64_t c = (0x47544341u >> (8u * (uint32_t)((code >> (2u * j)) & 3u))) & 0xFFu;  // A C T G by code (pack4)
                    if (j < 4) hi |= c << (8 * (3 - j));
                    else lo |= c << (8 * (9 - j));
                }
            } else {
                const uint64_t payload = ann_payload(a[u]);
                ulen = (uint32_t)(payload & 31u);
                const uint8_t* p = r2buf + (size_t)(payload >> 5);
                for (uint32_t j = 0; j < ulen; j++) {
                    const uint64_t c = p[j];
                    if (j < 4) hi |= c << (8 * (3 - j));
                    else lo |= c << (8 * (9 - j));
                }
            }
            IrrKey k;
            k.hi = ((uint64_t)ann_cell(a[u]) << kIrrCellShift) | ((uint64_t)ulen << 32) | hi;
            k.lo = lo;
            out[1 + slot] = k;
        }
    }
}

// header of the packed list: hi = all ones (the counting kernels skip such records as padding), lo = number of keys the
// rank had (more than the capacity = the list is incomplete)
__global__ void k_undecided_header(IrrKey* out, const unsigned long long* cursor)
{
    out[0].hi = ~0ull;
    out[0].lo = *cursor;
}

// Shard phase of the multi-GPU file path: newlines per tile of 2^tile_log2 bytes of a loaded byte range.  One block
// per tile, 16-byte loads (the range starts 16-byte aligned in its buffer), the last bytes one by one.
__global__ void __launch_bounds__(256) k_count_newlines(const uint8_t* __restrict__ buf, uint64_t len, int tile_log2, uint32_t* __restrict__ counts)
{
    const uint64_t lo = (uint64_t)blockIdx.x << tile_log2;
    const uint64_t hi = lo + (1ull << tile_log2) < len ? lo + (1ull << tile_log2) : len;
    uint32_t n = 0;
    const uint64_t full = (hi - lo) & ~15ull;
    const uint4* v = reinterpret_cast<const uint4*>(buf + lo);
    for (uint64_t i = threadIdx.x; i < full / 16; i += blockDim.x) {
        const uint4 w = __ldcs(v + i);
        const uint32_t x[4] = {w.x ^ 0x0A0A0A0Au, w.y ^ 0x0A0A0A0Au, w.z ^ 0x0A0A0A0Au, w.w ^ 0x0A0A0A0Au};
#pragma unroll
        for (int k = 0; k < 4; k++) n += (uint32_t)__popc(~(((x[k] & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x[k] | 0x7F7F7F7Fu));  // exact zero-byte test
    }
    for (uint64_t i = lo + full + threadIdx.x; i < hi; i += blockDim.x) n += buf[i] == '\n';
    n = __reduce_add_sync(0xFFFFFFFFu, n);
    __shared__ uint32_t s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = n;
    __syncthreads();
    if (threadIdx.x == 0) counts[blockIdx.x] = s[0] + s[1] + s[2] + s[3] + s[4] + s[5] + s[6] + s[7];
}

// Position learner on the device (V2:164-199): over the first n_lines lines of read 2, for every sequence line (line index
// % 4 == 1) the position of the first occurrence of each of three static sequences in the strip()ped line (LIB:105 writes
// the sample strip()ped; str.find), -1 when absent.  One block: warp 0 lists the line starts, then one thread per sequence
// line searches.  found[k * cap + t]: anchor k in sequence line t; *n_seq = sequence lines seen.
constexpr int kLearnMaxLines = 4096;
__global__ void __launch_bounds__(256) k_anchor_positions(const uint8_t* __restrict__ text, unsigned long long len, uint32_t n_lines, const uint8_t* __restrict__ anchors,
                                                          const int* __restrict__ anchor_len, int32_t* __restrict__ found, uint32_t cap, uint32_t* __restrict__ n_seq)
{
    __shared__ unsigned long long start[kLearnMaxLines + 1];  // start[i] = first byte of line i; start[n] = end of the last line + 1
    __shared__ uint32_t s_lines;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < 32) {
        uint32_t n = 0;  // complete lines listed so far
        if (lane == 0) start[0] = 0;
        for (unsigned long long base = 0; base < len && n < n_lines; base += 32) {
            const unsigned long long p = base + lane;
            const uint32_t nl = __ballot_sync(0xFFFFFFFFu, p < len && text[p] == '\n');
            if (lane == 0) {
                uint32_t m = nl;
                while (m && n < n_lines) {
                    const uint32_t b = (uint32_t)__ffs((int)m) - 1u;
                    m &= m - 1u;
                    start[++n] = base + b + 1;
                }
            }
            n = __shfl_sync(0xFFFFFFFFu, n, 0);
        }
        if (lane == 0) {
            if (n < n_lines && start[n] < len) start[++n] = len + 1;  // a last line without its newline still is a line
            s_lines = n;
        }
    }
    __syncthreads();
    const uint32_t lines = s_lines, seq_lines = lines >= 2 ? (lines - 2) / 4 + 1 : 0;
    if (threadIdx.x == 0) *n_seq = seq_lines;
    for (uint32_t t = threadIdx.x; t < seq_lines && t < cap; t += blockDim.x) {
        unsigned long long lo = start[4 * t + 1], hi = start[4 * t + 2] - 1;  // [lo, hi) without the newline
        while (lo < hi && py_space(text[lo])) lo++;
        while (hi > lo && py_space(text[hi - 1])) hi--;
        for (int k = 0; k < 3; k++) {
            const int al = anchor_len[k];
            int32_t pos = -1;
            for (unsigned long long p = lo; p + al <= hi && pos < 0; p++) {
                bool eq = true;
                for (int j = 0; j < al && eq; j++) eq = text[p + j] == anchors[k * 16 + j];
                if (eq) pos = (int32_t)(p - lo);
            }
            if (al == 0) pos = 0;
            found[(size_t)k * cap + t] = pos;
        }
    }
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

}  // namespace ssd
