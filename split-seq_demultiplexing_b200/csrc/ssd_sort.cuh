// ssd_sort.cuh — first-hit LUT builder, device-wide exclusive scan, LSD radix sort, unique compaction and
// per-cell distinct counting (the len(set(UMIs)) of V2:411).
#pragma once
#include "ssd_device.cuh"

namespace ssd {

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

// Multi-GPU -m decision: a cell is UNDECIDED when no rank alone holds >= m distinct UMIs of it (dmax[cell] < m, the all-reduced
// MAX of the ranks' local counts) although all ranks together hold >= m reads of it (reads[cell], the all-reduced SUM).  Only
// the (cell, UMI) keys of such cells have to be exchanged: every rank holds fewer than m distinct ones per cell.  This gathers
// this batch's keys of the undecided cells (duplicates kept): regular keys and irregular keys, one warp-aggregated cursor each.
__global__ void __launch_bounds__(256) k_undecided_keys(const uint64_t* __restrict__ ann, uint64_t n, const uint8_t* r2buf, const uint32_t* __restrict__ reads,
                                                        const uint32_t* __restrict__ dmax, uint32_t m, int umi_bits, uint64_t* keys_out, IrrKey* irr_out,
                                                        unsigned long long* cursors)
{
    const int lane = threadIdx.x & 31;
    for (uint64_t base = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) & ~31ull; base < n; base += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = base + lane;
        uint64_t a = r < n ? ann[r] : 0ull;
        uint32_t st = ann_state(a);
        if (st) {
            const uint32_t cell = ann_cell(a);
            if (!(__ldg(dmax + cell) < m && __ldg(reads + cell) >= m)) st = 0;
        }
        const uint32_t reg = __ballot_sync(0xFFFFFFFFu, st == 1u), irr = __ballot_sync(0xFFFFFFFFu, st == 2u);
        unsigned long long b0 = 0, b1 = 0;
        if (lane == 0) {
            if (reg) b0 = atomicAdd(&cursors[0], (unsigned long long)__popc(reg));
            if (irr) b1 = atomicAdd(&cursors[1], (unsigned long long)__popc(irr));
        }
        b0 = __shfl_sync(0xFFFFFFFFu, b0, 0);
        b1 = __shfl_sync(0xFFFFFFFFu, b1, 0);
        const uint32_t below = (1u << lane) - 1u;
        if (st == 1u) keys_out[b0 + (uint32_t)__popc(reg & below)] = ((uint64_t)ann_cell(a) << umi_bits) | ann_payload(a);
        if (st == 2u) {
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
            irr_out[b1 + (uint32_t)__popc(irr & below)] = k;
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
