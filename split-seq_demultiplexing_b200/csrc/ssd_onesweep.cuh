// ssd_onesweep.cuh — per-cell distinct-UMI counts (the len(set(...)) of V2:411) without unique-key lists, sm_100a.
//
// Regular keys (cell << umi_bits | 2-bit UMI, <= 48 bits): LSD radix sort with 8-bit digits where every pass is ONE
// kernel.  The digit histograms of all passes come from a single read of the keys; a pass then assigns tiles in
// launch order (ticket), ranks its 4096 keys per digit (match.any), learns the digit offsets of the earlier tiles by a
// decoupled look-back over per-tile digit counts, reorders the tile in shared memory and writes digit runs
// contiguously.  The first pass reads the annotation words directly (compaction for free).  A last kernel counts the
// first occurrences per cell in the sorted keys, one atomic per (warp, cell).
// On one GPU (and for the keys a rank receives for its cells) the sort is not needed at all: k_umi_sets below inserts
// the keys into per-cell hash sets sized from the per-cell key counts; the sort remains for hosts that want the
// sorted unique lists (ssd_local_unique_keys) and behind SSD_DISTINCT_SORT=1.
// Irregular keys (UMI with non-ACGT bytes or short; raw bytes, 128 bits) are few: an open-addressing hash set with
// 128-bit compare-and-swap, exact, one kernel, on the side stream.
#pragma once
#include "ssd_sort.cuh"

namespace ssd {

#ifndef SSD_OS_MINBLOCKS
#define SSD_OS_MINBLOCKS 4
#endif
constexpr int kOsBits = 8;
constexpr int kOsRadix = 1 << kOsBits;
constexpr int kOsItems = 16;
constexpr int kOsTile = kThreads * kOsItems;  // keys per tile
constexpr int kOsMaxPasses = 6;
constexpr uint32_t kOsAgg = 1u << 30, kOsPrefix = 2u << 30, kOsValue = (1u << 30) - 1u;
static_assert(kThreads == kOsRadix, "one thread per digit");

// digit histograms of all passes in one read of the keys
template <class Src>
__global__ void __launch_bounds__(kThreads) k_os_hist(Src keys, uint64_t n, int passes, uint32_t* ghist, int shift0)
{
    __shared__ uint32_t s_h[kOsMaxPasses][kOsRadix];
    for (int j = threadIdx.x; j < kOsMaxPasses * kOsRadix; j += kThreads) (&s_h[0][0])[j] = 0;
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (uint64_t)gridDim.x * kThreads) {
        const uint64_t k = keys(i);
        if (k != kNoKey) {
#pragma unroll
            for (int p = 0; p < kOsMaxPasses; p++)
                if (p < passes) atomicAdd(&s_h[p][(uint32_t)(k >> (shift0 + kOsBits * p)) & (kOsRadix - 1)], 1u);
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < passes * kOsRadix; j += kThreads) {
        const uint32_t v = (&s_h[0][0])[j];
        if (v) atomicAdd(&ghist[j], v);
    }
}

// one pass: stable scatter of the keys by the digit at `shift`
template <class Src>
__global__ void __launch_bounds__(kThreads, SSD_OS_MINBLOCKS) k_os_pass(Src keys_in, uint64_t n, int shift, uint64_t* keys_out, const uint32_t* ghist,
                                                      uint32_t* status, uint32_t* ticket)
{
    __shared__ uint64_t s_keys[kOsTile];
    __shared__ uint32_t s_wcount[kThreads / 32][kOsRadix];
    __shared__ uint32_t s_lstart[kOsRadix];
    __shared__ unsigned long long s_base[kOsRadix];
    __shared__ unsigned long long s_wsum[kThreads / 32];
    __shared__ uint32_t s_tile, s_nvalid;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);  // tiles are taken in launch order: what a tile looks back at is running or done
    for (int i = tid; i < (kThreads / 32) * kOsRadix; i += kThreads) (&s_wcount[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;

    // warp w owns keys [w * 512, (w + 1) * 512) of the tile; iteration i, lane l -> key w * 512 + i * 32 + l
    const uint64_t wbase = (uint64_t)tile * kOsTile + (uint64_t)warp * (32 * kOsItems);
    uint64_t key[kOsItems];
    uint16_t rank[kOsItems], dig[kOsItems];
#pragma unroll
    for (int i = 0; i < kOsItems; i++) {
        const uint64_t j = wbase + (uint64_t)i * 32 + lane;
        bool valid = j < n;
        key[i] = valid ? keys_in(j) : kNoKey;
        valid = key[i] != kNoKey;
        const uint32_t d = valid ? (uint32_t)(key[i] >> shift) & (kOsRadix - 1) : (uint32_t)kOsRadix;
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, d);
        const uint32_t leader = (uint32_t)__ffs((int)peers) - 1u;
        uint32_t pre = 0;
        if ((uint32_t)lane == leader && valid) pre = s_wcount[warp][d];
        pre = __shfl_sync(0xFFFFFFFFu, pre, (int)leader);
        __syncwarp();
        if ((uint32_t)lane == leader && valid) s_wcount[warp][d] = pre + (uint32_t)__popc(peers);
        __syncwarp();
        rank[i] = (uint16_t)(pre + (uint32_t)__popc(peers & ((1u << lane) - 1u)));
        dig[i] = (uint16_t)d;
    }
    __syncthreads();

    // thread d: the tile's count of digit d (warp counts become exclusive offsets), then the exclusive scans over
    // the digits of the tile counts (low word) and of the global histogram (high word) in one go
    uint32_t cnt = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; w++) {
        const uint32_t cw = s_wcount[w][tid];
        s_wcount[w][tid] = cnt;
        cnt += cw;
    }
    const unsigned long long mine = ((unsigned long long)ghist[tid] << 32) | cnt;
    unsigned long long incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long u = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += u;
    }
    if (lane == 31) s_wsum[warp] = incl;
    // publish the tile count of this digit right away: later tiles are waiting for it
    volatile uint32_t* st = status + (size_t)tile * kOsRadix + tid;
    if (tile != 0) *st = kOsAgg | cnt;
    __syncthreads();
    unsigned long long before = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; w++)
        if (w < warp) before += s_wsum[w];
    const unsigned long long excl = before + incl - mine;
    const uint32_t lstart = (uint32_t)excl, gstart = (uint32_t)(excl >> 32);
    s_lstart[tid] = lstart;
    if (tid == kOsRadix - 1) s_nvalid = lstart + cnt;

    // decoupled look-back: keys with this digit in all earlier tiles
    uint32_t prior = 0;
    for (uint32_t t = tile; t > 0;) {
        const uint32_t f = *reinterpret_cast<volatile const uint32_t*>(&status[(size_t)(t - 1) * kOsRadix + tid]);
        if ((f >> 30) == 0u) continue;  // not published yet
        prior += f & kOsValue;
        if ((f >> 30) == 2u) break;     // an inclusive prefix: everything before it is counted in
        t--;
    }
    *st = kOsPrefix | (prior + cnt);
    s_base[tid] = (unsigned long long)gstart + prior - lstart;  // global index of the tile-sorted position 0, per digit (wraps, fine)
    __syncthreads();

    // tile-sorted order in shared memory, then digit runs go out contiguously
#pragma unroll
    for (int i = 0; i < kOsItems; i++) {
        const uint32_t d = dig[i];
        if (d < (uint32_t)kOsRadix) s_keys[s_lstart[d] + s_wcount[warp][d] + rank[i]] = key[i];
    }
    __syncthreads();
    const uint32_t nvalid = s_nvalid;
#pragma unroll
    for (int i = 0; i < kOsItems; i++) {
        const uint32_t idx = (uint32_t)tid + (uint32_t)i * kThreads;
        if (idx < nvalid) {
            const uint64_t k = s_keys[idx];
            keys_out[s_base[(uint32_t)(k >> shift) & (kOsRadix - 1)] + idx] = k;
        }
    }
}

// sorted keys -> distinct[cell] += number of first occurrences; one atomic per (warp, cell)
__global__ void __launch_bounds__(kThreads) k_count_sorted(const uint64_t* keys, uint64_t n, int shift, uint32_t* distinct)
{
    const uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool valid = i < n;
    const uint64_t k = valid ? keys[i] : 0ull;
    const bool head = valid && (i == 0 || keys[i - 1] != k);
    const uint32_t cell = valid ? (uint32_t)(k >> shift) : 0xFFFFFFFFu;
    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, cell);
    const uint32_t heads = __ballot_sync(0xFFFFFFFFu, head);
    if (valid && (uint32_t)lane == (uint32_t)__ffs((int)peers) - 1u) {
        const uint32_t c = (uint32_t)__popc(heads & peers);
        if (c) atomicAdd(&distinct[cell], c);
    }
}

// ---------------------------------------------------------------------------------------------
// regular keys, without sorting: one exact open-addressing set PER CELL
// ---------------------------------------------------------------------------------------------
// A cell with h keys (its entry of the read histogram, or of a count of the keys at hand) owns the 2 h consecutive
// 32-bit slots starting at seg[cell] (seg = exclusive scan of the per-cell set sizes), so every set is at most half
// full and a slot only has to hold the UMI (<= 20 bits, stored + 1; zero = empty).  A cell with so many keys that one
// bit per possible UMI is smaller than that (2 h > 4^umi_len / 32 words) gets the bitmap instead.  The first insertion
// of a (cell, UMI) adds one to distinct[cell]; that is the len(set(...)) of V2:411 with one read of the keys and no
// ordering at all.
struct UmiSetWords {
    const uint32_t* counts;
    uint32_t bitmap_words;
    __device__ __forceinline__ uint32_t operator()(uint64_t i) const
    {
        const uint64_t w = 2ull * counts[i];
        return w > bitmap_words ? bitmap_words : (uint32_t)w;
    }
};
// (counted through a small direct-mapped table in shared memory per block, like the `distinct` updates of k_umi_sets
// below: the largest cell holds a fifth of the keys a rank receives, and a million and a half atomics on one address take
// 0.7 ms by themselves)
template <class Src>
__global__ void __launch_bounds__(kThreads) k_cell_key_counts(Src keys, uint64_t n, int umi_bits, uint32_t* counts)
{
    __shared__ uint32_t s_cell[256], s_cnt[256];
    for (int j = threadIdx.x; j < 256; j += kThreads) {
        s_cell[j] = 0xFFFFFFFFu;
        s_cnt[j] = 0;
    }
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (uint64_t)gridDim.x * kThreads) {
        const uint64_t k = keys(i);
        if (k == kNoKey) continue;
        const uint32_t cell = (uint32_t)(k >> umi_bits);
        const uint32_t e = (cell * 0x9E3779B1u) >> 24;
        const uint32_t prev = atomicCAS(&s_cell[e], 0xFFFFFFFFu, cell);
        if (prev == 0xFFFFFFFFu || prev == cell) atomicAdd(&s_cnt[e], 1u);
        else atomicAdd(&counts[cell], 1u);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < 256; j += kThreads)
        if (s_cnt[j]) atomicAdd(&counts[s_cell[j]], s_cnt[j]);
}

// Cells with many keys would serialise their `distinct` updates on one address (the largest cell of the bench workload
// holds 8 % of the reads): a block collects the first occurrences of such cells in a small direct-mapped table in
// shared memory and adds them once at the end; a conflict in that table falls back to the global atomic.
constexpr uint32_t kHotCellKeys = 512;  // a cell with at least this many keys goes through the block's table
constexpr int kHotSlots = 256;
// The work per key is a chain of dependent long-latency accesses (key -> per-cell count and offset -> compare-and-swap), so
// the loop is software-pipelined: while key i is inserted, the count and offset of key i + 1 and key i + 2 itself are on
// their way.
template <class Src>
__global__ void __launch_bounds__(kThreads) k_umi_sets(Src keys, uint64_t n, int umi_bits, const uint32_t* __restrict__ counts,
                                                       const uint32_t* __restrict__ seg, uint32_t* table, uint32_t* distinct)
{
    const uint32_t bitmap_words = umi_bits >= 5 ? 1u << (umi_bits - 5) : 1u;
    __shared__ uint32_t s_cell[kHotSlots], s_cnt[kHotSlots];
    for (int j = threadIdx.x; j < kHotSlots; j += kThreads) {
        s_cell[j] = 0xFFFFFFFFu;
        s_cnt[j] = 0;
    }
    __syncthreads();
    const uint64_t G = (uint64_t)gridDim.x * kThreads;
    uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x;
    uint64_t k1 = i < n ? keys(i) : kNoKey;          // key i
    uint64_t k2 = i + G < n ? keys(i + G) : kNoKey;  // key i + 1 (in units of the grid stride)
    uint32_t cnt1 = 0, sg1 = 0;
    if (k1 != kNoKey) {
        cnt1 = __ldg(counts + (uint32_t)(k1 >> umi_bits));
        sg1 = __ldg(seg + (uint32_t)(k1 >> umi_bits));
    }
    for (; i < n; i += G) {
        const uint64_t k = k1;
        const uint32_t cnt = cnt1, sg = sg1;
        // next stage of the two keys behind this one
        k1 = k2;
        cnt1 = sg1 = 0;
        if (k1 != kNoKey) {
            cnt1 = __ldg(counts + (uint32_t)(k1 >> umi_bits));
            sg1 = __ldg(seg + (uint32_t)(k1 >> umi_bits));
        }
        k2 = i + 2 * G < n ? keys(i + 2 * G) : kNoKey;
        if (k == kNoKey || cnt == 0u) continue;  // cnt == 0 cannot happen: the key itself was counted
        const uint32_t cell = (uint32_t)(k >> umi_bits), umi = (uint32_t)k & ((1u << umi_bits) - 1u);
        const uint64_t slots = 2ull * cnt;
        uint32_t* set = table + sg;
        if (slots > (uint64_t)bitmap_words) {
            // a large cell: one bit per possible UMI is smaller than its hash set (and stays in the cache).  The bit is set
            // without waiting for the old value (RED); k_umi_bitmap_counts counts the bits of these cells afterwards.
            atomicOr(set + (umi >> 5), 1u << (umi & 31u));
            continue;
        }
        uint32_t h = umi * 0x9E3779B1u;
        h = (h ^ (h >> 15)) * 0x85EBCA77u;
        uint64_t s = ((uint64_t)h * slots) >> 32;
        const uint32_t val = umi + 1u;
        for (;;) {
            const uint32_t old = atomicCAS(set + s, 0u, val);
            if (old == val) break;  // seen before
            if (old == 0u) {        // first occurrence of this (cell, UMI)
                bool local = false;
                if (cnt >= kHotCellKeys) {
                    const uint32_t e = (cell * 0x9E3779B1u) >> 24;
                    static_assert(kHotSlots == 256, "8-bit slot index");
                    const uint32_t prev = atomicCAS(&s_cell[e], 0xFFFFFFFFu, cell);
                    local = prev == 0xFFFFFFFFu || prev == cell;
                    if (local) atomicAdd(&s_cnt[e], 1u);
                }
                if (!local) atomicAdd(&distinct[cell], 1u);
                break;
            }
            if (++s == slots) s = 0;
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < kHotSlots; j += kThreads)
        if (s_cnt[j]) atomicAdd(&distinct[s_cell[j]], s_cnt[j]);
}

// distinct[cell] += set bits of the bitmap of every cell that has one (the cells k_umi_sets did not count itself)
__global__ void __launch_bounds__(kThreads) k_umi_bitmap_counts(const uint32_t* __restrict__ counts, const uint32_t* __restrict__ seg,
                                                                const uint32_t* __restrict__ table, uint32_t n_cells, int umi_bits, uint32_t* distinct)
{
    const uint32_t bitmap_words = umi_bits >= 5 ? 1u << (umi_bits - 5) : 1u;
    __shared__ uint32_t s_list[kThreads], s_n, s_sum;
    for (uint32_t base = blockIdx.x * kThreads; base < n_cells; base += gridDim.x * kThreads) {
        if (threadIdx.x == 0) s_n = 0;
        __syncthreads();
        const uint32_t cell = base + threadIdx.x;
        if (cell < n_cells && 2ull * counts[cell] > (uint64_t)bitmap_words) s_list[atomicAdd(&s_n, 1u)] = cell;
        __syncthreads();
        const uint32_t n = s_n;
        for (uint32_t k = 0; k < n; k++) {
            const uint32_t c = s_list[k];
            const uint32_t* bm = table + seg[c];
            uint32_t sum = 0;
            for (uint32_t w = threadIdx.x; w < bitmap_words; w += kThreads) sum += (uint32_t)__popc(bm[w]);
            sum = __reduce_add_sync(0xFFFFFFFFu, sum);
            if (threadIdx.x == 0) s_sum = 0;
            __syncthreads();
            if ((threadIdx.x & 31) == 0 && sum) atomicAdd(&s_sum, sum);
            __syncthreads();
            if (threadIdx.x == 0) atomicAdd(&distinct[c], s_sum);
            __syncthreads();
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// irregular keys: exact hash set, 128-bit compare-and-swap (ATOMG.E.CAS.128)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cas128(void* p, unsigned long long cmp_lo, unsigned long long cmp_hi, unsigned long long new_lo,
                                       unsigned long long new_hi, unsigned long long& old_lo, unsigned long long& old_hi)
{
    asm volatile(
        "{\n\t.reg .b128 c, s, o;\n\tmov.b128 c, {%2, %3};\n\tmov.b128 s, {%4, %5};\n\tatom.global.cas.b128 o, [%6], c, s;\n\tmov.b128 {%0, %1}, o;\n\t}"
        : "=l"(old_lo), "=l"(old_hi)
        : "l"(cmp_lo), "l"(cmp_hi), "l"(new_lo), "l"(new_hi), "l"(p)
        : "memory");
}

// insert one irregular key; true when it was not in the set yet
__device__ __forceinline__ bool irr_insert(ulonglong2* table, uint32_t slot_mask, uint64_t hi, uint64_t lo)
{
    hi |= 1ull << 63;  // never all zero: zero means empty
    unsigned long long h = (hi ^ (lo * 0x9E3779B97F4A7C15ull)) * 0xFF51AFD7ED558CCDull;
    h ^= h >> 31;
    for (uint32_t s = (uint32_t)h & slot_mask;; s = (s + 1u) & slot_mask) {
        unsigned long long olo, ohi;
        cas128(&table[s], 0ull, 0ull, lo, hi, olo, ohi);
        if (olo == 0ull && ohi == 0ull) return true;
        if (olo == lo && ohi == hi) return false;
    }
}

// irregular keys given as an array (what peers sent): distinct[cell] += first occurrences
// world > 0: only keys whose cell this rank owns count (owner = ((cell >> shift) * world) / n_buckets); a key with hi == ~0
// is padding
__global__ void __launch_bounds__(kThreads) k_irr_hash_keys(const IrrKey* keys, uint64_t n, ulonglong2* table, uint32_t slot_mask, uint32_t* distinct,
                                                            int rank, int world, int shift, int n_buckets)
{
    const uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    const IrrKey k = keys[i];
    if (k.hi == ~0ull) return;
    const uint32_t cell = (uint32_t)(k.hi >> kIrrCellShift) & (uint32_t)kAnnCellMask;
    if (world > 0 && (int)(((uint64_t)(cell >> shift) * (uint64_t)world) / (uint64_t)n_buckets) != rank) return;
    if (irr_insert(table, slot_mask, k.hi, k.lo)) atomicAdd(&distinct[cell], 1u);
}

// table: slots of {lo, hi | 1 << 63}, zero = empty; n_slots is a power of two >= 2 * (number of irregular records)
__global__ void __launch_bounds__(kThreads) k_irr_hash(const uint64_t* ann, uint64_t n, const uint8_t* r2buf, ulonglong2* table, uint32_t slot_mask,
                                                       uint32_t* distinct)
{
    const uint64_t r = (uint64_t)blockIdx.x * kThreads + threadIdx.x;
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
    const uint32_t cell = ann_cell(a);
    hi |= ((uint64_t)cell << kIrrCellShift) | ((uint64_t)ulen << 32);
    if (irr_insert(table, slot_mask, hi, lo)) atomicAdd(&distinct[cell], 1u);  // first time this (cell, UMI) is seen
}

}  // namespace ssd
