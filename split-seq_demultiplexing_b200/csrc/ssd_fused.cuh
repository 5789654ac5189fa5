// ssd_fused.cuh — ONE pass over read 1 that finds its records AND writes the annotated FASTQ (sm_100a).
//
// Replaces, for the merged outputs, the read-1 scan + output planning + table-driven writer (three kernels, read 1
// crossing HBM twice, 32 bytes of per-record tables written and read back): V2:223-262 (read-1 parse), V2:349-358 (mate
// join), V2:134-141 + 367-370 / 415-454 (header rewrite, writers).  Read 2 has been scanned before (ssd_scan.cuh): its
// annotation word and id fingerprint per record, and the -m decision per cell, are inputs here.
//
// Persistent, warp-specialised, one block per SM, tiles of 16 KiB taken by ticket:
//   copy warp     keeps kFSlots tile windows (tile + 2 KiB halo) arriving with bulk copies (TMA, mbarrier complete_tx)
//   stage A       (8 warps) newline bitmap from 16-byte shared loads, tile newline count PUBLISHED (chain 1), list of line starts
//   prefix warp 1 (block 0) turns the published newline counts into inclusive prefixes, in tile order
//   stage B       (3 groups of 3 warps, tiles dealt round robin: the stage is latency bound) waits for the tile's line
//                 index, then one thread per record: header walk (id fingerprint, name.split("/")[0].replace(" ","")
//                 .strip(), squeezed IN PLACE in the window), read / quality bounds, the record's annotation + read 2's
//                 fingerprint (mate check, V2:350), pass decision, output length; exclusive scan of the lengths inside
//                 the tile, tile output bytes PUBLISHED (chain 2); descriptors into a ring of kFDesc slots
//   prefix warp 2 (block 0) the same look-back over the tiles' output bytes: where each tile's slice of the output starts
//   poll warp     waits for a described tile's output offset, so that stage C never does
//   stage C       (8 warps) assembles the tile's output slice shared -> shared ("_cell_umi" literal, word-wise copies of the
//                 short pieces, 16-byte chunks of read + quality funnel-shifted into place) and sends it off with a bulk
//                 store (two buffers)
// Two dependent look-back chains (record numbers, then output offsets) cost latency, not bandwidth: the windows and
// descriptor slots in flight cover it.  Nothing is guessed: a tile is only processed with its exact record numbers.
// Records outside the fast path's preconditions (control bytes in the header, more than two trailing blanks, records
// that outgrow the window, tiles with more than kFMaxRec records) make their tile take an exact path from global memory.
#pragma once
#include "ssd_emit_tiled.cuh"

namespace ssd {

#ifndef SSD_FUSED_SLOTS
#define SSD_FUSED_SLOTS 7
#endif
constexpr int kFTile = 16384;
constexpr int kFHalo = 2048;
constexpr int kFWin = kFTile + kFHalo;
constexpr int kFWinWords = kFWin / 32, kFTileWords = kFTile / 32, kFHaloWords = kFHalo / 32;
constexpr int kFSlots = SSD_FUSED_SLOTS;           // tile windows per block: being filled / in A / in B (3) / described, waiting / in C
constexpr int kFMaxRecTile = kFTile / 32;          // records starting in one tile (>= 32 bytes each on average)
constexpr int kFMaxLines = 4 * kFMaxRecTile + 72;  // line-start list entries (tile + halo)
constexpr int kFMaxRec = 128;                      // records of a tile that stage B describes for stage C
constexpr int kFOut = 20480;                       // staged output bytes per assembly batch (< 32768: offsets share words with flags)
constexpr int kFDesc = 4;                          // descriptor slots between stage B and stage C
constexpr int kFAWarps = 8, kFBWarps = 3, kFBGroups = 3, kFCWarps = 8;
constexpr int kFAThreads = kFAWarps * 32, kFBThreads = kFBWarps * 32, kFCThreads = kFCWarps * 32;
constexpr int kFWarpB0 = kFAWarps, kFWarpC0 = kFWarpB0 + kFBGroups * kFBWarps, kFWarpCopy = kFWarpC0 + kFCWarps;
constexpr int kFWarpPrefix1 = kFWarpCopy + 1, kFWarpPrefix2 = kFWarpCopy + 2, kFWarpPoll = kFWarpCopy + 3;
constexpr int kFThreads = (kFWarpPoll + 1) * 32;
constexpr uint32_t kFGeneric = 1u, kFSequential = 2u;  // tile flags: exact path per record / one warp walks the tile
static_assert(kFTileWords % kFAThreads == 0 && kFHaloWords <= 64, "tile shape");
static_assert(kFWin + 48 < kFOut, "one record of the fast path always fits an assembly batch");
static_assert(kFThreads <= 1024, "block size");

struct FusedArgs {
    const uint8_t* buf;  // read 1
    size_t len;
    unsigned long long* st1;  // per tile: newline count, then inclusive prefix (flag << 62 | value)
    unsigned long long* st2;  // per tile: output bytes, then inclusive prefix
    uint32_t n_tiles;
    DevCounters* cnt;
    unsigned long long* err_off;
    const uint64_t* ann;               // read 2's annotation words
    const unsigned long long* r2_tok;  // read 2's id fingerprints (matched records)
    uint64_t n_rec_r2;
    const uint8_t* pass;     // NULL = every matched record (MergedCells_1.fastq)
    const uint32_t* remap;   // NULL = no collapse
    const uint8_t* r2buf;    // irregular UMIs are copied from read 2
    const unsigned long long* wl64;  // whitelist entries as 8-byte words
    uint8_t* out;
    uint64_t cap;
};

// What stage B leaves for stage C about one tile.
struct FDesc {
    unsigned long long ann[kFMaxRec];  // annotation of the record, 0 = not written
    uint4 info[kFMaxRec];              // x = start in the window | cut << 16, y = kept | tail << 16, z = seq_rel | seq_len << 16, w = qual_rel | qual_len << 16
    uint32_t rel[kFMaxRec + 4];        // exclusive scan of the output lengths (n_rec + 1 entries)
    unsigned long long r_first, out_hi;  // first record number; inclusive prefix of the output bytes (set by the poll warp)
    uint32_t tile, n_rec, total, n_kept, flags, first_s0, cps, cps_magic;
};

struct __align__(128) FusedSmem {
    uint8_t win[kFSlots][kFWin + 128];  // tile windows (bulk-copy destinations)
    uint8_t out[2][kFOut + 32];         // assembled output slices (bulk-store sources)
    FDesc desc[kFDesc];
    uint2 c_seg[3][kFMaxRec];           // stage C, per record of the batch: {dst | src << 16, length} of header', read (+ "+" + quality when adjacent), quality
    uint2 c_chk[2][kFMaxRec];           // whole 16-byte chunks of the two long segments
    uint32_t c_lit[kFMaxRec];           // where the literal goes | two long segments << 31
    uint16_t lstart[kFBGroups][kFMaxLines];  // line-start lists between stage A and stage B
    alignas(16) uint32_t mask[kFWinWords];
    unsigned long long wl64[kMaxWl];
    unsigned long long full[kFSlots], empty[kFSlots];              // mbarriers: bytes landed / stage C is done with the window
    unsigned long long list_full[kFBGroups], list_empty[kFBGroups];  // stage A -> stage B
    unsigned long long desc_full[kFDesc], desc_ready[kFDesc], desc_empty[kFDesc];  // stage B -> poll warp -> stage C -> stage B
    unsigned long long line_base[kFBGroups];
    uint32_t tile[kFSlots];
    uint32_t total[kFBGroups], n_listed[kFBGroups], list_tile[kFBGroups];
    uint32_t warp_tile[kFAWarps], halo_sum[2];
    uint32_t b_warp[kFBGroups][kFBWarps], b_flags[kFBGroups], b_kept[kFBGroups];
    uint32_t c_flags[2];  // per assembly batch (alternating): 1 = some record has two long segments
};
static_assert(sizeof(FusedSmem) <= 232448, "the block's shared memory must fit the 227 KiB an sm_100 block may ask for");

// In-order inclusive prefix over published per-tile values (decoupled look-back done by ONE warp for everybody): every
// lane takes 8 consecutive entries, 256 tiles per round trip to L2.  Entries: flag << 62 | value, flag 1 = published.
__device__ __forceinline__ void prefix_chain(volatile unsigned long long* st, uint32_t n_tiles, int lane)
{
    unsigned long long running = 0;
    uint32_t next = 0, ns = 20;
    while (next < n_tiles) {
        unsigned long long v[8];
        const uint32_t first = next + 8u * (uint32_t)lane;
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = first + k < n_tiles ? st[first + k] : 0ull;
        uint32_t mine = 0;  // leading published entries of this lane
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
}

__device__ __forceinline__ unsigned long long poll_prefix(volatile unsigned long long* st, uint32_t tile)
{
    unsigned long long v;
    uint32_t ns = 20;
    while (((v = st[tile]) >> 62) != 2ull) {
        __nanosleep(ns);
        if (ns < 200u) ns += ns;
    }
    return v & kStatusValue;
}

// A read-1 record measured exactly and bytewise (any accessor): what the fast path cannot take.
struct R1Measure {
    bool complete, noat;
    uint32_t kept, seq_len, qual_len;
    unsigned long long fp;
};
template <class Acc>
__device__ __noinline__ void r1_measure_generic(const Acc& a, size_t len, size_t start, R1Measure& m)
{
    Lines L;
    m.complete = m.noat = false;
    m.kept = m.seq_len = m.qual_len = 0;
    m.fp = 0;
    if (!find_lines(a, start, len, true, L) || !L.complete) return;
    m.complete = true;
    m.noat = a.byte(L.s[0]) != '@';
    // name.split("/")[0].replace(" ", "").strip()  (V2:135-138)
    uint32_t kept = 0, kept_at_last_nonspace = 0;
    for (size_t p = L.s[0]; p < L.e[0]; p++) {
        const uint32_t c = a.byte(p);
        if (c == '/') break;
        if (c != ' ') kept++;
        if (!py_space(c)) kept_at_last_nonspace = kept;
    }
    m.kept = kept_at_last_nonspace;
    m.seq_len = (uint32_t)(rstrip_end(a, L.s[1], L.e[1]) - L.s[1]);
    m.qual_len = (uint32_t)(rstrip_end(a, L.s[3], L.e[3]) - L.s[3]);
    size_t tok_e = L.s[0];
    while (tok_e < L.e[0] && !py_space(a.byte(tok_e))) tok_e++;
    m.fp = tok_hash_generic(a, L.s[0], tok_e);
}


#ifdef SSD_FUSED_PROFILE
// debug builds only: clock totals per stage and wait reason, summed over the blocks (one thread per stage group counts)
//   0 A wait list slot  1 A wait data  2 A work   3 B wait list  4 B wait desc slot  5 B poll chain 1  6 B work
//   7 C wait ready  8 poll warp: chain 2  9 C work  10 tiles  11 copy wait window  12 C describe  13 C assemble (thread 0)
//   14 C fence + bulk wait + sync  15 C store
__device__ unsigned long long g_fused_prof[16];
#define FP_T(var) const long long var = clock64()
#define FP_ADD(slot, a, b) fprof[slot] += (unsigned long long)((b) - (a))
#define FP_DECL unsigned long long fprof[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}
#define FP_FLUSH(cond)                                                     \
    if (cond)                                                              \
        for (int k_ = 0; k_ < 16; k_++)                                    \
            if (fprof[k_]) atomicAdd(&g_fused_prof[k_], fprof[k_])
#else
#define FP_T(var)
#define FP_ADD(slot, a, b)
#define FP_DECL
#define FP_FLUSH(cond)
#endif

__device__ __forceinline__ void sync_fa() { asm volatile("bar.sync 1, %0;" ::"n"(kFAThreads) : "memory"); }
__device__ __forceinline__ void sync_fb(int g) { asm volatile("bar.sync %0, %1;" ::"r"(2 + g), "n"(kFBThreads) : "memory"); }
__device__ __forceinline__ void sync_fc() { asm volatile("bar.sync 5, %0;" ::"n"(kFCThreads) : "memory"); }

// n bytes shared -> shared at any alignment: dst-aligned 32-bit words (two source words funnel-shifted), at most three
// bytes on either side bytewise.  Reads up to 3 bytes past src + n (inside the window's padding).
__device__ __forceinline__ void copy_words(uint8_t* s_out, uint32_t dst, const uint8_t* s_in, uint32_t src, uint32_t n)
{
    uint32_t h = (4u - (dst & 3u)) & 3u;
    h = h < n ? h : n;
    for (uint32_t i = 0; i < h; i++) s_out[dst + i] = s_in[src + i];
    dst += h;
    src += h;
    n -= h;
    const uint32_t nw = n >> 2, sh = (src & 3u) * 8u;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(s_in + (src & ~3u));
    uint32_t* o = reinterpret_cast<uint32_t*>(s_out + dst);
    uint32_t prev = nw ? w[0] : 0u;
    for (uint32_t i = 0; i < nw; i++) {
        const uint32_t nxt = w[i + 1];
        o[i] = __funnelshift_r(prev, nxt, sh);
        prev = nxt;
    }
    for (uint32_t i = nw << 2; i < n; i++) s_out[dst + i] = s_in[src + i];
}

// the bytes of a long segment that do not fill a 16-byte output chunk (up to 15 in front, up to 15 behind), word-wise
__device__ __forceinline__ void edges_words(const uint8_t* s_in, uint8_t* s_out, uint2 sg)
{
    const uint32_t dst = sg.x & 0xFFFFu, src = sg.x >> 16, len = sg.y;
    uint32_t nh = (16u - (dst & 15u)) & 15u;
    nh = nh < len ? nh : len;
    copy_words(s_out, dst, s_in, src, nh);
    const uint32_t e = dst + len;
    uint32_t t0 = e & ~15u;
    t0 = t0 > dst + nh ? t0 : dst + nh;
    copy_words(s_out, t0, s_in, src + (t0 - dst), e - t0);
}

__global__ void __launch_bounds__(kFThreads, 1) k_fused_r1(const __grid_constant__ FusedArgs args, const __grid_constant__ DevCfg cfg)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    FusedSmem& S = *reinterpret_cast<FusedSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < kFSlots; s++) {
            mbar_init(reinterpret_cast<uint64_t*>(&S.full[s]), 1);
            mbar_init(reinterpret_cast<uint64_t*>(&S.empty[s]), 1);
        }
        for (int s = 0; s < kFBGroups; s++) {
            mbar_init(reinterpret_cast<uint64_t*>(&S.list_full[s]), 1);
            mbar_init(reinterpret_cast<uint64_t*>(&S.list_empty[s]), 1);
        }
        for (int s = 0; s < kFDesc; s++) {
            mbar_init(reinterpret_cast<uint64_t*>(&S.desc_full[s]), 1);
            mbar_init(reinterpret_cast<uint64_t*>(&S.desc_ready[s]), 1);
            mbar_init(reinterpret_cast<uint64_t*>(&S.desc_empty[s]), 1);
        }
    }
    for (int i = tid; i < kMaxWl; i += kFThreads) S.wl64[i] = args.wl64[i];
    __syncthreads();
    volatile unsigned long long* st1 = args.st1;
    volatile unsigned long long* st2 = args.st2;

    if (warp == kFWarpPrefix1) {
        if (blockIdx.x == 0) prefix_chain(st1, args.n_tiles, lane);
        return;
    }
    if (warp == kFWarpPrefix2) {
        if (blockIdx.x == 0) prefix_chain(st2, args.n_tiles, lane);
        return;
    }
    if (warp == kFWarpPoll) {
        // ================================ poll warp: where a described tile's slice of the output ends ================================
        if (lane != 0) return;
        FP_DECL;
        for (uint32_t it = 0;; it++) {
            const uint32_t ds = it % kFDesc;
            FDesc& D = S.desc[ds];
            mbar_wait(reinterpret_cast<uint64_t*>(&S.desc_full[ds]), (it / kFDesc) & 1u);
            const uint32_t tile = D.tile;
            if (tile != kNoTile) {
                FP_T(p0);
                const unsigned long long incl = poll_prefix(st2, tile);  // inclusive prefix of the output bytes, written by prefix warp 2
                FP_T(p1);
                FP_ADD(8, p0, p1);
                D.out_hi = incl;
                if (tile == args.n_tiles - 1) args.cnt->emit_bytes = incl;
            }
            mbar_arrive(&S.desc_ready[ds]);
            if (tile == kNoTile) break;
        }
        FP_FLUSH(true);
        return;
    }
    if (warp == kFWarpCopy) {
        // ================================ copy warp ================================
        FP_DECL;
        for (uint32_t it = 0;; it++) {
            const uint32_t ws = it % kFSlots;
            FP_T(p0);
            if (it >= (uint32_t)kFSlots) mbar_wait(reinterpret_cast<uint64_t*>(&S.empty[ws]), ((it / kFSlots) - 1u) & 1u);
            FP_T(p1);
            FP_ADD(11, p0, p1);
            uint32_t t = 0;
            if (lane == 0) t = atomicAdd(&args.cnt->ticket_fused, 1u);  // tiles are claimed in order: whatever a tile waits for is running or done
            t = __shfl_sync(0xFFFFFFFFu, t, 0);
            if (t >= args.n_tiles) {
                if (lane == 0) {
                    S.tile[ws] = kNoTile;
                    mbar_arrive(&S.full[ws]);
                }
                break;
            }
            const size_t tile_lo = (size_t)t * kFTile;
            const uint32_t wb = (uint32_t)(args.len - tile_lo < (size_t)kFWin ? args.len - tile_lo : (size_t)kFWin);
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
        FP_FLUSH(lane == 0);
        return;
    }

    if (warp < kFAWarps) {
        // ================================ stage A: newline bitmap, count, line-start list ================================
        FP_DECL;
        for (uint32_t it = 0;; it++) {
            const uint32_t ws = it % kFSlots, ls = it % kFBGroups;
            FP_T(p0);
            if (it >= (uint32_t)kFBGroups) mbar_wait(reinterpret_cast<uint64_t*>(&S.list_empty[ls]), ((it / kFBGroups) - 1u) & 1u);
            FP_T(p1);
            mbar_wait(reinterpret_cast<uint64_t*>(&S.full[ws]), (it / kFSlots) & 1u);
            FP_T(p2);
            FP_ADD(0, p0, p1);
            FP_ADD(1, p1, p2);
            const uint32_t tile = S.tile[ws];
            if (tile == kNoTile) {
                // the end reaches every stage-B group: this iteration's and the following ones'
                for (uint32_t j = 0; j < (uint32_t)kFBGroups; j++) {
                    const uint32_t it2 = it + j, ls2 = it2 % kFBGroups;
                    if (j != 0u && it2 >= (uint32_t)kFBGroups) mbar_wait(reinterpret_cast<uint64_t*>(&S.list_empty[ls2]), ((it2 / kFBGroups) - 1u) & 1u);
                    if (tid == 0) {
                        S.list_tile[ls2] = kNoTile;
                        mbar_arrive(&S.list_full[ls2]);
                    }
                }
                break;
            }
            const size_t tile_lo = (size_t)tile * kFTile;
            const uint32_t win_bytes = (uint32_t)(args.len - tile_lo < (size_t)kFWin ? args.len - tile_lo : (size_t)kFWin);
            const uint8_t* s_win = S.win[ws];
            uint32_t* s_mask = S.mask;
            uint16_t* s_lstart = S.lstart[ls];
            constexpr int WPT = kFTileWords / kFAThreads;  // mask words (32 bytes each) per thread
            uint32_t mw[WPT];
            uint32_t cnt = 0;
            build_nl_mask<kFAThreads, (kFWinWords * 2 + kFAThreads - 1) / kFAThreads>(s_win, win_bytes, reinterpret_cast<uint16_t*>(s_mask), kFWinWords * 2);
            sync_fa();
#pragma unroll
            for (int w = 0; w < WPT; w++) {
                mw[w] = s_mask[WPT * tid + w];
                cnt += (uint32_t)__popc(mw[w]);
            }
            const uint32_t hm = tid < kFHaloWords ? s_mask[kFTileWords + tid] : 0u;
            // both prefix sums in one: a warp's tile count in the low half, its halo count in the high half
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
            sync_fa();
            uint32_t warp_base = 0, total = 0;
#pragma unroll
            for (int w = 0; w < kFAWarps; w++) {
                const uint32_t v = S.warp_tile[w];
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
                st1[tile] = kFlagAgg | total;  // published: prefix warp 1 takes it from here
                s_lstart[0] = 0;
            }
            // line-start list: s_lstart[1 + k] = byte after the k-th newline of the window (tile first, then halo)
            {
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
                            if (idx < (uint32_t)kFMaxLines) s_lstart[idx] = (uint16_t)(pos0 + b);
                            idx++;
                        }
                    } else {
                        if (p >= 1u && idx < (uint32_t)kFMaxLines) s_lstart[idx] = (uint16_t)(pos0 + (uint32_t)__ffs((int)m) - 1u);
                        if (p == 2u && idx + 1u < (uint32_t)kFMaxLines) s_lstart[idx + 1u] = (uint16_t)(pos0 + 31u - (uint32_t)__clz((int)m));
                        idx += p;
                    }
                }
                uint32_t hrank = hexcl, mm = hm;
                while (mm) {
                    const uint32_t b = (uint32_t)__ffs((int)mm) - 1u;
                    mm &= mm - 1u;
                    if (hrank + 1u < (uint32_t)kFMaxLines) s_lstart[hrank + 1u] = (uint16_t)((kFTileWords + tid) * 32 + b + 1);
                    hrank++;
                }
            }
            sync_fa();  // list complete; warp_tile / halo_sum / mask may be reused
            if (tid == 0) mbar_arrive(&S.list_full[ls]);  // hand the tile to stage B
            FP_T(p3);
            FP_ADD(2, p2, p3);
        }
        FP_FLUSH(tid == 0);
        return;
    }

    if (warp < kFWarpC0) {
        // ================================ stage B: the records that start in the tile ================================
        const int g = (warp - kFWarpB0) / kFBWarps;                         // group: takes the block's tiles it = g, g + 3, ...
        const int tb = tid - (kFWarpB0 + g * kFBWarps) * 32, wb = tb >> 5;  // thread / warp inside the group
        GlobAcc ga{args.buf, args.len};
        FP_DECL;
        for (uint32_t it = (uint32_t)g;; it += (uint32_t)kFBGroups) {
            const uint32_t ws = it % kFSlots, ls = (uint32_t)g, ds = it % kFDesc;
            FDesc& D = S.desc[ds];
            FP_T(p0);
            mbar_wait(reinterpret_cast<uint64_t*>(&S.list_full[ls]), (it / kFBGroups) & 1u);
            FP_T(p1);
            if (it >= (uint32_t)kFDesc) mbar_wait(reinterpret_cast<uint64_t*>(&S.desc_empty[ds]), ((it / kFDesc) - 1u) & 1u);  // stage C is done with the slot
            FP_T(p2);
            FP_ADD(3, p0, p1);
            FP_ADD(4, p1, p2);
            const uint32_t tile = S.list_tile[ls];
            if (tile == kNoTile) {
                if (tb == 0) {
                    D.tile = kNoTile;
                    mbar_arrive(&S.desc_full[ds]);
                }
                break;
            }
            const size_t tile_lo = (size_t)tile * kFTile;
            const uint32_t win_bytes = (uint32_t)(args.len - tile_lo < (size_t)kFWin ? args.len - tile_lo : (size_t)kFWin);
            uint8_t* s_win = S.win[ws];
            const uint16_t* s_lstart = S.lstart[ls];
            const uint32_t total = S.total[ls], n_listed = S.n_listed[ls];
            const uint32_t* words = reinterpret_cast<const uint32_t*>(s_win);
            if (tb == 0) {
                // the inclusive newline prefix of this tile, written by prefix warp 1
                const unsigned long long incl = poll_prefix(st1, tile);
                FP_T(p3);
                FP_ADD(5, p2, p3);
                S.line_base[g] = incl - total;
                S.b_flags[g] = 0;
                S.b_kept[g] = 0;
                if (tile == args.n_tiles - 1) {
                    const unsigned long long n_lines = incl + ((args.len && s_win[win_bytes - 1] != '\n') ? 1ull : 0ull);
                    args.cnt->n_lines[0] = n_lines;
                    args.cnt->n_records[0] = n_lines >> 2;
                }
            }
            sync_fb(g);
            const unsigned long long line_base = S.line_base[g];
            // records that start in this tile: the line after a newline whose next-line index is 0 mod 4
            const unsigned long long r_first = tile == 0 ? 0ull : (line_base + 4) >> 2;
            const unsigned long long r_end = ((line_base + total) >> 2) + 1;  // exclusive
            uint32_t n_rec = r_end > r_first ? (uint32_t)(r_end - r_first) : 0u;
            if (tile == 0 && args.len == 0) n_rec = 0;
            if (n_rec > (uint32_t)kFMaxRecTile || n_listed + 1u > (uint32_t)kFMaxLines) {
                if (tb == 0) atomicOr(&args.cnt->err_flags, kErrDense);
                n_rec = 0;
            }
            const uint32_t j_first = (uint32_t)(4ull * r_first - line_base);  // list index of the first record's start

            uint32_t run = 0;  // output bytes of the records before this round
            for (uint32_t base = 0; base < n_rec || base == 0; base += kFBThreads) {
                const uint32_t i = base + (uint32_t)tb;
                const bool have = i < n_rec;
                const uint64_t r = r_first + i;
                // the record's annotation and read 2's id fingerprint: requested first, used after the header walk
                uint64_t an = 0;
                unsigned long long tok2 = 0;
                if (have && r < args.n_rec_r2) {
                    an = __ldg(args.ann + r);
                    tok2 = __ldg(args.r2_tok + r);
                }
                const uint32_t j0 = j_first + 4u * i;
                uint32_t s0 = 0, s1 = 1, s2 = 1, s3 = 1, s4 = 1;
                bool fast = false;
                if (have) {
                    s0 = s_lstart[j0];
                    if (j0 + 4u <= n_listed) {
                        s1 = s_lstart[j0 + 1];
                        s2 = s_lstart[j0 + 2];
                        s3 = s_lstart[j0 + 3];
                        s4 = s_lstart[j0 + 4];
                        fast = true;
                    }
                }
                // header bytes [s0, s0 + hdr_len), one trailing '\r' dropped; token = bytes before the first byte <= 0x20
                uint32_t hdr_len = fast ? s1 - 1u - s0 : 0u;
                if (fast && hdr_len && s_win[s0 + hdr_len - 1u] == '\r') hdr_len--;
                const uint32_t a = s0 & 3u, nwords = fast ? (hdr_len + 3u) >> 2 : 0u;
                const uint32_t trip = __reduce_max_sync(0xFFFFFFFFu, nwords);
                // one pass over the header, four bytes at a time from its first byte: id fingerprint and the pieces of
                // name.split("/")[0].replace(" ", "").strip() (V2:135-138)
                uint32_t cut = hdr_len, spaces = 0, bad = 0, tok_len = hdr_len, first_sp = 0xFFFFFFFFu;
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
                            if (!found_slash) {
                                const uint32_t zs = zero_flags(x ^ 0x2F2F2F2Fu) & V;
                                uint32_t zp = zero_flags(x ^ 0x20202020u) & V;
                                if (zs) {
                                    const uint32_t bit = (uint32_t)__ffs((int)zs) - 1u;
                                    cut = 4u * w + (bit >> 3);
                                    zp &= (1u << bit) - 1u;
                                    found_slash = true;
                                }
                                if (zp && first_sp == 0xFFFFFFFFu) first_sp = 4u * w + (((uint32_t)__ffs((int)zp) - 1u) >> 3);
                                spaces += (uint32_t)__popc(zp);
                            }
                        }
                    }
                }
                unsigned long long fp = tok_hash_finish(h, tok_len);
                uint32_t seq_e = s2 - 1u, qual_e = s4 - 1u;
                bool ok = fast && bad == 0u;
                if (ok) ok = rstrip2(s_win, s1, seq_e) && rstrip2(s_win, s3, qual_e);
                bool complete = ok, noat = false;
                uint32_t kept = cut - spaces, seq_len = seq_e - s1, qual_len = qual_e - s3, tail = 0;
                if (ok) {
                    noat = s_win[s0] != '@';
                    tail = (seq_e == s2 - 1u && s3 - s2 == 2u && s_win[s2] == '+' && qual_e == s4 - 1u) ? 1u : 0u;
                }
                __syncwarp();
                if (have && !ok) {
                    // outside the fast path: measured bytewise from global memory; the whole tile is then written the exact way
                    R1Measure m;
                    r1_measure_generic(ga, args.len, tile_lo + s0, m);
                    complete = m.complete;
                    noat = m.noat;
                    kept = m.kept;
                    seq_len = m.seq_len;
                    qual_len = m.qual_len;
                    fp = m.fp;
                    atomicOr(&S.b_flags[g], kFGeneric);
                }
                __syncwarp();
                uint32_t out_len = 0;
                bool keep = false;
                if (have && complete) {
                    if (noat) raise(args.cnt, args.err_off, 0, kErrNoAt, tile_lo + s0);
                    if (ann_state(an) != 0u) {
                        // V2:349-350: the matched read-2 record must meet a read-1 record of the same id (fingerprint compare)
                        if (tok2 != fp) raise_key(args.cnt, args.err_off, r);
                        keep = args.pass == nullptr || args.pass[ann_cell(an)] != 0;
                    }
                    if (keep) out_len = kept + 27u + ann_umi_len(an, cfg.umi_len) + seq_len + qual_len + 4u;
                }
                if (keep && ok && spaces) {
                    // name.replace(" ", ""): squeezed in place, so that the writer copies header' like any other piece
                    uint8_t* hw = s_win + s0;
                    uint32_t o = first_sp;
                    for (uint32_t k = first_sp + 1u; k < cut; k++) {
                        const uint8_t ch = hw[k];
                        if (ch != ' ') hw[o++] = ch;
                    }
                }
                // exclusive scan of the output lengths over the group (and over the rounds)
                uint32_t incl = out_len;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                    if (lane >= d) incl += v;
                }
                const uint32_t nk = (uint32_t)__popc(__ballot_sync(0xFFFFFFFFu, keep));
                if (lane == 31) S.b_warp[g][wb] = incl;
                if (lane == 0 && nk) atomicAdd(&S.b_kept[g], nk);
                sync_fb(g);
                uint32_t before = 0, round_total = 0;
#pragma unroll
                for (int w = 0; w < kFBWarps; w++) {
                    const uint32_t v = S.b_warp[g][w];
                    if (w < wb) before += v;
                    round_total += v;
                }
                const uint32_t rel = run + before + incl - out_len;
                if (have && i < (uint32_t)kFMaxRec) {
                    D.ann[i] = keep ? an : 0ull;
                    D.info[i] = make_uint4(s0 | (cut << 16), kept | (tail << 16), (s1 - s0) | (seq_len << 16), (s3 - s0) | (qual_len << 16));
                    D.rel[i] = rel;
                }
                run += round_total;
                sync_fb(g);  // b_warp may be reused; the descriptors and the squeezed headers are in place
            }
            if (tb == 0) {
                const uint32_t n_kept = S.b_kept[g];
                uint32_t flags = S.b_flags[g];
                if (n_rec > (uint32_t)kFMaxRec) flags |= kFGeneric | kFSequential;
                else D.rel[n_rec] = run;
                D.tile = tile;
                D.n_rec = n_rec;
                D.total = run;
                D.n_kept = n_kept;
                D.flags = flags;
                D.r_first = r_first;
                D.first_s0 = n_rec ? (tile == 0 ? 0u : (uint32_t)s_lstart[j_first]) : 0u;
                const uint32_t avg = n_kept ? run / n_kept : 0u;
                uint32_t cps = (avg > 64u ? avg - 64u : 0u) / 16u;
                cps = cps < 1u ? 1u : (cps > 64u ? 64u : cps);
                D.cps = cps;
                D.cps_magic = ((1u << 20) + cps - 1u) / cps;
                st2[tile] = kFlagAgg | (unsigned long long)run;  // published: prefix warp 2 takes it from here
                mbar_arrive(&S.desc_full[ds]);   // over to the poll warp and stage C (which also releases the window)
                mbar_arrive(&S.list_empty[ls]);  // the list slot may be refilled
                FP_T(p4);
                FP_ADD(6, p2, p4);
            }
        }
        FP_FLUSH(tb == 0);
        return;
    }

    // ================================ stage C: assemble the tile's slice of the output and send it off ================================
    {
        const int tc = tid - kFWarpC0 * 32, wc = tc >> 5;
        GlobAcc ga{args.buf, args.len};
        unsigned long long my_records = 0;
        uint32_t nbuf = 0;
        FP_DECL;
        for (uint32_t it = 0;; it++) {
            const uint32_t ds = it % kFDesc, ws = it % kFSlots;
            const FDesc& D = S.desc[ds];
            FP_T(p0);
            mbar_wait(reinterpret_cast<uint64_t*>(&S.desc_ready[ds]), (it / kFDesc) & 1u);
            FP_T(p1);
            FP_ADD(7, p0, p1);
            const uint32_t tile = D.tile;
            if (tile == kNoTile) break;
            const uint32_t total = D.total, n_rec = D.n_rec, flags = D.flags;
            const unsigned long long out_hi = D.out_hi, out_lo = out_hi - total;
            if (tc == 0) my_records += D.n_kept;
            const size_t tile_lo = (size_t)tile * kFTile;
            const uint8_t* s_win = S.win[ws];
            if (total != 0u && out_hi > args.cap) {
                if (tc == 0) atomicOr(&args.cnt->err_flags, kErrRange);  // nothing is written past the end of the buffer
            } else if (total != 0u && (flags & kFSequential)) {
                // more records than stage B describes: one warp walks the tile's records in order, from global memory
                if (wc == 0) {
                    size_t pos = tile_lo + D.first_s0;
                    unsigned long long off = out_lo;
                    for (uint32_t i = 0; i < n_rec; i++) {
                        Lines L;
                        find_lines(ga, pos, args.len, true, L);
                        if (!L.complete) break;
                        const uint64_t r = D.r_first + i;
                        const uint64_t an = r < args.n_rec_r2 ? args.ann[r] : 0ull;
                        if (ann_state(an) != 0u && (args.pass == nullptr || args.pass[ann_cell(an)] != 0))
                            off += emit_record_plain(ga, args.len, pos, an, args.remap, args.r2buf, cfg, args.out + off);
                        pos = L.e[3] + 1;
                    }
                    if (lane == 0 && off != out_hi) atomicOr(&args.cnt->err_flags, kErrEmitLen);
                }
            } else if (total != 0u && (flags & kFGeneric)) {
                // a record outside the fast path: every record of the tile one warp each, global memory to global memory
                for (uint32_t k = (uint32_t)wc; k < n_rec; k += kFCWarps) {
                    const unsigned long long an = D.ann[k];
                    if (an == 0ull) continue;
                    const uint32_t wrote = emit_record_plain(ga, args.len, tile_lo + (D.info[k].x & 0xFFFFu), an, args.remap, args.r2buf, cfg,
                                                             args.out + out_lo + D.rel[k]);
                    if (lane == 0 && wrote != D.rel[k + 1] - D.rel[k]) atomicOr(&args.cnt->err_flags, kErrEmitLen);
                }
            } else if (total != 0u) {
                const uint32_t cps = D.cps, cps_magic = D.cps_magic;
                for (uint32_t k_lo = 0; k_lo < n_rec;) {
                    // the next batch: the records whose output fits one staging buffer (all of the tile's, as a rule)
                    const uint32_t base = D.rel[k_lo];
                    const unsigned long long q_lo = out_lo + base;
                    const uint32_t skew = (uint32_t)(((uintptr_t)args.out + q_lo) & 15u);
                    uint32_t k_hi = n_rec;
                    if (D.rel[k_hi] - base + skew > (uint32_t)kFOut) {
                        uint32_t lo = k_lo + 1u, hi = k_hi;  // rel[lo] fits (one record always does), rel[hi] does not
                        while (hi - lo > 1u) {
                            const uint32_t mid = (lo + hi) >> 1;
                            if (D.rel[mid] - base + skew <= (uint32_t)kFOut) lo = mid;
                            else hi = mid;
                        }
                        k_hi = lo;
                    }
                    const uint32_t nq = k_hi - k_lo, nout = D.rel[k_hi] - base;
                    if (nout != 0u) {
                        FP_T(c0);
                        uint8_t* s_out = S.out[nbuf & 1u];
                        uint32_t* c_flags = &S.c_flags[nbuf & 1u];  // the slot the previous batch reads is the other one
                        if (tc == 0) *c_flags = 0;
                        // ---- one thread per record: where its pieces go ----
                        if ((uint32_t)tc < nq) {
                            const uint32_t k = k_lo + (uint32_t)tc;
                            const unsigned long long an = D.ann[k];
                            uint2 sh = make_uint2(0u, 0u), sa = make_uint2(0u, 0u), sb = make_uint2(0u, 0u);
                            uint32_t lit_at = 0;
                            if (an != 0ull) {
                                const uint4 nf = D.info[k];
                                const uint32_t src0 = nf.x & 0xFFFFu, kept = nf.y & 0xFFFFu, seq_rel = nf.z & 0xFFFFu, seq_len = nf.z >> 16,
                                               qual_rel = nf.w & 0xFFFFu, qual_len = nf.w >> 16;
                                const uint32_t lit = 27u + ann_umi_len(an, cfg.umi_len);
                                const uint32_t d = skew + (D.rel[k] - base), dA = d + kept + lit;
                                sh = make_uint2(d | (src0 << 16), kept);  // header', squeezed in place by stage B
                                lit_at = d + kept;
                                if (nf.y >> 16) {  // the source already reads  read "\n+\n" quality "\n"
                                    sa = make_uint2(dA | ((src0 + seq_rel) << 16), qual_rel + qual_len + 1u - seq_rel);
                                } else {
                                    sa = make_uint2(dA | ((src0 + seq_rel) << 16), seq_len);
                                    sb = make_uint2((dA + seq_len + 3u) | ((src0 + qual_rel) << 16), qual_len);
                                    lit_at |= 0x80000000u;
                                }
                            }
                            S.c_seg[0][tc] = sh;
                            S.c_seg[1][tc] = sa;
                            S.c_seg[2][tc] = sb;
                            S.c_chk[0][tc] = tile_chunk_desc(sa);
                            S.c_chk[1][tc] = tile_chunk_desc(sb);
                            S.c_lit[tc] = lit_at;
                        }
                        sync_fc();
                        FP_T(c1);
                        FP_ADD(12, c0, c1);
                        if (tc < 128) {
                            // ---- '_' + barcode1 + barcode2 + barcode3 + '_' + umi + '\n' ----
                            const uint32_t t = (uint32_t)tc;
                            const unsigned long long an = t < nq ? D.ann[k_lo + t] : 0ull;
                            if (an != 0ull) {
                                const uint32_t la = S.c_lit[t];
                                if (la & 0x80000000u) atomicOr(c_flags, 1u);
                                uint8_t* q = s_out + (la & 0x7FFFu);
                                uint32_t cell = ann_cell(an);
                                if (args.remap) cell = args.remap[cell];
                                uint32_t i1, i2, i3;
                                cell_indices(cfg, cell, i1, i2, i3);
                                const unsigned long long b1 = S.wl64[i1], b2 = S.wl64[i2], b3 = S.wl64[i3];
                                const uint32_t b1l = (uint32_t)b1, b1h = (uint32_t)(b1 >> 32), b2l = (uint32_t)b2, b2h = (uint32_t)(b2 >> 32), b3l = (uint32_t)b3,
                                               b3h = (uint32_t)(b3 >> 32);
                                q[0] = '_';
#pragma unroll
                                for (int j = 0; j < 4; j++) {
                                    q[1 + j] = (uint8_t)(b1l >> (8 * j));
                                    q[5 + j] = (uint8_t)(b1h >> (8 * j));
                                    q[9 + j] = (uint8_t)(b2l >> (8 * j));
                                    q[13 + j] = (uint8_t)(b2h >> (8 * j));
                                    q[17 + j] = (uint8_t)(b3l >> (8 * j));
                                    q[21 + j] = (uint8_t)(b3h >> (8 * j));
                                }
                                q[25] = '_';
                                uint32_t ulen;
                                if (ann_state(an) == 1u) {
                                    ulen = (uint32_t)cfg.umi_len;
                                    const uint32_t code = (uint32_t)an;  // 2 bits per base, at most 10 bases
#pragma unroll
                                    for (int j = 0; j < 10; j++)
                                        if ((uint32_t)j < ulen) q[26 + j] = (uint8_t)__byte_perm(0x47544341u, 0u, (code >> (2 * j)) & 3u);
                                } else {
                                    const uint64_t payload = ann_payload(an);
                                    ulen = (uint32_t)(payload & 31u);
                                    const uint8_t* u = args.r2buf + (size_t)(payload >> 5);
                                    for (uint32_t j = 0; j < ulen; j++) q[26u + j] = u[j];
                                }
                                q[26u + ulen] = '\n';
                            }
                        } else {
                            // ---- header' and the ragged ends of the long segments (word-wise), the separators of records whose lines are not adjacent ----
                            const uint32_t t = (uint32_t)tc - 128u;
                            if (t < nq && D.ann[k_lo + t] != 0ull) {
                                const uint2 sh = S.c_seg[0][t], sa = S.c_seg[1][t];
                                copy_words(s_out, sh.x & 0xFFFFu, s_win, sh.x >> 16, sh.y);
                                edges_words(s_win, s_out, sa);
                                if (S.c_lit[t] & 0x80000000u) {
                                    const uint2 sb = S.c_seg[2][t];
                                    edges_words(s_win, s_out, sb);
                                    uint8_t* q = s_out + (sa.x & 0xFFFFu) + sa.y;
                                    q[0] = '\n';
                                    q[1] = '+';
                                    q[2] = '\n';
                                    q[3u + sb.y] = '\n';
                                }
                            }
                        }
                        // ---- whole 16-byte chunks of the long segments: cps slots per record, slot i = (record i / cps, chunk i % cps) ----
                        {
                            const uint32_t n_items = nq * cps;
                            for (uint32_t i = (uint32_t)tc; i < n_items; i += kFCThreads) {
                                const uint32_t k = (i * cps_magic) >> 20, j = i - k * cps;
                                const uint2 ck = S.c_chk[0][k];
                                const uint32_t nfull = ck.y & 0xFFFFu;
                                if (j < nfull) tile_chunk(s_win, s_out, ck, j);
                                if (nfull > cps)
                                    for (uint32_t c = j + cps; c < nfull; c += cps) tile_chunk(s_win, s_out, ck, c);  // segments longer than the slots
                            }
                        }
                        FP_T(c2);
                        FP_ADD(13, c1, c2);
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        if (tc == 0) bulk_wait_read();  // the slice sent off from this buffer two batches ago has left it (and so has every other one)
                        sync_fc();
                        FP_T(c3);
                        FP_ADD(14, c2, c3);
                        if (*c_flags & 1u) {
                            // records whose read and quality lines are two segments (rare): the second segments' chunks
                            const uint32_t n_items = nq * cps;
                            for (uint32_t i = (uint32_t)tc; i < n_items; i += kFCThreads) {
                                const uint32_t k = (i * cps_magic) >> 20, j = i - k * cps;
                                const uint2 ck = S.c_chk[1][k];
                                for (uint32_t c = j; c < (ck.y & 0xFFFFu); c += cps) tile_chunk(s_win, s_out, ck, c);
                            }
                            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                            sync_fc();
                        }
                        // ---- the finished slice leaves with one bulk copy; shared and global addresses agree modulo 16 ----
                        uint8_t* gdst = args.out + q_lo;
                        const uint32_t head = skew ? (16u - skew < nout ? 16u - skew : nout) : 0u;
                        const uint32_t body = (nout - head) & ~15u;
                        if (tc == 0 && body) bulk_s2g(gdst + head, s_out + skew + head, body);
                        if ((uint32_t)tc < head) gdst[tc] = s_out[skew + tc];
                        if (tc >= 32 && tc < 48) {
                            const uint32_t i = head + body + (uint32_t)(tc - 32);
                            if (i < nout) gdst[i] = s_out[skew + i];
                        }
                        nbuf++;
                        FP_T(c4);
                        FP_ADD(15, c3, c4);
                    }
                    k_lo = k_hi;
                }
            }
            sync_fc();  // every thread is done with the window and the descriptors
            if (tc == 0) {
                mbar_arrive(&S.desc_empty[ds]);
                mbar_arrive(&S.empty[ws]);  // the window may be refilled
                FP_T(p3);
                FP_ADD(9, p1, p3);
                FP_ADD(10, 0, 1);
            }
        }
        FP_FLUSH(tc == 0);
        if (tc == 0) {
            bulk_wait_read();  // shared memory must outlive the last bulk store
            if (my_records) atomicAdd(&args.cnt->emit_records, my_records);
        }
    }
}

// Error path of the mate check (V2:349-350): the first matched read-2 record at or after record `from` (read 1 ended there).
__global__ void k_first_matched_from(const uint64_t* ann, uint64_t from, uint64_t n, unsigned long long* err_rec)
{
    for (uint64_t r = from + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (uint64_t)gridDim.x * blockDim.x)
        if (ann_state(ann[r]) != 0u) atomicMin(err_rec, (unsigned long long)r);
}

// Records of this batch that the post-filter output holds (ssd_build_filter reports it before anything is written).
__global__ void __launch_bounds__(kThreads) k_count_retained(const uint64_t* ann, uint64_t n, const uint8_t* pass, DevCounters* cnt)
{
    unsigned long long mine = 0;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t a = __ldg(ann + r);
        if (ann_state(a) != 0u && pass[ann_cell(a)] != 0) mine++;
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) mine += __shfl_xor_sync(0xFFFFFFFFu, mine, d);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&cnt->n_retained, mine);
}

}  // namespace ssd
