// ssd_emit_tiled.cuh — the writer for the common case (input order, whitelist entries 8 bytes long), sm_100a.
//
// A block owns up to kTR consecutive records.  Their read-1 bytes are one contiguous range of the input and their
// output is one contiguous range of the result, so both sides move with bulk copies (TMA): global -> shared for the
// input range, shared -> global for the finished output slice.  In between the records are assembled shared -> shared:
//   * one thread per record describes it (where its header, literal and read/quality bytes go),
//   * one thread per record squeezes the header (name.split("/")[0].replace(" ", "").strip(), V2:135-138) and copies
//     the few bytes of the long segments that do not fill a 16-byte output chunk,
//   * one thread per record writes "_" barcode1 barcode2 barcode3 "_" umi "\n" (V2:139-141),
//   * every thread copies whole 16-byte output chunks of the long segments: two aligned 16-byte shared loads,
//     funnel-shifted into one aligned 16-byte shared store.
// About 30 warp instructions per record (the eight-lanes-per-record writer in ssd_emit.cuh needs ~150).
// Blocks that do not fit the staging buffers (very long records) take the exact one-warp-per-record path.
#pragma once
#include "ssd_emit.cuh"

namespace ssd {

#ifndef SSD_EMIT_TR
#define SSD_EMIT_TR 64
#endif
#ifndef SSD_EMIT_MINBLOCKS
#define SSD_EMIT_MINBLOCKS 2
#endif
constexpr int kTR = SSD_EMIT_TR;          // records per tile, at most (32 or 64)
constexpr int kTIn = kTR * 384;           // staged read-1 bytes per tile
constexpr int kTOut = kTR * 416;          // staged output bytes per tile
constexpr int kTPad = 16;          // bytes in front of the staged input (16-byte loads may start before the first record)

struct __align__(128) TileSmem {
    uint8_t in[kTPad + kTIn + 48];
    uint8_t out[kTOut + 32];
    unsigned long long ann[kTR];   // annotation of the record, 0 = not written
    uint2 seg[2][kTR];             // long segments {dst | src << 16, length}: read (+ "+" + quality when adjacent), quality
    uint2 chk[2][kTR];             // their whole 16-byte chunks {first dst | aligned src << 16, count | word shift << 16 | bit shift << 24}
    uint2 head[kTR];               // {dst | src << 16, cut length | literal dst << 16}
    unsigned long long bar;        // mbarrier of the input copy
    unsigned long long g0, g1;     // file range of the block's records
    uint32_t fits, flags;
};

__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// [b0, b1): the part of the file range [g0, g1) between 16-byte aligned addresses; false when there is none
__device__ __forceinline__ bool tile_bulk_range(const uint8_t* buf, uint64_t g0, uint64_t g1, uint64_t& b0, uint64_t& b1)
{
    const uintptr_t a0 = (uintptr_t)(buf + g0), a1 = (uintptr_t)(buf + g1);
    const uintptr_t ab0 = (a0 + 15u) & ~(uintptr_t)15, ab1 = a1 & ~(uintptr_t)15;
    b0 = g0 + (uint64_t)(ab0 - a0);
    b1 = ab1 > ab0 ? g1 - (uint64_t)(a1 - ab1) : b0;
    return ab1 > ab0;
}

// the bytes of a long segment that do not fill a 16-byte output chunk: up to 15 in front, up to 15 behind
__device__ __forceinline__ void tile_edges(const uint8_t* s_in, uint8_t* s_out, uint2 sg)
{
    const uint32_t dst = sg.x & 0xFFFFu, src = sg.x >> 16, len = sg.y;
    uint32_t nh = (16u - (dst & 15u)) & 15u;
    nh = nh < len ? nh : len;
    for (uint32_t i = 0; i < nh; i++) s_out[dst + i] = s_in[src + i];
    const uint32_t e = dst + len;
    uint32_t t0 = e & ~15u;
    t0 = t0 > dst + nh ? t0 : dst + nh;
    for (uint32_t p = t0; p < e; p++) s_out[p] = s_in[src + (p - dst)];
}

// descriptor of the whole 16-byte output chunks of a long segment
__device__ __forceinline__ uint2 tile_chunk_desc(uint2 sg)
{
    const uint32_t dst = sg.x & 0xFFFFu, src = sg.x >> 16, len = sg.y;
    const uint32_t first = (dst + 15u) & ~15u, last = (dst + len) & ~15u;
    if (last <= first) return make_uint2(0u, 0u);
    const uint32_t sa0 = src + (first - dst);  // source of the first whole chunk
    return make_uint2(first | ((sa0 & ~15u) << 16), ((last - first) >> 4) | (((sa0 & 15u) >> 2) << 16) | (((sa0 & 3u) * 8u) << 24));
}

// whole chunk c of a long segment: two aligned 16-byte loads funnel-shifted into one aligned 16-byte store
__device__ __forceinline__ void tile_chunk(const uint8_t* s_in, uint8_t* s_out, uint2 ck, uint32_t c)
{
    const uint32_t bs = ck.y >> 24;
    const bool s2 = ck.y & 0x20000u, s1 = ck.y & 0x10000u;
    const uint4* in16 = reinterpret_cast<const uint4*>(s_in + (ck.x >> 16)) + c;
    const uint4 A = in16[0], B = in16[1];
    const uint32_t t0 = s2 ? A.z : A.x, t1 = s2 ? A.w : A.y, t2 = s2 ? B.x : A.z, t3 = s2 ? B.y : A.w, t4 = s2 ? B.z : B.x, t5 = s2 ? B.w : B.y;
    const uint32_t w0 = s1 ? t1 : t0, w1 = s1 ? t2 : t1, w2 = s1 ? t3 : t2, w3 = s1 ? t4 : t3, w4 = s1 ? t5 : t4;
    reinterpret_cast<uint4*>(s_out + (ck.x & 0xFFFFu))[c] =
        make_uint4(__funnelshift_r(w0, w1, bs), __funnelshift_r(w1, w2, bs), __funnelshift_r(w2, w3, bs), __funnelshift_r(w3, w4, bs));
}

// records [r0, r1) one warp each, global memory to global memory (exact, any record)
template <typename OffT>
__device__ __noinline__ void emit_block_generic(const EmitArgs<OffT>& args, const DevCfg& cfg, uint64_t r0, uint64_t r1, int n_warps = kThreads / 32)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    GlobAcc ga{args.r1buf, args.r1_len};
    for (uint64_t k = r0 + warp; k < r1; k += n_warps) {
        const uint64_t o0 = args.out_off[k], o1 = args.out_off[k + 1];
        if (o1 == o0) continue;
        const uint64_t r = args.order ? args.order[k] : k;
        const uint32_t wrote = emit_record_generic(ga, args, cfg, r, args.out + o0);
        if (lane == 0 && wrote != (uint32_t)(o1 - o0)) atomicOr(&args.cnt->err_flags, kErrEmitLen);
    }
}

template <typename OffT>
__global__ void __launch_bounds__(kThreads) k_emit_tiled(const __grid_constant__ EmitArgs<OffT> args, const __grid_constant__ DevCfg cfg)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    TileSmem& S = *reinterpret_cast<TileSmem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint64_t r0 = (uint64_t)blockIdx.x * args.rpb;
    if (r0 >= args.n_rec) return;
    const uint64_t r1 = r0 + args.rpb < args.n_rec ? r0 + args.rpb : args.n_rec;
    // the block's file range: requested together with the output range (one round trip before the bulk copy can start)
    uint64_t g0t = 0, g1t = 0;
    uint4 ml = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0) {
        g0t = (uint64_t)args.r1_start[r0];
        g1t = (uint64_t)args.r1_start[r1 - 1];
        ml = args.r1_meta[r1 - 1];
    }
    const uint64_t out_lo = args.out_off[r0], out_hi = args.out_off[r1];
    if (out_hi == out_lo) return;  // nothing of this block survives the filter
    const uint32_t nrec = (uint32_t)(r1 - r0);
    const uint32_t out_skew = (uint32_t)(((uintptr_t)args.out + out_lo) & 15u);

    // ---- the block's input range; its 16-byte aligned middle arrives by bulk copy ----
    if (tid == 0) {
        const uint64_t g0 = g0t;
        uint64_t g1 = g1t + (ml.z & 0xFFFFu) + (ml.z >> 16) + 1u;  // behind the last quality line
        g1 = g1 < (uint64_t)args.r1_len ? g1 : (uint64_t)args.r1_len;
        const uint32_t mis = (uint32_t)((uintptr_t)(args.r1buf + g0) & 15u);
        const bool fits = nrec <= (uint32_t)kTR && g1 > g0 && g1 - g0 + mis <= (uint64_t)kTIn && out_hi - out_lo + out_skew <= (uint64_t)kTOut;
        S.g0 = g0;
        S.g1 = g1;
        S.fits = fits ? 1u : 0u;
        S.flags = 0;
        mbar_init(reinterpret_cast<uint64_t*>(&S.bar), 1);
        if (fits) {
            uint64_t b0, b1;
            if (tile_bulk_range(args.r1buf, g0, g1, b0, b1)) {
                mbar_expect_tx(reinterpret_cast<uint64_t*>(&S.bar), (uint32_t)(b1 - b0));
                bulk_g2s(&S.in[kTPad + mis + (uint32_t)(b0 - g0)], args.r1buf + b0, (uint32_t)(b1 - b0), reinterpret_cast<uint64_t*>(&S.bar));
            } else {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&S.bar)) : "memory");
            }
        }
    }
    __syncthreads();
    if (!S.fits) {
        emit_block_generic(args, cfg, r0, r1);
        return;
    }
    const uint64_t g0 = S.g0, g1 = S.g1;
    const uint32_t mis = (uint32_t)((uintptr_t)(args.r1buf + g0) & 15u);
    const uint32_t in0 = (uint32_t)kTPad + mis;  // S.in[in0 + (x - g0)] holds file byte x
    if (tid < 32) {
        // the unaligned ends of the range (the whole range when it has no aligned middle)
        uint64_t b0, b1;
        if (tile_bulk_range(args.r1buf, g0, g1, b0, b1)) {
            if (tid < 16) {
                const uint64_t x = g0 + (uint32_t)tid;
                if (x < b0) S.in[in0 + (uint32_t)tid] = args.r1buf[x];
            } else {
                const uint64_t x = b1 + (uint32_t)(tid - 16);
                if (x < g1) S.in[in0 + (uint32_t)(x - g0)] = args.r1buf[x];
            }
        } else {
            const uint64_t x = g0 + (uint32_t)tid;
            if (x < g1) S.in[in0 + (uint32_t)tid] = args.r1buf[x];
        }
    }

    // ---- one thread per record: where its pieces go ----
    if ((uint32_t)tid < nrec) {
        const uint64_t r = r0 + (uint32_t)tid;
        const uint64_t o0 = args.out_off[r], o1 = args.out_off[r + 1];
        uint2 sa = make_uint2(0u, 0u), sb = make_uint2(0u, 0u), hd = make_uint2(0u, 0u);
        unsigned long long an = 0;
        if (o1 != o0) {
            const uint64_t start = (uint64_t)args.r1_start[r];
            const uint4 mt = args.r1_meta[r];
            an = args.ann[r];
            const uint32_t cut = mt.x & 0xFFFFu, kept = mt.x >> 16, seq_rel = mt.y & 0xFFFFu, seq_len = mt.y >> 16, qual_rel = mt.z & 0xFFFFu,
                           qual_len = mt.z >> 16;
            const uint32_t lit = 27u + ann_umi_len(an, cfg.umi_len);
            if ((mt.w & kMetaLong) || start < g0 || start + qual_rel + qual_len + ((mt.w & kMetaTail) ? 1u : 0u) > g1) {
                atomicOr(&S.flags, 1u);  // the whole block goes the exact way
                an = 0;
            } else if (kept + lit + seq_len + qual_len + 4u != (uint32_t)(o1 - o0)) {
                atomicOr(&args.cnt->err_flags, kErrEmitLen);
                an = 0;
            } else {
                const uint32_t d = out_skew + (uint32_t)(o0 - out_lo), src0 = in0 + (uint32_t)(start - g0), dA = d + kept + lit;
                hd = make_uint2(d | (src0 << 16), cut | ((d + kept) << 16));
                if (mt.w & kMetaTail) {  // the source already reads  read "\n+\n" quality "\n"
                    sa = make_uint2(dA | ((src0 + seq_rel) << 16), qual_rel + qual_len + 1u - seq_rel);
                } else {
                    sa = make_uint2(dA | ((src0 + seq_rel) << 16), seq_len);
                    uint8_t* q = S.out + dA + seq_len;
                    q[0] = '\n';
                    q[1] = '+';
                    q[2] = '\n';
                    sb = make_uint2((dA + seq_len + 3u) | ((src0 + qual_rel) << 16), qual_len);
                    q[3u + qual_len] = '\n';
                    atomicOr(&S.flags, 2u);
                }
            }
        }
        S.seg[0][tid] = sa;
        S.seg[1][tid] = sb;
        S.chk[0][tid] = tile_chunk_desc(sa);
        S.chk[1][tid] = tile_chunk_desc(sb);
        S.head[tid] = hd;
        S.ann[tid] = an;
    }
    __syncthreads();
    mbar_wait(reinterpret_cast<uint64_t*>(&S.bar), 0);
    const uint32_t flags = S.flags;
    if (flags & 1u) {
        emit_block_generic(args, cfg, r0, r1);
        return;
    }

    if (warp < 2) {
        // ---- header without its blanks ----
        const uint32_t k = (uint32_t)tid;
        if (k < nrec && S.ann[k] != 0ull) {
            const uint2 hd = S.head[k];
            const uint32_t cut = hd.y & 0xFFFFu;
            const uint8_t* sp = S.in + (hd.x >> 16);
            uint8_t* dp = S.out + (hd.x & 0xFFFFu);
            for (uint32_t i = 0; i < cut; i++) {
                const uint8_t ch = sp[i];
                if (ch != ' ') *dp++ = ch;
            }
        }
    } else if (warp < 4) {
        // ---- '_' + barcode1 + barcode2 + barcode3 + '_' + umi + '\n' ----
        const uint32_t k = (uint32_t)tid - 64u;
        const unsigned long long an = k < nrec ? S.ann[k] : 0ull;
        if (an != 0ull) {
            uint8_t* q = S.out + (S.head[k].y >> 16);
            uint32_t cell = ann_cell(an);
            if (args.remap) cell = args.remap[cell];
            uint32_t i1, i2, i3;
            cell_indices(cfg, cell, i1, i2, i3);
            const unsigned long long b1 = __ldg(args.wl64 + i1), b2 = __ldg(args.wl64 + kMaxWl + i2), b3 = __ldg(args.wl64 + 2 * kMaxWl + i3);
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
    } else if (warp < 6) {
        // ---- the ragged ends of the long segments ----
        const uint32_t k = (uint32_t)tid - 128u;
        if (k < nrec && S.ann[k] != 0ull) {
            tile_edges(S.in, S.out, S.seg[0][k]);
            if (flags & 2u) tile_edges(S.in, S.out, S.seg[1][k]);
        }
    }

    // ---- whole 16-byte chunks of the long segments: cps slots per record, slot i = (record i / cps, chunk i % cps) ----
    {
        const uint32_t n_items = nrec * args.cps;
        for (uint32_t i = (uint32_t)tid; i < n_items; i += kThreads) {
            const uint32_t k = (i * args.cps_magic) >> 20, j = i - k * args.cps;
            const uint2 ck = S.chk[0][k];
            const uint32_t nfull = ck.y & 0xFFFFu;
            if (j < nfull) tile_chunk(S.in, S.out, ck, j);
            for (uint32_t c = j + args.cps; c < nfull; c += args.cps) tile_chunk(S.in, S.out, ck, c);  // segments longer than the slots
        }
        if (flags & 2u) {
            for (uint32_t i = (uint32_t)tid; i < n_items; i += kThreads) {
                const uint32_t k = (i * args.cps_magic) >> 20, j = i - k * args.cps;
                const uint2 ck = S.chk[1][k];
                for (uint32_t c = j; c < (ck.y & 0xFFFFu); c += args.cps) tile_chunk(S.in, S.out, ck, c);
            }
        }
    }

    // ---- the finished slice leaves with one bulk copy; shared and global addresses agree modulo 16 ----
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const uint32_t nout = (uint32_t)(out_hi - out_lo);
    uint8_t* gdst = args.out + out_lo;
    const uint32_t head = out_skew ? (16u - out_skew < nout ? 16u - out_skew : nout) : 0u;
    const uint32_t body = (nout - head) & ~15u;
    if (tid == 0 && body) bulk_s2g(gdst + head, S.out + out_skew + head, body);
    if ((uint32_t)tid < head) gdst[tid] = S.out[out_skew + tid];
    if (tid >= 32 && tid < 48) {
        const uint32_t i = head + body + (uint32_t)(tid - 32);
        if (i < nout) gdst[i] = S.out[out_skew + i];
    }
    if (tid == 0 && body) bulk_wait_read();
}

// ---------------------------------------------------------------------------------------------
// The same writer as a persistent pipeline: a producer warp describes the records of the next tile and starts the
// bulk copy of its input while eight consumer warps assemble the current tile; the finished slice leaves with a bulk
// copy that overlaps the next tile's assembly (two output buffers).  No thread of the consumers ever waits for global
// memory.
// ---------------------------------------------------------------------------------------------
#ifndef SSD_EMIT_STAGES
#define SSD_EMIT_STAGES 2
#endif
constexpr int kStreamStages = SSD_EMIT_STAGES;
#ifndef SSD_EMIT_OUTBUFS
#define SSD_EMIT_OUTBUFS 2
#endif
constexpr int kStreamOutBufs = SSD_EMIT_OUTBUFS;  // 2: a slice leaves while the next is assembled; 1: the next tile waits for the slice to leave
constexpr int kStreamConsumers = 8;                              // warps
constexpr int kStreamThreads = (kStreamConsumers + 1) * 32;      // + the producer warp
constexpr uint32_t kStreamGeneric = 1u, kStreamTwoSeg = 2u, kStreamEmpty = 4u;

struct __align__(128) StreamStage {
    uint8_t in[kTPad + kTIn + 48];
    unsigned long long ann[kTR];   // annotation of the record, 0 = not written
    uint2 seg[2][kTR];             // long segments {dst | src << 16, length}
    uint2 chk[2][kTR];             // their whole 16-byte chunks
    uint2 head[kTR];               // {dst | src << 16, cut length | literal dst << 16 | two-segment record << 31}
    unsigned long long out_lo, out_hi;
    uint32_t tile, nrec, flags, pad;
};

struct __align__(128) StreamSmem {
    StreamStage st[kStreamStages];
    uint8_t out[kStreamOutBufs][kTOut + 32];
    unsigned long long full[kStreamStages], empty[kStreamStages];
};

__device__ __forceinline__ void sync_consumers() { asm volatile("bar.sync 1, %0;" ::"n"(kStreamConsumers * 32) : "memory"); }

template <typename OffT>
__global__ void __launch_bounds__(kStreamThreads, SSD_EMIT_MINBLOCKS) k_emit_stream(const __grid_constant__ EmitArgs<OffT> args, const __grid_constant__ DevCfg cfg,
                                                                    uint32_t n_tiles)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    StreamSmem& S = *reinterpret_cast<StreamSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < kStreamStages; s++) {
            mbar_init(reinterpret_cast<uint64_t*>(&S.full[s]), 2);  // the input bytes' expect-tx arrival + the descriptors' arrival
            mbar_init(reinterpret_cast<uint64_t*>(&S.empty[s]), 1);
        }
    }
    __syncthreads();

    if (warp == kStreamConsumers) {
        // ================================ producer warp ================================
        // the per-record arrays of a tile are requested one tile ahead (registers), so that the round trip to global memory
        // overlaps the wait for a free stage
        uint64_t o0[2], o1[2], start[2];
        uint4 mt[2];
        unsigned long long an[2];
        auto request = [&](uint64_t t) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                o0[h] = o1[h] = start[h] = 0;
                mt[h] = make_uint4(0u, 0u, 0u, 0u);
                an[h] = 0;
            }
            if (t >= n_tiles) return;
            const uint64_t r0 = t * args.rpb;
            const uint32_t nrec = (uint32_t)(r0 + args.rpb < args.n_rec ? args.rpb : args.n_rec - r0);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const uint32_t k = (uint32_t)lane + 32u * h;
                if (k < nrec) {
                    o0[h] = args.out_off[r0 + k];
                    o1[h] = args.out_off[r0 + k + 1];
                    start[h] = (uint64_t)args.r1_start[r0 + k];
                    mt[h] = args.r1_meta[r0 + k];
                    an[h] = args.ann[r0 + k];
                }
            }
        };
        request(blockIdx.x);
        for (uint32_t it = 0;; it++) {
            const uint32_t s = it % kStreamStages;
            StreamStage& T = S.st[s];
            if (it >= (uint32_t)kStreamStages) mbar_wait(reinterpret_cast<uint64_t*>(&S.empty[s]), ((it / kStreamStages) - 1u) & 1u);
            const uint64_t t = (uint64_t)blockIdx.x + (uint64_t)it * gridDim.x;
            if (t >= n_tiles) {
                if (lane == 0) {
                    T.tile = 0xFFFFFFFFu;
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&S.full[s])) : "memory");
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&S.full[s])) : "memory");
                }
                break;
            }
            const uint64_t r0 = t * args.rpb;
            const uint32_t nrec = (uint32_t)(r0 + args.rpb < args.n_rec ? args.rpb : args.n_rec - r0);
            const uint32_t last = nrec - 1u;
            const uint64_t out_lo = __shfl_sync(0xFFFFFFFFu, o0[0], 0);
            const uint64_t oh0 = __shfl_sync(0xFFFFFFFFu, o1[0], (int)(last & 31u)), oh1 = __shfl_sync(0xFFFFFFFFu, o1[1], (int)(last & 31u));
            const uint64_t out_hi = last < 32u ? oh0 : oh1;
            const uint64_t g0 = __shfl_sync(0xFFFFFFFFu, start[0], 0);
            const uint64_t ls0 = __shfl_sync(0xFFFFFFFFu, start[0], (int)(last & 31u)), ls1 = __shfl_sync(0xFFFFFFFFu, start[1], (int)(last & 31u));
            const uint32_t lz0 = __shfl_sync(0xFFFFFFFFu, mt[0].z, (int)(last & 31u)), lz1 = __shfl_sync(0xFFFFFFFFu, mt[1].z, (int)(last & 31u));
            const uint32_t lz = last < 32u ? lz0 : lz1;
            uint64_t g1 = (last < 32u ? ls0 : ls1) + (lz & 0xFFFFu) + (lz >> 16) + 1u;  // behind the last quality line
            g1 = g1 < (uint64_t)args.r1_len ? g1 : (uint64_t)args.r1_len;
            const uint32_t out_skew = (uint32_t)(((uintptr_t)args.out + out_lo) & 15u);
            const uint32_t mis = (uint32_t)((uintptr_t)(args.r1buf + g0) & 15u);
            const uint32_t in0 = (uint32_t)kTPad + mis;  // T.in[in0 + (x - g0)] holds file byte x
            uint32_t flags = 0;
            if (out_hi == out_lo) flags = kStreamEmpty;  // nothing of this tile survives the filter: no loads at all
            else if (!(nrec <= (uint32_t)kTR && g1 > g0 && g1 - g0 + mis <= (uint64_t)kTIn && out_hi - out_lo + out_skew <= (uint64_t)kTOut)) flags = kStreamGeneric;
            uint32_t bulk = 0;
            if (flags == 0) {
                uint64_t b0, b1;
                if (tile_bulk_range(args.r1buf, g0, g1, b0, b1)) {
                    bulk = (uint32_t)(b1 - b0);
                    if (lane == 0) {
                        mbar_expect_tx(reinterpret_cast<uint64_t*>(&S.full[s]), bulk);
                        bulk_g2s(&T.in[in0 + (uint32_t)(b0 - g0)], args.r1buf + b0, bulk, reinterpret_cast<uint64_t*>(&S.full[s]));
                    }
                    if (lane < 16) {
                        const uint64_t x = g0 + (uint32_t)lane;
                        if (x < b0) T.in[in0 + (uint32_t)lane] = args.r1buf[x];
                    } else {
                        const uint64_t x = b1 + (uint32_t)(lane - 16);
                        if (x < g1) T.in[in0 + (uint32_t)(x - g0)] = args.r1buf[x];
                    }
                } else {
                    const uint64_t x = g0 + (uint32_t)lane;
                    if (x < g1) T.in[in0 + (uint32_t)lane] = args.r1buf[x];
                }
                // where the pieces of every record go
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const uint32_t k = (uint32_t)lane + 32u * h;
                    if (k >= nrec) continue;
                    uint2 sa = make_uint2(0u, 0u), sb = make_uint2(0u, 0u), hd = make_uint2(0u, 0u);
                    unsigned long long a = 0;
                    if (o1[h] != o0[h]) {
                        a = an[h];
                        const uint32_t cut = mt[h].x & 0xFFFFu, kept = mt[h].x >> 16, seq_rel = mt[h].y & 0xFFFFu, seq_len = mt[h].y >> 16,
                                       qual_rel = mt[h].z & 0xFFFFu, qual_len = mt[h].z >> 16;
                        const uint32_t lit = 27u + ann_umi_len(a, cfg.umi_len);
                        if ((mt[h].w & kMetaLong) || start[h] < g0 || start[h] + qual_rel + qual_len + ((mt[h].w & kMetaTail) ? 1u : 0u) > g1) {
                            flags |= kStreamGeneric;  // the whole tile goes the exact way
                            a = 0;
                        } else if (kept + lit + seq_len + qual_len + 4u != (uint32_t)(o1[h] - o0[h])) {
                            atomicOr(&args.cnt->err_flags, kErrEmitLen);
                            a = 0;
                        } else {
                            const uint32_t d = out_skew + (uint32_t)(o0[h] - out_lo), src0 = in0 + (uint32_t)(start[h] - g0), dA = d + kept + lit;
                            hd = make_uint2(d | (src0 << 16), cut | ((d + kept) << 16));
                            if (mt[h].w & kMetaTail) {  // the source already reads  read "\n+\n" quality "\n"
                                sa = make_uint2(dA | ((src0 + seq_rel) << 16), qual_rel + qual_len + 1u - seq_rel);
                            } else {
                                sa = make_uint2(dA | ((src0 + seq_rel) << 16), seq_len);
                                sb = make_uint2((dA + seq_len + 3u) | ((src0 + qual_rel) << 16), qual_len);
                                hd.y |= 0x80000000u;
                                flags |= kStreamTwoSeg;
                            }
                        }
                    }
                    T.seg[0][k] = sa;
                    T.seg[1][k] = sb;
                    T.chk[0][k] = tile_chunk_desc(sa);
                    T.chk[1][k] = tile_chunk_desc(sb);
                    T.head[k] = hd;
                    T.ann[k] = a;
                }
            }
            flags = __reduce_or_sync(0xFFFFFFFFu, flags);
            if (lane == 0) {
                T.tile = (uint32_t)t;
                T.nrec = nrec;
                T.flags = flags;
                T.out_lo = out_lo;
                T.out_hi = out_hi;
            }
            __syncwarp();
            if (lane == 0) {
                if (!bulk) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&S.full[s])) : "memory");
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&S.full[s])) : "memory");
            }
            request(t + gridDim.x);
        }
        return;
    }

    // ================================ consumer warps ================================
    for (uint32_t it = 0;; it++) {
        const uint32_t s = it % kStreamStages;
        const StreamStage& T = S.st[s];
        uint8_t* s_out = S.out[kStreamOutBufs == 2 ? (it & 1u) : 0u];
        mbar_wait(reinterpret_cast<uint64_t*>(&S.full[s]), (it / kStreamStages) & 1u);
        if (T.tile == 0xFFFFFFFFu) break;
        const uint32_t nrec = T.nrec, flags = T.flags;
        const uint64_t out_lo = T.out_lo, out_hi = T.out_hi;
        const bool staged = (flags & (kStreamGeneric | kStreamEmpty)) == 0u;
        if (kStreamOutBufs == 1) {
            if (tid == 0) bulk_wait_read();  // the previous slice has left the buffer
            sync_consumers();
        }
        if (flags & kStreamGeneric) {
            const uint64_t r0 = (uint64_t)T.tile * args.rpb;
            emit_block_generic(args, cfg, r0, r0 + nrec, kStreamConsumers);
        }
        if (staged) {
            if (warp < 2) {
                // ---- header without its blanks ----
                const uint32_t k = (uint32_t)tid;
                if (k < nrec && T.ann[k] != 0ull) {
                    const uint2 hd = T.head[k];
                    const uint32_t cut = hd.y & 0xFFFFu;
                    const uint8_t* sp = T.in + (hd.x >> 16);
                    uint8_t* dp = s_out + (hd.x & 0xFFFFu);
                    for (uint32_t i = 0; i < cut; i++) {
                        const uint8_t ch = sp[i];
                        if (ch != ' ') *dp++ = ch;
                    }
                }
            } else if (warp < 4) {
                // ---- '_' + barcode1 + barcode2 + barcode3 + '_' + umi + '\n' ----
                const uint32_t k = (uint32_t)tid - 64u;
                const unsigned long long an = k < nrec ? T.ann[k] : 0ull;
                if (an != 0ull) {
                    uint8_t* q = s_out + ((T.head[k].y >> 16) & 0x7FFFu);
                    uint32_t cell = ann_cell(an);
                    if (args.remap) cell = args.remap[cell];
                    uint32_t i1, i2, i3;
                    cell_indices(cfg, cell, i1, i2, i3);
                    const unsigned long long b1 = __ldg(args.wl64 + i1), b2 = __ldg(args.wl64 + kMaxWl + i2), b3 = __ldg(args.wl64 + 2 * kMaxWl + i3);
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
            } else if (warp < 6) {
                // ---- the ragged ends of the long segments, the separators of records whose lines are not adjacent ----
                const uint32_t k = (uint32_t)tid - 128u;
                if (k < nrec && T.ann[k] != 0ull) {
                    const uint2 sa = T.seg[0][k];
                    tile_edges(T.in, s_out, sa);
                    if (T.head[k].y & 0x80000000u) {
                        const uint2 sb = T.seg[1][k];
                        tile_edges(T.in, s_out, sb);
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
                const uint32_t n_items = nrec * args.cps;
                for (uint32_t i = (uint32_t)tid; i < n_items; i += kStreamConsumers * 32) {
                    const uint32_t k = (i * args.cps_magic) >> 20, j = i - k * args.cps;
                    const uint2 ck = T.chk[0][k];
                    const uint32_t nfull = ck.y & 0xFFFFu;
                    if (j < nfull) tile_chunk(T.in, s_out, ck, j);
                    if (nfull > args.cps)
                        for (uint32_t c = j + args.cps; c < nfull; c += args.cps) tile_chunk(T.in, s_out, ck, c);  // segments longer than the slots
                }
                if (flags & kStreamTwoSeg) {
                    for (uint32_t i = (uint32_t)tid; i < n_items; i += kStreamConsumers * 32) {
                        const uint32_t k = (i * args.cps_magic) >> 20, j = i - k * args.cps;
                        const uint2 ck = T.chk[1][k];
                        for (uint32_t c = j; c < (ck.y & 0xFFFFu); c += args.cps) tile_chunk(T.in, s_out, ck, c);
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        if (kStreamOutBufs == 2 && tid == 0) bulk_wait_read();  // the slice written two tiles ago has left the other buffer
        sync_consumers();
        // ---- the finished slice leaves with one bulk copy; shared and global addresses agree modulo 16 ----
        if (staged) {
            const uint32_t out_skew = (uint32_t)(((uintptr_t)args.out + out_lo) & 15u);
            const uint32_t nout = (uint32_t)(out_hi - out_lo);
            uint8_t* gdst = args.out + out_lo;
            const uint32_t head = out_skew ? (16u - out_skew < nout ? 16u - out_skew : nout) : 0u;
            const uint32_t body = (nout - head) & ~15u;
            if (tid == 0 && body) bulk_s2g(gdst + head, s_out + out_skew + head, body);
            if ((uint32_t)tid < head) gdst[tid] = s_out[out_skew + tid];
            if (tid >= 32 && tid < 48) {
                const uint32_t i = head + body + (uint32_t)(tid - 32);
                if (i < nout) gdst[i] = s_out[out_skew + i];
            }
        }
        if (tid == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&S.empty[s])) : "memory");
    }
    if (tid == 0) bulk_wait_read();  // shared memory must outlive the last bulk store
}

}  // namespace ssd
