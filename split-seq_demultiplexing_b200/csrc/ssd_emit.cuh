// ssd_emit.cuh — the -m decision, output planning and the annotated-FASTQ writer (sm_100a).
//
//   k_filter / k_collapse_map   per-cell pass decision (V2:409-412), RanHex->OligoDT map (Collapse_RanHex_Odt.py:44-71)
//   OutLen + exclusive scan      byte offset of every record's output (prefix-sum compaction)
//   k_emit                       the writer (V2:134-141, 367-370, 415-454): eight lanes per record read the
//                                read-1 record straight from global memory with 16-byte loads, assemble the
//                                output record in a shared-memory staging buffer that holds the block's
//                                contiguous slice of the output, and the block flushes that slice with aligned
//                                16-byte stores.  Records that do not fit the staging buffer are written by
//                                one warp each, directly to global memory.
#pragma once
#include "ssd_scan.cuh"

namespace ssd {

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
    // one pair of atomics per block
    const int n1 = __syncthreads_count(present), n2 = __syncthreads_count(ok);
    if (threadIdx.x == 0) {
        if (n1) atomicAdd(&cnt->n_cells, (unsigned long long)n1);
        if (n2) atomicAdd(&cnt->n_passing_cells, (unsigned long long)n2);
    }
}

// n1, n23: size of the round-1 list, product of the sizes of the round-2 and round-3 lists
__global__ void k_collapse_map(const uint8_t* pass, int n1, int n23, int pairs, uint32_t* remap)
{
    uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t total = (uint64_t)n1 * n23;
    if (c >= total) return;
    uint32_t i1 = (uint32_t)(c / (uint64_t)n23);
    uint32_t out = (uint32_t)c;
    if (i1 >= (uint32_t)pairs && i1 < 2u * (uint32_t)pairs && pass[c]) {
        uint64_t odt = c - (uint64_t)pairs * n23;
        if (pass[odt]) out = (uint32_t)odt;
    }
    remap[c] = out;
}

// ---------------------------------------------------------------------------------------------
// output planning: length of each record's output (0 when not emitted)
// ---------------------------------------------------------------------------------------------
struct OutLen {
    const uint64_t* ann;
    const uint32_t* r1_base;  // header' + seq + qual + 7 fixed bytes, from the read-1 scan
    const uint8_t* pass;      // NULL = every matched record (MergedCells_1)
    const uint32_t* remap;    // NULL = no collapse
    const uint8_t* wl_len;    // device copy of the per-entry whitelist lengths: kMaxWl per round
    int n2, n3, umi_len, bc_const;  // bc_const > 0: the three entries of every cell have that many bytes together, added without lookups
    int copy_mode;            // annotated input: records are copied as they are, r1_base already is the length
    __device__ __forceinline__ uint32_t bc_len(uint32_t cell, bool collapse) const
    {
        if (bc_const) return (uint32_t)bc_const;
        if (collapse && remap) cell = remap[cell];
        const uint32_t i3 = cell % (uint32_t)n3, i2 = (cell / (uint32_t)n3) % (uint32_t)n2, i1 = cell / ((uint32_t)n2 * (uint32_t)n3);
        return wl_len[i1] + wl_len[kMaxWl + i2] + wl_len[2 * kMaxWl + i3];
    }
    __device__ __forceinline__ uint32_t operator()(uint64_t r) const
    {
        const uint64_t a = ann[r];
        if (ann_state(a) == 0) return 0u;
        const uint32_t cell = ann_cell(a);
        if (pass && !pass[cell]) return 0u;
        if (copy_mode) return r1_base[r];
        return r1_base[r] + bc_len(cell, true) + ann_umi_len(a, umi_len);
    }
    // length in the pre-filter output (never collapsed) and in the post-filter output
    __device__ __forceinline__ void both(uint64_t r, uint32_t& matched, uint32_t& passing) const
    {
        const uint64_t a = ann[r];
        matched = passing = 0u;
        if (ann_state(a) == 0) return;
        const uint32_t cell = ann_cell(a);
        if (copy_mode) {
            matched = r1_base[r];
            if (pass[cell]) passing = matched;
            return;
        }
        const uint32_t common = r1_base[r] + ann_umi_len(a, umi_len);
        matched = common + bc_len(cell, false);
        if (pass[cell]) passing = common + bc_len(cell, true);
    }
};

// Block sums of the post-filter lengths (first kernel of their exclusive scan) plus, on the side, the
// number of retained records and the size of the pre-filter output.
__global__ void __launch_bounds__(kThreads) k_plan_reduce(OutLen f, uint64_t n, unsigned long long* block_sums, DevCounters* cnt)
{
    __shared__ unsigned long long s_tmp[kThreads / 32];
    const uint64_t base = (uint64_t)blockIdx.x * kScanBlock + (uint64_t)threadIdx.x * kScanItems;
    unsigned long long v = 0, vm = 0;
    uint32_t nz = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++)
        if (base + k < n) {
            uint32_t lm, lp;
            f.both(base + k, lm, lp);
            v += lp;
            vm += lm;
            nz += lp ? 1u : 0u;
        }
    unsigned long long total;
    block_exclusive_scan(v, s_tmp, total);
    nz = __reduce_add_sync(0xFFFFFFFFu, nz);
#pragma unroll
    for (int d = 16; d; d >>= 1) vm += __shfl_xor_sync(0xFFFFFFFFu, vm, d);
    if ((threadIdx.x & 31) == 0) {
        if (nz) atomicAdd(&cnt->n_retained, (unsigned long long)nz);
        if (vm) atomicAdd(&cnt->bytes_matched, vm);
    }
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// The same in one pass: exclusive scan of the post-filter lengths with a decoupled look-back over per-tile sums
// (tiles taken by ticket), retained-record count and pre-filter size on the side.  status: one u64 per tile, zeroed.
constexpr unsigned long long kPlanAgg = 1ull << 62, kPlanPrefix = 2ull << 62, kPlanValue = (1ull << 62) - 1ull;
__global__ void __launch_bounds__(kThreads) k_plan_onepass(OutLen f, uint64_t n, unsigned long long* out_off, unsigned long long* status,
                                                           uint32_t* ticket, DevCounters* cnt)
{
    __shared__ unsigned long long s_tmp[kThreads / 32];
    __shared__ unsigned long long s_prior;
    __shared__ uint32_t s_tile;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t base = (uint64_t)tile * kScanBlock + (uint64_t)threadIdx.x * kScanItems;
    uint32_t len[kScanItems];
    unsigned long long v = 0, vm = 0;
    uint32_t nz = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        uint32_t lm = 0, lp = 0;
        if (base + k < n) f.both(base + k, lm, lp);
        len[k] = lp;
        v += lp;
        vm += lm;
        nz += lp ? 1u : 0u;
    }
    unsigned long long total;
    unsigned long long run = block_exclusive_scan(v, s_tmp, total);
    if (threadIdx.x == 0) {
        volatile unsigned long long* st = status;
        if (tile != 0) st[tile] = kPlanAgg | total;
        unsigned long long prior = 0;
        for (uint32_t t = tile; t > 0;) {
            const unsigned long long x = st[t - 1];
            if ((x >> 62) == 0ull) continue;
            prior += x & kPlanValue;
            if ((x >> 62) == 2ull) break;
            t--;
        }
        st[tile] = kPlanPrefix | (prior + total);
        s_prior = prior;
        if ((uint64_t)(tile + 1) * kScanBlock >= n) {  // the last tile closes the array
            out_off[n] = prior + total;
            cnt->bytes_passing = prior + total;
        }
    }
    nz = __reduce_add_sync(0xFFFFFFFFu, nz);
#pragma unroll
    for (int d = 16; d; d >>= 1) vm += __shfl_xor_sync(0xFFFFFFFFu, vm, d);
    if ((threadIdx.x & 31) == 0) {
        if (nz) atomicAdd(&cnt->n_retained, (unsigned long long)nz);
        if (vm) atomicAdd(&cnt->bytes_matched, vm);
    }
    __syncthreads();
    run += s_prior;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        if (base + k < n) out_off[base + k] = run;
        run += len[k];
    }
}

// ---------------------------------------------------------------------------------------------
// the writer
// ---------------------------------------------------------------------------------------------
constexpr int kEmitOut = 32768;  // staged output bytes per block
constexpr int kGroup = 8;        // lanes per record

template <typename OffT>
struct EmitArgs {
    const uint8_t* r1buf;
    size_t r1_len;
    const uint8_t* r2buf;
    size_t r2_len;
    const OffT* r1_start;
    const uint4* r1_meta;
    const uint64_t* ann;
    const uint64_t* out_off;  // n_rec + 1 entries
    const uint32_t* remap;    // NULL = none
    uint8_t* out;
    uint64_t n_rec;           // records of this batch (min of the two files)
    uint32_t rpb;             // records per block
    const uint32_t* order;    // split mode: position k of the output holds record order[k] (NULL = input order)
    int copy_mode;            // annotated input: header, read, third line and quality are copied strip()ped (V2:435-447)
    DevCounters* cnt;
    const unsigned long long* wl64;  // tiled writer: the whitelist entries as 8-byte words, kMaxWl per round
    uint32_t cps, cps_magic;         // tiled writer: 16-byte chunk slots per record, ceil(2^20 / cps)
};

// One warp writes one output record (V2:134-141).  `dst` is a generic pointer (shared staging buffer or
// global memory) to the first byte of the record's output; returns the number of bytes written.
template <class Acc, typename OffT>
__device__ __forceinline__ uint32_t emit_record_generic(const Acc& a, const EmitArgs<OffT>& args, const DevCfg& cfg, uint64_t r, uint8_t* dst)
{
    const int lane = threadIdx.x & 31;
    const size_t start = (size_t)args.r1_start[r];
    Lines L;
    find_lines(a, start, args.r1_len, true, L);
    uint32_t o = 0;
    if (args.copy_mode) {  // the four lines, strip()ped
        for (int k = 0; k < 4; k++) {
            size_t le = rstrip_end(a, L.s[k], L.e[k]), ls = L.s[k];
            while (ls < le && py_space(a.byte(ls))) ls++;
            for (size_t i = lane; i < le - ls; i += 32) dst[o + i] = (uint8_t)a.byte(ls + i);
            o += (uint32_t)(le - ls);
            if (lane == 0) dst[o] = '\n';
            o += 1;
        }
        return o;
    }

    size_t tok1_s = 0, tok1_e = 0;  // header style 1: the second token follows the UMI
    if (cfg.header_style == 1) {
        // Extract_BC_UMI.py:78-85: ReadName.split()[0], then _cell_umi, a blank, ReadName.split()[1]
        size_t p = L.s[0], t0s, t0e;
        while (p < L.e[0] && py_space(a.byte(p))) p++;
        t0s = p;
        while (p < L.e[0] && !py_space(a.byte(p))) p++;
        t0e = p;
        while (p < L.e[0] && py_space(a.byte(p))) p++;
        tok1_s = p;
        while (p < L.e[0] && !py_space(a.byte(p))) p++;
        tok1_e = p;
        for (size_t i = lane; i < t0e - t0s; i += 32) dst[o + i] = (uint8_t)a.byte(t0s + i);
        o += (uint32_t)(t0e - t0s);
    }
    // header: bytes before the first '/', blanks removed, then strip()
    size_t cut = cfg.header_style == 1 ? L.s[0] : L.e[0];
    for (size_t p = L.s[0]; p < cut; p += 32) {
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
    uint32_t ci1, ci2, ci3;
    cell_indices(cfg, cell, ci1, ci2, ci3);
    const uint32_t idx[3] = {ci1, ci2, ci3};
    if (lane == 0) dst[o] = '_';
    o += 1;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        uint32_t wl = cfg.wl_len[k][idx[k]];
        if ((uint32_t)lane < wl) dst[o + lane] = cfg.wl[k][idx[k] * 8 + lane];
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
    if (cfg.header_style == 1) {
        if (lane == 0) dst[o] = ' ';
        o += 1;
        for (size_t i = lane; i < tok1_e - tok1_s; i += 32) dst[o + i] = (uint8_t)a.byte(tok1_s + i);
        o += (uint32_t)(tok1_e - tok1_s);
    }
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


__device__ __forceinline__ uint4 ldg16_safe(const uint8_t* buf, size_t off, size_t len)
{
    if (off + 16 <= len) return __ldg(reinterpret_cast<const uint4*>(buf + off));
    uint32_t w[4] = {0, 0, 0, 0};
    for (int b = 0; b < 16; b++)
        if (off + b < len) w[b >> 2] |= (uint32_t)buf[off + b] << (8 * (b & 3));
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// Copy n bytes from global memory (file offset src) to shared memory (byte offset d of s_out) with one
// eight-lane group: byte copies up to the next 16-byte boundary of the destination, 16-byte chunks
// (two aligned 16-byte loads funnel-shifted into one aligned 16-byte store), byte copies for the rest.
// Head and tail are at most 15 bytes each: two predicated byte copies per lane, no variable-trip loops.
__device__ __forceinline__ void group_copy(uint8_t* s_out, uint32_t d, const uint8_t* buf, size_t src, size_t len, uint32_t n, int gl)
{
    uint32_t head = (16u - (d & 15u)) & 15u;
    head = head < n ? head : n;
    const uint8_t* sp = buf + src;
    uint8_t* dp = s_out + d;
#pragma unroll
    for (int t = 0; t < 2; t++) {
        const uint32_t i = (uint32_t)gl + 8u * t;
        if (i < head) dp[i] = sp[i];
    }
    const uint32_t nchunks = (n - head) >> 4;
    const size_t s = src + head;
    const size_t sal = s & ~(size_t)15;
    const uint32_t ws = (uint32_t)(s & 15) >> 2, bs = ((uint32_t)s & 3u) * 8u;
    uint4* dst = reinterpret_cast<uint4*>(dp + head);
    const bool inside = sal + 16u * (size_t)(nchunks + 1u) <= len;  // every 16-byte load below stays inside the file
    const uint4* src16 = reinterpret_cast<const uint4*>(buf + sal);
    for (uint32_t c = gl; c < nchunks; c += kGroup) {
        uint4 A, B;
        if (inside) {
            A = __ldg(src16 + c);
            B = __ldg(src16 + c + 1);
        } else {
            A = ldg16_safe(buf, sal + 16u * c, len);
            B = ldg16_safe(buf, sal + 16u * c + 16u, len);
        }
        // words [ws, ws + 5) of A:B, selected without branches (ws differs between the groups of a warp)
        const bool s2 = ws & 2u, s1 = ws & 1u;
        const uint32_t t0 = s2 ? A.z : A.x, t1 = s2 ? A.w : A.y, t2 = s2 ? B.x : A.z, t3 = s2 ? B.y : A.w, t4 = s2 ? B.z : B.x,
                       t5 = s2 ? B.w : B.y;
        const uint32_t w0 = s1 ? t1 : t0, w1 = s1 ? t2 : t1, w2 = s1 ? t3 : t2, w3 = s1 ? t4 : t3, w4 = s1 ? t5 : t4;
        dst[c] = make_uint4(__funnelshift_r(w0, w1, bs), __funnelshift_r(w1, w2, bs), __funnelshift_r(w2, w3, bs), __funnelshift_r(w3, w4, bs));
    }
    const uint32_t done = head + (nchunks << 4), tail = n - done;
#pragma unroll
    for (int t = 0; t < 2; t++) {
        const uint32_t i = (uint32_t)gl + 8u * t;
        if (i < tail) dp[done + i] = sp[done + i];
    }
}

// Eight lanes write one output record into the staging buffer at byte offset d; returns its length.
template <typename OffT>
__device__ __forceinline__ uint32_t emit_group(const EmitArgs<OffT>& args, const DevCfg& cfg, uint64_t r, uint8_t* s_out, uint32_t d, int gl,
                                               uint32_t gmask)
{
    const size_t start = (size_t)args.r1_start[r];
    const uint4 mt = args.r1_meta[r];
    const uint32_t cut_len = mt.x & 0xFFFFu, seq_rel = mt.y & 0xFFFFu, seq_len = mt.y >> 16, qual_rel = mt.z & 0xFFFFu, qual_len = mt.z >> 16;
    uint32_t o = 0;

    if (args.copy_mode) {
        const uint32_t plus_rel = mt.w >> 16, plus_len = (mt.w >> 1) & 0x7FFFu;
        group_copy(s_out, d, args.r1buf, start, args.r1_len, cut_len, gl);
        o = cut_len;
        if (gl == 0) s_out[d + o] = '\n';
        o += 1u;
        group_copy(s_out, d + o, args.r1buf, start + seq_rel, args.r1_len, seq_len, gl);
        o += seq_len;
        if (gl == 0) s_out[d + o] = '\n';
        o += 1u;
        group_copy(s_out, d + o, args.r1buf, start + plus_rel, args.r1_len, plus_len, gl);
        o += plus_len;
        if (gl == 0) s_out[d + o] = '\n';
        o += 1u;
        group_copy(s_out, d + o, args.r1buf, start + qual_rel, args.r1_len, qual_len, gl);
        o += qual_len;
        if (gl == 0) s_out[d + o] = '\n';
        return o + 1u;
    }

    // header: non-blank bytes of the first cut_len bytes (name.split("/")[0].replace(" ", "").strip())
    {
        const uint32_t a = (uint32_t)(start & 3);
        const size_t al = start - a;
        const uint32_t nbytes = a + cut_len;
        for (uint32_t c0 = 0; c0 < nbytes; c0 += 32u) {
            const uint32_t wpos = c0 + 4u * (uint32_t)gl;  // first byte of this lane's word, relative to al
            uint32_t x = 0, nib = 0;
            if (wpos < nbytes) {
                x = ldg_word_safe(args.r1buf, al + wpos, args.r1_len);
                const uint32_t lo = wpos < a ? a - wpos : 0u, rem = nbytes - wpos;
                const uint32_t V = byte_range_flags(lo < 4u ? lo : 4u, rem < 4u ? rem : 4u);
                const uint32_t K = ~zero_flags(x ^ 0x20202020u) & V;  // kept: inside the range and not a blank
                nib = (K * 0x00204081u) >> 28;
            }
            const uint32_t all = __reduce_or_sync(gmask, nib << (4 * gl));
#pragma unroll
            for (uint32_t b = 0; b < 4; b++)
                if ((nib >> b) & 1u) s_out[d + o + (uint32_t)__popc(all & ((1u << (4u * gl + b)) - 1u))] = (uint8_t)(x >> (8u * b));
            o += (uint32_t)__popc(all);
        }
    }

    // '_' + barcode1 + barcode2 + barcode3 + '_' + umi + '\n'
    const uint64_t an = args.ann[r];
    uint32_t cell = ann_cell(an);
    if (args.remap) cell = args.remap[cell];
    uint32_t i1, i2, i3;
    cell_indices(cfg, cell, i1, i2, i3);
    const bool regular = ann_state(an) == 1;
    const uint64_t payload = ann_payload(an);
    if (cfg.bc8 && regular && cfg.umi_len == 10) {
        // fixed layout, 37 bytes: k = 0 '_', 1..24 barcodes, 25 '_', 26..35 UMI, 36 '\n'; lane gl writes k = gl + 8t
        uint8_t* q = s_out + d + o;
#pragma unroll
        for (uint32_t t = 0; t < 3; t++) {
            const uint32_t k = (uint32_t)gl + 8u * t;
            uint32_t ch = '_';
            if (k) {
                const uint32_t which = (k - 1u) >> 3, idx = which == 0u ? i1 : (which == 1u ? i2 : i3);
                ch = cfg.wl[which][idx * 8u + ((k - 1u) & 7u)];
            }
            q[k] = (uint8_t)ch;
        }
        {
            const uint32_t k = 24u + (uint32_t)gl;  // 24: last barcode byte, 25: '_', 26..31: UMI 0..5
            uint32_t ch = gl == 0 ? (uint32_t)cfg.wl[2][i3 * 8u + 7u] : (uint32_t)'_';
            if (gl >= 2) ch = __byte_perm(0x47544341u, 0u, (uint32_t)(payload >> (2u * (uint32_t)(gl - 2))) & 3u) & 0xFFu;
            q[k] = (uint8_t)ch;
        }
        if (gl < 5) {  // 32..35: UMI 6..9, 36: '\n'
            uint32_t ch = '\n';
            if (gl < 4) ch = __byte_perm(0x47544341u, 0u, (uint32_t)(payload >> (2u * (uint32_t)(gl + 6))) & 3u) & 0xFFu;
            q[32 + gl] = (uint8_t)ch;
        }
        o += 37u;
    } else {
        const uint32_t l1 = cfg.wl_len[0][i1], l2 = cfg.wl_len[1][i2], l3 = cfg.wl_len[2][i3];
        const uint32_t ulen = ann_umi_len(an, cfg.umi_len);
        const uint32_t c1 = 1u, c2 = c1 + l1, c3 = c2 + l2, c4 = c3 + l3, c5 = c4 + 1u, c6 = c5 + ulen, lit = c6 + 1u;
        for (uint32_t k = gl; k < lit; k += kGroup) {
            uint32_t ch;
            if (k == 0u || k == c4) ch = '_';
            else if (k < c2) ch = cfg.wl[0][i1 * 8u + (k - c1)];
            else if (k < c3) ch = cfg.wl[1][i2 * 8u + (k - c2)];
            else if (k < c4) ch = cfg.wl[2][i3 * 8u + (k - c3)];
            else if (k < c6) ch = regular ? (uint32_t)base2_char((uint32_t)(payload >> (2u * (k - c5)))) : (uint32_t)args.r2buf[(size_t)(payload >> 5) + (k - c5)];
            else ch = '\n';
            s_out[d + o + k] = (uint8_t)ch;
        }
        o += lit;
    }

    // read, "+", quality — each rstrip()ped (V2:246, 249)
    if (mt.w & kMetaTail) {  // the source already reads  read "\n+\n" quality "\n"
        const uint32_t tail_len = qual_rel + qual_len + 1u - seq_rel;
        group_copy(s_out, d + o, args.r1buf, start + seq_rel, args.r1_len, tail_len, gl);
        return o + tail_len;
    }
    group_copy(s_out, d + o, args.r1buf, start + seq_rel, args.r1_len, seq_len, gl);
    o += seq_len;
    if (gl < 3) s_out[d + o + gl] = (uint8_t)"\n+\n"[gl];
    o += 3u;
    group_copy(s_out, d + o, args.r1buf, start + qual_rel, args.r1_len, qual_len, gl);
    o += qual_len;
    if (gl == 0) s_out[d + o] = '\n';
    return o + 1u;
}

template <typename OffT>
__global__ void __launch_bounds__(kThreads) k_emit(const __grid_constant__ EmitArgs<OffT> args, const __grid_constant__ DevCfg cfg)
{
    __shared__ __align__(128) uint8_t s_out[kEmitOut + 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t r0 = (uint64_t)blockIdx.x * args.rpb;
    if (r0 >= args.n_rec) return;
    const uint64_t r1 = r0 + args.rpb < args.n_rec ? r0 + args.rpb : args.n_rec;
    const uint64_t out_lo = args.out_off[r0], out_hi = args.out_off[r1];
    if (out_hi == out_lo) return;  // nothing of this block survives the filter: no loads at all

    const uint32_t out_skew = (uint32_t)(((uintptr_t)args.out + out_lo) & 15u);
    const bool staged = cfg.header_style == 0 && out_hi - out_lo + out_skew <= (uint64_t)kEmitOut;  // the other header style: warp per record

    if (!staged) {
        // records too long for the staging buffer: one warp per record, global memory to global memory
        GlobAcc ga{args.r1buf, args.r1_len};
        for (uint64_t k = r0 + warp; k < r1; k += kThreads / 32) {
            uint64_t o0 = args.out_off[k], o1 = args.out_off[k + 1];
            if (o1 == o0) continue;
            const uint64_t r = args.order ? args.order[k] : k;
            uint32_t wrote = emit_record_generic(ga, args, cfg, r, args.out + o0);
            if (lane == 0 && wrote != (uint32_t)(o1 - o0)) atomicOr(&args.cnt->err_flags, kErrEmitLen);
        }
        return;
    }

    const int gl = lane & (kGroup - 1);
    const uint32_t gmask = 0xFFu << (lane & ~(kGroup - 1));
    for (uint64_t k = r0 + (uint32_t)(tid / kGroup); k < r1; k += kThreads / kGroup) {
        const uint64_t o0 = args.out_off[k], o1 = args.out_off[k + 1];
        if (o1 == o0) continue;
        const uint64_t r = args.order ? args.order[k] : k;
        const uint32_t wrote = emit_group(args, cfg, r, s_out, out_skew + (uint32_t)(o0 - out_lo), gl, gmask);
        if (gl == 0 && wrote != (uint32_t)(o1 - o0)) atomicOr(&args.cnt->err_flags, kErrEmitLen);
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
// split mode: per-cell buckets (demultiplex_using_barcodes.py:350-372 keeps one buffer per cell)
// ---------------------------------------------------------------------------------------------
struct PassFlag {  // 1 for records that are written to the post-filter output
    OutLen f;
    __device__ __forceinline__ uint32_t operator()(uint64_t r) const { return f(r) ? 1u : 0u; }
};

// (cell, record) pairs of the written records, in input order: the input of a stable sort by cell
__global__ void k_split_pairs(OutLen f, const uint32_t* pos, uint64_t n, uint64_t* cells, uint32_t* recs)
{
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n || !f(r)) return;
    uint32_t cell = ann_cell(f.ann[r]);
    if (f.remap) cell = f.remap[cell];
    cells[pos[r]] = cell;
    recs[pos[r]] = (uint32_t)r;
}

struct SplitLen {  // output length of the record at position k of the bucketed order
    OutLen f;
    const uint32_t* order;
    __device__ __forceinline__ uint32_t operator()(uint64_t k) const { return f(order[k]); }
};

struct BucketHead {  // 1 where a new cell starts in the sorted (cell) array
    const uint64_t* cells;
    __device__ __forceinline__ uint32_t operator()(uint64_t k) const { return k == 0 || cells[k] != cells[k - 1] ? 1u : 0u; }
};

// bucket table: for every run of equal cells its cell id, first byte and first position
__global__ void k_bucket_table(const uint64_t* cells, const uint32_t* bpos, const unsigned long long* out_off, uint64_t n, uint32_t* b_cell,
                               unsigned long long* b_off, unsigned long long* b_first)
{
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    if (k == 0 || cells[k] != cells[k - 1]) {
        const uint32_t b = bpos[k];
        b_cell[b] = (uint32_t)cells[k];
        b_off[b] = out_off[k];
        b_first[b] = k;
    }
}

__global__ void k_record_cells(const uint64_t* ann, uint64_t n, int32_t* cells)
{
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) cells[r] = ann_state(ann[r]) ? (int32_t)ann_cell(ann[r]) : -1;
}

}  // namespace ssd
