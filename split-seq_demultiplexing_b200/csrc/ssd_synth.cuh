// ssd_synth.cuh — counter-based synthetic SPLiT-seq read pairs (SURVEY.md 8d "synthetic record shape").
// Integer-only, so the device generator and its host twin produce identical bytes for any index range.
//
// Read 1:  "@SYN.<i> <i>/1\n" + L uniform ACGT + "\n+\n" + L quality chars + "\n"
// Read 2:  "@SYN.<i> <i>/2\n" + UMI(10) BC3(8) LINK(30) BC2(8) LINK(30) BC1(8) tail(L-94) + "\n+\n" + quality + "\n"
// with planted cells drawn from an octave ("Zipf-like", density ~ 1/rank) distribution over a pool of
// random (i1,i2,i3) triples, substitutions inside the barcode windows, N bases, junk barcodes,
// low-entropy UMIs (so reads != distinct UMIs), exact 2/2 ties and truncated read 2.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SSD_HD __host__ __device__ __forceinline__
#else
#define SSD_HD inline
#endif

namespace ssd_synth {

struct Params {
    uint64_t seed;
    int32_t read_len;
    int32_t n_wl;
    uint32_t cell_pool;
    uint32_t cell_octaves;  // floor(log2(cell_pool)) + 1
    uint32_t p_sub, p_n, p_junk, p_dup, p_tie, p_trunc, p_atqual;  // ppm
    uint32_t n_ties;        // entries in the tie table (8-mers at distance 2/2 from a whitelist pair)
};

enum Field : uint64_t {
    F_R1SEQ = 0x1000, F_R1Q = 0x2000, F_R2Q = 0x3000, F_TAIL = 0x4000, F_N = 0x5000, F_SUB = 0x6000,
    F_CELL = 1, F_CELL2 = 2, F_JUNK = 3, F_JUNKBC = 4, F_UMI = 5, F_DUP = 6, F_DUPSEL = 7, F_TIE = 8, F_TIESEL = 9,
    F_TRUNC = 10, F_TRUNCLEN = 11, F_ATQ1 = 12, F_ATQ2 = 13, F_POOL = 0x77
};

SSD_HD uint64_t mix(uint64_t seed, uint64_t idx, uint64_t field)
{
    uint64_t z = seed + idx * 0x9E3779B97F4A7C15ull + field * 0xD1B54A32D192ED03ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

SSD_HD uint32_t ppm(uint64_t h) { return (uint32_t)(h % 1000000ull); }

SSD_HD int digits10(uint64_t v)
{
    int d = 1;
    while (v >= 10) {
        v /= 10;
        d++;
    }
    return d;
}

SSD_HD int r2_seq_len(const Params& p, uint64_t i)
{
    if (p.p_trunc && ppm(mix(p.seed, i, F_TRUNC)) < p.p_trunc) return 40 + (int)(mix(p.seed, i, F_TRUNCLEN) % 54ull);
    return p.read_len;
}

SSD_HD uint32_t header_len(uint64_t i) { return (uint32_t)(9 + 2 * digits10(i)); }  // incl. '\n'
SSD_HD uint32_t r1_record_len(const Params& p, uint64_t i) { return header_len(i) + 2u * (uint32_t)p.read_len + 4u; }
SSD_HD uint32_t r2_record_len(const Params& p, uint64_t i) { return header_len(i) + 2u * (uint32_t)r2_seq_len(p, i) + 4u; }

// Per-record plan for read 2, computed once per record.
struct R2Plan {
    int len;            // sequence length
    uint32_t bc[3];     // 16-bit 2-bit codes (position j at bits 2j) of BC1, BC2, BC3 after substitutions
    uint32_t umi;       // 20-bit code
};

SSD_HD uint32_t code_of(const char* s8)
{
    uint32_t c = 0;
    for (int j = 0; j < 8; j++) {
        char ch = s8[j];
        uint32_t v = ch == 'A' ? 0u : ch == 'C' ? 1u : ch == 'G' ? 2u : 3u;
        c |= v << (2 * j);
    }
    return c;
}

SSD_HD R2Plan plan_r2(const Params& p, uint64_t i, const char* wl, const char* ties)
{
    R2Plan pl;
    pl.len = r2_seq_len(p, i);
    bool junk = ppm(mix(p.seed, i, F_JUNK)) < p.p_junk;
    uint64_t rank = 0;
    if (junk) {
        uint64_t h = mix(p.seed, i, F_JUNKBC);
        pl.bc[0] = (uint32_t)(h & 0xFFFF);
        pl.bc[1] = (uint32_t)((h >> 16) & 0xFFFF);
        pl.bc[2] = (uint32_t)((h >> 32) & 0xFFFF);
    } else {
        uint32_t oct = (uint32_t)(mix(p.seed, i, F_CELL) % p.cell_octaves);
        rank = ((1ull << oct) - 1 + (mix(p.seed, i, F_CELL2) & ((1ull << oct) - 1))) % p.cell_pool;
        uint64_t h = mix(p.seed ^ 0xC3115EEDull, rank, F_POOL);
        uint32_t i1 = (uint32_t)(h % (uint64_t)p.n_wl), i2 = (uint32_t)((h >> 20) % (uint64_t)p.n_wl),
                 i3 = (uint32_t)((h >> 40) % (uint64_t)p.n_wl);
        pl.bc[0] = code_of(wl + 8 * i1);
        pl.bc[1] = code_of(wl + 8 * i2);
        pl.bc[2] = code_of(wl + 8 * i3);
        if (p.n_ties && ppm(mix(p.seed, i, F_TIE)) < p.p_tie)
            pl.bc[1] = code_of(ties + 8 * (mix(p.seed, i, F_TIESEL) % p.n_ties));
    }
    for (int w = 0; w < 3; w++)
        for (int j = 0; j < 8; j++) {
            uint64_t h = mix(p.seed, i, F_SUB + (uint64_t)(w * 8 + j));
            if (ppm(h) < p.p_sub) {
                uint32_t old = (pl.bc[w] >> (2 * j)) & 3u;
                uint32_t nw = (old + 1u + (uint32_t)((h >> 32) % 3ull)) & 3u;
                pl.bc[w] = (pl.bc[w] & ~(3u << (2 * j))) | (nw << (2 * j));
            }
        }
    if (!junk && ppm(mix(p.seed, i, F_DUP)) < p.p_dup)
        pl.umi = (uint32_t)(mix(p.seed ^ 0xD0BULL, rank, F_DUPSEL + (mix(p.seed, i, F_DUPSEL) & 3ull)) & 0xFFFFF);
    else
        pl.umi = (uint32_t)(mix(p.seed, i, F_UMI) & 0xFFFFF);
    return pl;
}

SSD_HD char base_char(uint32_t code) { return "ACGT"[code & 3u]; }

// byte j (0-based, within the sequence line) of read 2
SSD_HD char r2_base(const Params& p, uint64_t i, const R2Plan& pl, int j)
{
    const char* l1 = "GTGGCCGCTGTTTCGCATCGGCGTACGACT";
    const char* l2 = "ATCCACGTGCTTGAGAGGCCAGAGCATTCG";
    char c;
    if (j < 10) c = base_char(pl.umi >> (2 * j));
    else if (j < 18) c = base_char(pl.bc[2] >> (2 * (j - 10)));
    else if (j < 48) c = l1[j - 18];
    else if (j < 56) c = base_char(pl.bc[1] >> (2 * (j - 48)));
    else if (j < 86) c = l2[j - 56];
    else if (j < 94) c = base_char(pl.bc[0] >> (2 * (j - 86)));
    else c = base_char((uint32_t)(mix(p.seed, i, F_TAIL + (uint64_t)(j >> 5)) >> (2 * (j & 31))));
    if (p.p_n && ppm(mix(p.seed, i, F_N + (uint64_t)j)) < p.p_n) c = 'N';
    return c;
}

SSD_HD char qual_char(uint64_t h, int j) { return "EEEEEEEEEAAA6/<<"[(h >> (4 * (j & 15))) & 15u]; }

SSD_HD char r2_qual(const Params& p, uint64_t i, const R2Plan& pl, int j)
{
    if (j == 0 && ppm(mix(p.seed, i, F_ATQ2)) < p.p_atqual) return '@';
    if (r2_base(p, i, pl, j) == 'N') return '#';
    return qual_char(mix(p.seed, i, F_R2Q + (uint64_t)(j >> 4)), j);
}

SSD_HD char r1_base(const Params& p, uint64_t i, int j) { return base_char((uint32_t)(mix(p.seed, i, F_R1SEQ + (uint64_t)(j >> 5)) >> (2 * (j & 31)))); }

SSD_HD char r1_qual(const Params& p, uint64_t i, int j)
{
    if (j == 0 && ppm(mix(p.seed, i, F_ATQ1)) < p.p_atqual) return '@';
    return qual_char(mix(p.seed, i, F_R1Q + (uint64_t)(j >> 4)), j);
}

// byte k of the header line "@SYN.<i> <i>/<mate>\n"
SSD_HD char header_char(uint64_t i, int mate, uint32_t k)
{
    int d = digits10(i);
    if (k < 5) return "@SYN."[k];
    k -= 5;
    if (k < (uint32_t)d) {
        uint64_t v = i;
        for (int t = 0; t < d - 1 - (int)k; t++) v /= 10;
        return (char)('0' + v % 10);
    }
    k -= (uint32_t)d;
    if (k == 0) return ' ';
    k -= 1;
    if (k < (uint32_t)d) {
        uint64_t v = i;
        for (int t = 0; t < d - 1 - (int)k; t++) v /= 10;
        return (char)('0' + v % 10);
    }
    k -= (uint32_t)d;
    if (k == 0) return '/';
    if (k == 1) return (char)('0' + mate);
    return '\n';
}

// byte k of record i of read 1 / read 2
SSD_HD char r1_byte(const Params& p, uint64_t i, uint32_t k)
{
    uint32_t h = header_len(i), L = (uint32_t)p.read_len;
    if (k < h) return header_char(i, 1, k);
    k -= h;
    if (k < L) return r1_base(p, i, (int)k);
    k -= L;
    if (k < 3) return "\n+\n"[k];
    k -= 3;
    if (k < L) return r1_qual(p, i, (int)k);
    return '\n';
}

SSD_HD char r2_byte(const Params& p, uint64_t i, const R2Plan& pl, uint32_t k)
{
    uint32_t h = header_len(i), L = (uint32_t)pl.len;
    if (k < h) return header_char(i, 2, k);
    k -= h;
    if (k < L) return r2_base(p, i, pl, (int)k);
    k -= L;
    if (k < 3) return "\n+\n"[k];
    k -= 3;
    if (k < L) return r2_qual(p, i, pl, (int)k);
    return '\n';
}

}  // namespace ssd_synth
