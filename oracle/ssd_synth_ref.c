/*
 * ssd_synth_ref.c — TEST / BENCH INFRASTRUCTURE ONLY.
 *
 * The synthetic read pairs of SURVEY.md 8(d) ("synthetic record shape", "value distributions", "seeds / generator")
 * restated in plain C, so that the reference arm of bench.py and the oracle-side tests can make the very workload the
 * GPU arm measures without loading libssd_b200.so.  The product's own generator (csrc/ssd_synth.cuh, device kernel +
 * host twin) must produce the same bytes for the same spec; tests/test_oracle_golden.py compares the two.
 *
 *   read 1:  "@SYN.<i> <i>/1\n" + L uniform ACGT + "\n+\n" + L quality chars + "\n"
 *   read 2:  "@SYN.<i> <i>/2\n" + UMI(10) BC3(8) LINK(30) BC2(8) LINK(30) BC1(8) tail(L-94) + "\n+\n" + quality + "\n"
 *
 * Every random decision is a hash of (seed, pair index, field): any index range can be generated on its own.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct ssd_oracle_synth_spec {
    uint64_t seed, first_index, n_pairs;
    int32_t read_len, n_whitelist;
    const char *whitelist; /* n_whitelist x 8 */
    uint32_t cell_pool, p_sub_ppm, p_n_ppm, p_junk_ppm, p_dup_umi_ppm, p_tie_ppm, p_trunc_ppm, p_atqual_ppm;
} ssd_oracle_synth_spec;

enum {
    F_R1SEQ = 0x1000, F_R1Q = 0x2000, F_R2Q = 0x3000, F_TAIL = 0x4000, F_N = 0x5000, F_SUB = 0x6000,
    F_CELL = 1, F_CELL2 = 2, F_JUNK = 3, F_JUNKBC = 4, F_UMI = 5, F_DUP = 6, F_DUPSEL = 7, F_TIE = 8, F_TIESEL = 9,
    F_TRUNC = 10, F_TRUNCLEN = 11, F_ATQ1 = 12, F_ATQ2 = 13, F_POOL = 0x77
};

static const char LINK32[] = "GTGGCCGCTGTTTCGCATCGGCGTACGACT"; /* between BC3 and BC2 */
static const char LINK21[] = "ATCCACGTGCTTGAGAGGCCAGAGCATTCG"; /* between BC2 and BC1 */
static const char QUALS[] = "EEEEEEEEEAAA6/<<";

static uint64_t mix(uint64_t seed, uint64_t idx, uint64_t field)
{
    uint64_t z = seed + idx * 0x9E3779B97F4A7C15ull + field * 0xD1B54A32D192ED03ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static uint32_t ppm(uint64_t h) { return (uint32_t)(h % 1000000ull); }

static int digits10(uint64_t v)
{
    int d = 1;
    while (v >= 10) {
        v /= 10;
        d++;
    }
    return d;
}

static int r2_len(const ssd_oracle_synth_spec *s, uint64_t i)
{
    if (s->p_trunc_ppm && ppm(mix(s->seed, i, F_TRUNC)) < s->p_trunc_ppm) return 40 + (int)(mix(s->seed, i, F_TRUNCLEN) % 54ull);
    return s->read_len;
}
static uint32_t hdr_len(uint64_t i) { return (uint32_t)(9 + 2 * digits10(i)); }

static uint32_t code_of(const char *s8)
{
    uint32_t c = 0;
    int j;
    for (j = 0; j < 8; j++) c |= (s8[j] == 'A' ? 0u : s8[j] == 'C' ? 1u : s8[j] == 'G' ? 2u : 3u) << (2 * j);
    return c;
}

/* 8-mers halfway between two whitelist entries at Hamming distance exactly 4 (the first and third differing position
 * taken from the second entry): 2/2 ties for e = 2.  At most 512 of them, in pair order. */
static uint32_t tie_table(const ssd_oracle_synth_spec *s, char *ties)
{
    uint32_t n = 0;
    int a, b, j;
    for (a = 0; a < s->n_whitelist && n < 512; a++)
        for (b = a + 1; b < s->n_whitelist && n < 512; b++) {
            const char *x = s->whitelist + 8 * a, *y = s->whitelist + 8 * b;
            int diff[8], nd = 0;
            for (j = 0; j < 8; j++)
                if (x[j] != y[j]) diff[nd++] = j;
            if (nd != 4) continue;
            memcpy(ties + 8 * n, x, 8);
            ties[8 * n + diff[0]] = y[diff[0]];
            ties[8 * n + diff[2]] = y[diff[2]];
            n++;
        }
    return n;
}

static char *put_header(char *o, uint64_t i, int mate)
{
    char num[24];
    int d = 0, k;
    uint64_t v = i;
    do {
        num[d++] = (char)('0' + v % 10);
        v /= 10;
    } while (v);
    memcpy(o, "@SYN.", 5);
    o += 5;
    for (k = d - 1; k >= 0; k--) *o++ = num[k];
    *o++ = ' ';
    for (k = d - 1; k >= 0; k--) *o++ = num[k];
    *o++ = '/';
    *o++ = (char)('0' + mate);
    *o++ = '\n';
    return o;
}

static void one_pair(const ssd_oracle_synth_spec *s, uint32_t cell_octaves, const char *ties, uint32_t n_ties, uint64_t i, char *o1, char *o2)
{
    const int L = s->read_len, L2 = r2_len(s, i);
    uint32_t bc[3], umi;
    uint64_t rank = 0, h;
    int j, w;
    const int junk = ppm(mix(s->seed, i, F_JUNK)) < s->p_junk_ppm;
    char *seq2, *q;

    /* ---- read 1 ---- */
    o1 = put_header(o1, i, 1);
    for (j = 0; j < L; j++) o1[j] = "ACGT"[(mix(s->seed, i, F_R1SEQ + (uint64_t)(j >> 5)) >> (2 * (j & 31))) & 3u];
    o1 += L;
    memcpy(o1, "\n+\n", 3);
    o1 += 3;
    for (j = 0; j < L; j++) o1[j] = QUALS[(mix(s->seed, i, F_R1Q + (uint64_t)(j >> 4)) >> (4 * (j & 15))) & 15u];
    if (L > 0 && ppm(mix(s->seed, i, F_ATQ1)) < s->p_atqual_ppm) o1[0] = '@';
    o1[L] = '\n';

    /* ---- read 2: planted cell (octave distribution over the pool) or junk barcodes, substitutions, UMI ---- */
    if (junk) {
        h = mix(s->seed, i, F_JUNKBC);
        bc[0] = (uint32_t)(h & 0xFFFF);
        bc[1] = (uint32_t)((h >> 16) & 0xFFFF);
        bc[2] = (uint32_t)((h >> 32) & 0xFFFF);
    } else {
        const uint32_t oct = (uint32_t)(mix(s->seed, i, F_CELL) % cell_octaves);
        rank = ((1ull << oct) - 1 + (mix(s->seed, i, F_CELL2) & ((1ull << oct) - 1))) % s->cell_pool;
        h = mix(s->seed ^ 0xC3115EEDull, rank, F_POOL);
        bc[0] = code_of(s->whitelist + 8 * (h % (uint64_t)s->n_whitelist));
        bc[1] = code_of(s->whitelist + 8 * ((h >> 20) % (uint64_t)s->n_whitelist));
        bc[2] = code_of(s->whitelist + 8 * ((h >> 40) % (uint64_t)s->n_whitelist));
        if (n_ties && ppm(mix(s->seed, i, F_TIE)) < s->p_tie_ppm) bc[1] = code_of(ties + 8 * (mix(s->seed, i, F_TIESEL) % n_ties));
    }
    for (w = 0; w < 3; w++)
        for (j = 0; j < 8; j++) {
            h = mix(s->seed, i, F_SUB + (uint64_t)(w * 8 + j));
            if (ppm(h) < s->p_sub_ppm) {
                const uint32_t old = (bc[w] >> (2 * j)) & 3u, nw = (old + 1u + (uint32_t)((h >> 32) % 3ull)) & 3u;
                bc[w] = (bc[w] & ~(3u << (2 * j))) | (nw << (2 * j));
            }
        }
    if (!junk && ppm(mix(s->seed, i, F_DUP)) < s->p_dup_umi_ppm)
        umi = (uint32_t)(mix(s->seed ^ 0xD0Bull, rank, F_DUPSEL + (mix(s->seed, i, F_DUPSEL) & 3ull)) & 0xFFFFF);
    else
        umi = (uint32_t)(mix(s->seed, i, F_UMI) & 0xFFFFF);

    o2 = put_header(o2, i, 2);
    seq2 = o2;
    for (j = 0; j < L2; j++) {
        char c;
        if (j < 10) c = "ACGT"[(umi >> (2 * j)) & 3u];
        else if (j < 18) c = "ACGT"[(bc[2] >> (2 * (j - 10))) & 3u];
        else if (j < 48) c = LINK32[j - 18];
        else if (j < 56) c = "ACGT"[(bc[1] >> (2 * (j - 48))) & 3u];
        else if (j < 86) c = LINK21[j - 56];
        else if (j < 94) c = "ACGT"[(bc[0] >> (2 * (j - 86))) & 3u];
        else c = "ACGT"[(mix(s->seed, i, F_TAIL + (uint64_t)(j >> 5)) >> (2 * (j & 31))) & 3u];
        if (s->p_n_ppm && ppm(mix(s->seed, i, F_N + (uint64_t)j)) < s->p_n_ppm) c = 'N';
        seq2[j] = c;
    }
    o2 += L2;
    memcpy(o2, "\n+\n", 3);
    o2 += 3;
    q = o2;
    for (j = 0; j < L2; j++) q[j] = seq2[j] == 'N' ? '#' : QUALS[(mix(s->seed, i, F_R2Q + (uint64_t)(j >> 4)) >> (4 * (j & 15))) & 15u];
    if (L2 > 0 && ppm(mix(s->seed, i, F_ATQ2)) < s->p_atqual_ppm) q[0] = '@';
    q[L2] = '\n';
}

static int spec_ok(const ssd_oracle_synth_spec *s)
{
    return s && s->whitelist && s->n_whitelist >= 1 && s->n_whitelist <= 128 && s->read_len >= 94 && s->read_len <= 1000 && s->cell_pool >= 1;
}

int ssd_oracle_synth_sizes(const ssd_oracle_synth_spec *s, uint64_t *r1_bytes, uint64_t *r2_bytes)
{
    uint64_t a = 0, b = 0, k;
    if (!spec_ok(s)) return -1;
    for (k = 0; k < s->n_pairs; k++) {
        const uint64_t i = s->first_index + k;
        a += hdr_len(i) + 2u * (uint32_t)s->read_len + 4u;
        b += hdr_len(i) + 2u * (uint32_t)r2_len(s, i) + 4u;
    }
    if (r1_bytes) *r1_bytes = a;
    if (r2_bytes) *r2_bytes = b;
    return 0;
}

/* r1 / r2: buffers of the sizes ssd_oracle_synth_sizes reports.  n_threads <= 1: serial. */
int ssd_oracle_synth_generate(const ssd_oracle_synth_spec *s, char *r1, char *r2, int n_threads)
{
    enum { BLOCK = 4096 };
    char *ties;
    uint32_t n_ties, cell_octaves = 1;
    uint64_t n_blocks, b, *off1, *off2, a1 = 0, a2 = 0;
    if (!spec_ok(s) || !r1 || !r2) return -1;
    ties = (char *)malloc(8 * 512);
    n_blocks = (s->n_pairs + BLOCK - 1) / BLOCK;
    off1 = (uint64_t *)malloc((size_t)(n_blocks + 1) * sizeof(uint64_t));
    off2 = (uint64_t *)malloc((size_t)(n_blocks + 1) * sizeof(uint64_t));
    if (!ties || !off1 || !off2) {
        free(ties);
        free(off1);
        free(off2);
        return -4;
    }
    n_ties = tie_table(s, ties);
    while ((1ull << cell_octaves) <= s->cell_pool) cell_octaves++;
    if (n_threads < 1) n_threads = 1;
    /* where every block of pairs starts in the two files */
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(n_threads)
#endif
    for (b = 0; b < n_blocks; b++) {
        uint64_t k, hi = (b + 1) * BLOCK < s->n_pairs ? (b + 1) * BLOCK : s->n_pairs, x = 0, y = 0;
        for (k = b * BLOCK; k < hi; k++) {
            const uint64_t i = s->first_index + k;
            x += hdr_len(i) + 2u * (uint32_t)s->read_len + 4u;
            y += hdr_len(i) + 2u * (uint32_t)r2_len(s, i) + 4u;
        }
        off1[b + 1] = x;
        off2[b + 1] = y;
    }
    off1[0] = off2[0] = 0;
    for (b = 0; b < n_blocks; b++) {
        a1 += off1[b + 1];
        a2 += off2[b + 1];
        off1[b + 1] = a1;
        off2[b + 1] = a2;
    }
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4) num_threads(n_threads)
#endif
    for (b = 0; b < n_blocks; b++) {
        uint64_t k, hi = (b + 1) * BLOCK < s->n_pairs ? (b + 1) * BLOCK : s->n_pairs;
        char *o1 = r1 + off1[b], *o2 = r2 + off2[b];
        for (k = b * BLOCK; k < hi; k++) {
            const uint64_t i = s->first_index + k;
            one_pair(s, cell_octaves, ties, n_ties, i, o1, o2);
            o1 += hdr_len(i) + 2u * (uint32_t)s->read_len + 4u;
            o2 += hdr_len(i) + 2u * (uint32_t)r2_len(s, i) + 4u;
        }
    }
    free(ties);
    free(off1);
    free(off2);
    return 0;
}
