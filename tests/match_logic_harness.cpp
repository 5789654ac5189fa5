// Host build of csrc/ssd_match.cuh for tests/test_match_logic.py (g++): the packed, masked first-hit rule the read-2 scan runs
// on windows the lookup tables do not answer, executed the way the warp does it (32 lanes, 32 whitelist entries per round of
// votes, lowest hit wins), next to the byte-wise rule of the reference (V2:72-73, 302-304) for a self-check in bulk.
#include <stdint.h>
#include <string.h>

#include "ssd_match.cuh"

using namespace ssd;

static uint32_t word_at(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

static uint16_t code_of(const uint8_t* w8)
{
    uint32_t c = 0;
    for (int j = 0; j < 8; j++) c |= ((uint32_t)(w8[j] >> 1) & 3u) << (2 * j);  // ssd_create's wl_code
    return (uint16_t)c;
}

extern "C" {

// window: at least 8 readable bytes, of which the first n exist in the read; wl: n_wl x 8 bytes (A/C/G/T only)
int ml_first_hit_packed(const uint8_t* window, int n, int e, const uint8_t* wl, int n_wl)
{
    uint32_t d0, d1, valid, k;
    const uint32_t code = pack4(word_at(window), d0) | (pack4(word_at(window + 4), d1) << 8);
    window_masks(d0, d1, (uint32_t)n, valid, k);
    if (n == 0) {
        valid = 0;  // the scan does not even load an empty window
        k = 0;
    }
    if (k > (uint32_t)e) return -1;
    for (int base = 0; base < n_wl; base += 32) {  // one round of votes
        uint32_t ballot = 0;
        for (int lane = 0; lane < 32; lane++) {
            const int i = base + lane;
            if (i < n_wl && masked_hit(code, code_of(wl + 8 * i), valid, k, (uint32_t)e)) ballot |= 1u << lane;
        }
        if (ballot) return base + __builtin_ctz(ballot);
    }
    return -1;
}

// the reference's rule on the bytes: first entry with sum(c1 != c2 for c1, c2 in zip(window[:n], entry)) <= e
int ml_first_hit_bytes(const uint8_t* window, int n, int e, const uint8_t* wl, int n_wl)
{
    for (int i = 0; i < n_wl; i++) {
        int d = 0;
        for (int j = 0; j < n && j < 8; j++) d += window[j] != wl[8 * i + j];
        if (d <= e) return i;
    }
    return -1;
}

// pack4 on four bytes: returns code | (bitmask of the bytes that are not A/C/G/T) << 8
int ml_pack4(const uint8_t* four)
{
    uint32_t diff;
    const uint32_t code = pack4(word_at(four), diff);
    uint32_t bad = 0;
    for (int j = 0; j < 4; j++)
        if ((diff >> (8 * j)) & 0xFFu) bad |= 1u << j;
    return (int)(code | (bad << 8));
}

// bulk self-check: n_cases random windows (xorshift from seed; alphabet weighted towards bases, with N, lower case, line
// ends, quality characters, bytes >= 0x80), every length 0..8, every e in 0..3; returns the number of disagreements
long ml_bulk_check(uint64_t seed, long n_cases, const uint8_t* wl, int n_wl)
{
    static const uint8_t alphabet[] = {'A', 'C', 'G', 'T', 'A', 'C', 'G', 'T', 'A', 'C', 'G', 'T', 'N', 'n', 'a', 'c', 'g', 't',
                                       '\n', '\r', '+', 'E', '6', '/', '<', '#', '@', ' ', 0, 0x80, 0xC1, 0xFF, 'B', 'U'};
    uint64_t s = seed ? seed : 1;
    long bad = 0;
    for (long t = 0; t < n_cases; t++) {
        uint8_t w[12];
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        // half of the windows start from a whitelist entry (so that hits exist), then get random substitutions
        const bool planted = (s >> 60) & 1;
        const uint8_t* src = wl + 8 * ((s >> 32) % (uint64_t)n_wl);
        uint64_t r = s;
        for (int j = 0; j < 12; j++) {
            r ^= r << 13; r ^= r >> 7; r ^= r << 17;
            const bool subst = !planted || j >= 8 || (r & 7) == 0;
            w[j] = subst ? alphabet[(r >> 8) % sizeof alphabet] : src[j];
        }
        for (int n = 0; n <= 8; n++)
            for (int e = 0; e <= 3; e++)
                if (ml_first_hit_packed(w, n, e, wl, n_wl) != ml_first_hit_bytes(w, n, e, wl, n_wl)) bad++;
    }
    return bad;
}

}  // extern "C"
