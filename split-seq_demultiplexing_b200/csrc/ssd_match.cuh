// ssd_match.cuh — the first-hit rule of the barcode match (V2:72-73, 302-304, 327-329) on 2-bit packed windows, as pure functions
// that compile for the device and for the host: tests/test_match_logic.py builds them with g++ and checks them, window by
// window, against the byte-wise rule (zip() truncation for short windows, bytes outside ACGT mismatching every base).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SSD_HD __host__ __device__ __forceinline__
#else
#define SSD_HD inline
#endif

namespace ssd {

SSD_HD uint32_t hd_popc(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return (uint32_t)__popc(x);
#else
    return (uint32_t)__builtin_popcount(x);
#endif
}

// byte j of the result = byte (nibble j of sel) of a, for selectors 0..3 (all this file needs of PRMT)
SSD_HD uint32_t hd_byte_perm_a(uint32_t a, uint32_t sel)
{
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, 0u, sel);
#else
    uint32_t r = 0;
    for (int j = 0; j < 4; j++) r |= ((a >> (8u * ((sel >> (4 * j)) & 3u))) & 0xFFu) << (8 * j);
    return r;
#endif
}

// four bases -> 8-bit 2-bit code (position j at bits 2j; A=0 C=1 T=2 G=3 = (c >> 1) & 3); `diff` = x ^ (the bases those codes
// stand for): byte j of it is zero exactly when byte j of x is one of A C G T
SSD_HD uint32_t pack4(uint32_t x, uint32_t& diff)
{
    const uint32_t y = (x >> 1) & 0x03030303u;
    const uint32_t sel = (y & 0x3u) | ((y >> 4) & 0x30u) | ((y >> 8) & 0x300u) | ((y >> 12) & 0x3000u);
    diff = x ^ hd_byte_perm_a(0x47544341u, sel);  // A C T G by code
    return (y * 0x01041040u) >> 24;
}

// bit 2j (j = 0..7) set for the window positions whose byte is not A/C/G/T; d0 / d1 = pack4's diff words of bytes 0-3 / 4-7
SSD_HD uint32_t invalid_positions(uint32_t d0, uint32_t d1)
{
    uint32_t m = 0;
#pragma unroll
    for (uint32_t j = 0; j < 4; j++) {
        if ((d0 >> (8u * j)) & 0xFFu) m |= 1u << (2u * j);
        if ((d1 >> (8u * j)) & 0xFFu) m |= 1u << (2u * (j + 4u));
    }
    return m;
}

// A window of which only the first n <= 8 bytes exist (lineRead[s:e] of a short read, V2:296-301): valid = the positions that
// hold A/C/G/T (bit 2j), k = how many of the n do not.  Positions >= n count for neither: hamming() zips to the shorter operand.
SSD_HD void window_masks(uint32_t d0, uint32_t d1, uint32_t n, uint32_t& valid, uint32_t& k)
{
    const uint32_t lenmask = n >= 8u ? 0x5555u : (((1u << (2u * n)) - 1u) & 0x5555u);
    const uint32_t inv = invalid_positions(d0, d1) & lenmask;
    valid = lenmask & ~inv;
    k = hd_popc(inv);
}

// Is the whitelist entry with 2-bit code wl_code within e of the window?  (a byte outside ACGT mismatches every base)
SSD_HD bool masked_hit(uint32_t code, uint32_t wl_code, uint32_t valid, uint32_t k, uint32_t e)
{
    const uint32_t x = code ^ wl_code;
    return k + hd_popc((x | (x >> 1)) & valid) <= e;
}

}  // namespace ssd
