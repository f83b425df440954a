// FracMinHash k-mer hashing as sourmash publishes it ("hash_function": "0.murmur64", seed 42): the canonical k-mer (the
// lexicographically smaller of the window and its reverse complement) goes through MurmurHash3_x64_128 and the first
// 64 bits are the hash.  Bases are handled as 2-bit codes (A, C, G, T = 0..3 -- the ASCII order, so comparing codes is
// comparing the strings; 4 = anything else) and turned back into the ASCII bytes only where the hash consumes them.
//
// The functions are __host__ __device__ so that tests/ can compile exactly this code for the CPU and check it without a
// GPU (tests/harness/sketch_hash_host.cpp); the product only ever calls them from sketch.cu's kernel.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define YSK_HD __host__ __device__ __forceinline__
#else
#define YSK_HD inline
#endif

YSK_HD uint8_t ysk_code(uint8_t c) {                 // upper/lower case A C G T -> 0..3, anything else -> 4
    c &= 0xDFu;                                      // clears the case bit: 'a' -> 'A' (only c and c|0x20 map to the same value)
    return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 4;
}
YSK_HD uint64_t ysk_ascii(uint32_t code) { return (0x54474341u >> (8u * code)) & 0xFFu; }      // "ACGT"[code]
YSK_HD uint64_t ysk_rotl(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
YSK_HD uint64_t ysk_fmix(uint64_t k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}

// s: codes of a window of k valid bases.  Byte i of the canonical k-mer is s[i] (forward) or the complement of s[k-1-i].
YSK_HD uint64_t ysk_canonical_hash(const uint8_t* s, int k, uint32_t seed) {
    bool rc = false;
    for (int j = 0; j < k; j++) {                    // first difference between the window and its reverse complement
        const uint32_t a = s[j], b = 3u - s[k - 1 - j];
        if (a != b) { rc = a > b; break; }
    }
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    uint64_t h1 = seed, h2 = seed;
    const int nblocks = k >> 4;
    int i = 0;
    for (int blk = 0; blk < nblocks; blk++) {
        uint64_t k1 = 0, k2 = 0;
#pragma unroll
        for (int b = 0; b < 8; b++, i++) k1 |= ysk_ascii(rc ? 3u - s[k - 1 - i] : s[i]) << (8 * b);
#pragma unroll
        for (int b = 0; b < 8; b++, i++) k2 |= ysk_ascii(rc ? 3u - s[k - 1 - i] : s[i]) << (8 * b);
        k1 *= c1; k1 = ysk_rotl(k1, 31); k1 *= c2; h1 ^= k1;
        h1 = ysk_rotl(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        k2 *= c2; k2 = ysk_rotl(k2, 33); k2 *= c1; h2 ^= k2;
        h2 = ysk_rotl(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    const int t = k & 15;
    uint64_t k1 = 0, k2 = 0;
    for (int b = 0; b < t; b++, i++) {
        const uint64_t ch = ysk_ascii(rc ? 3u - s[k - 1 - i] : s[i]);
        if (b < 8) k1 |= ch << (8 * b);
        else k2 |= ch << (8 * (b - 8));
    }
    if (t > 8) { k2 *= c2; k2 = ysk_rotl(k2, 33); k2 *= c1; h2 ^= k2; }
    if (t > 0) { k1 *= c1; k1 = ysk_rotl(k1, 31); k1 *= c2; h1 ^= k1; }
    h1 ^= (uint64_t)k; h2 ^= (uint64_t)k;
    h1 += h2; h2 += h1;
    h1 = ysk_fmix(h1); h2 = ysk_fmix(h2);
    h1 += h2;
    return h1;
}
