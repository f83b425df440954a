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

// ---- k <= 32: the window as one 64-bit word of 2-bit codes (base j in bits [2j+1 : 2j]) -----------------------------------
YSK_HD uint64_t ysk_brev64(uint64_t x) {              // reverse all 64 bits
#ifdef __CUDA_ARCH__
    return __brevll(x);
#else
    x = ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
    x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
    x = ((x >> 8) & 0x00FF00FF00FF00FFULL) | ((x & 0x00FF00FF00FF00FFULL) << 8);
    x = ((x >> 16) & 0x0000FFFF0000FFFFULL) | ((x & 0x0000FFFF0000FFFFULL) << 16);
    return (x >> 32) | (x << 32);
#endif
}
YSK_HD uint32_t ysk_ascii4(uint32_t codes) {           // low 8 bits = 4 bases -> their 4 ASCII bytes, base 0 in the low byte
    uint32_t s = codes & 0xFFu;
    s = (s | (s << 4)) & 0x0F0Fu;
    s = (s | (s << 2)) & 0x3333u;                      // nibble i = code of base i: a byte selector
#ifdef __CUDA_ARCH__
    return __byte_perm(0x54474341u, 0u, s);            // PRMT: byte i = "ACGT"[nibble i]
#else
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= ((0x54474341u >> (8u * ((s >> (4 * i)) & 3u))) & 0xFFu) << (8 * i);
    return r;
#endif
}
YSK_HD uint64_t ysk_ascii8(uint64_t c, int q) {        // bytes 8q .. 8q+7 of the k-mer held in c
    return (uint64_t)ysk_ascii4((uint32_t)(c >> (16 * q))) | ((uint64_t)ysk_ascii4((uint32_t)(c >> (16 * q + 8))) << 32);
}
YSK_HD uint64_t ysk_low_bytes(int n) { return n >= 8 ? ~0ULL : ((1ULL << (8 * n)) - 1ULL); }

// Same result as ysk_canonical_hash for a window of k <= 32 valid bases given as packed codes.
YSK_HD uint64_t ysk_canonical_hash_packed(uint64_t w, int k, uint32_t seed) {
    const uint64_t mask = (k >= 32) ? ~0ULL : ((1ULL << (2 * k)) - 1ULL);
    const int sh = 64 - 2 * k;
    w &= mask;
    uint64_t fwm = ysk_brev64(w);                                                       // bits reversed: code j in group 31 - j, its two bits swapped
    fwm = ((fwm & 0x5555555555555555ULL) << 1) | ((fwm >> 1) & 0x5555555555555555ULL);  // base 0 in the top two bits: integer order = string order
    const uint64_t rcm = (~w & mask) << sh;                                             // the reverse complement, packed the same way
    const uint64_t rcl = ~(fwm >> sh) & mask;                                           // ... and with its base 0 in the low bits
    const uint64_t c = (fwm <= rcm) ? w : rcl;                                          // canonical k-mer, byte i = base i
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    uint64_t h1 = seed, h2 = seed;
    const int nblocks = k >> 4;
    for (int blk = 0; blk < nblocks; blk++) {
        uint64_t k1 = ysk_ascii8(c, 2 * blk), k2 = ysk_ascii8(c, 2 * blk + 1);
        k1 *= c1; k1 = ysk_rotl(k1, 31); k1 *= c2; h1 ^= k1;
        h1 = ysk_rotl(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        k2 *= c2; k2 = ysk_rotl(k2, 33); k2 *= c1; h2 ^= k2;
        h2 = ysk_rotl(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    const int t = k & 15;
    if (t > 8) {
        uint64_t k2 = ysk_ascii8(c, 2 * nblocks + 1) & ysk_low_bytes(t - 8);
        k2 *= c2; k2 = ysk_rotl(k2, 33); k2 *= c1; h2 ^= k2;
    }
    if (t > 0) {
        uint64_t k1 = ysk_ascii8(c, 2 * nblocks) & ysk_low_bytes(t);
        k1 *= c1; k1 = ysk_rotl(k1, 31); k1 *= c2; h1 ^= k1;
    }
    h1 ^= (uint64_t)k; h2 ^= (uint64_t)k;
    h1 += h2; h2 += h1;
    h1 = ysk_fmix(h1); h2 = ysk_fmix(h2);
    h1 += h2;
    return h1;
}

// ---- tile layout of the packed kernel: 16 bases per code word, one "not A/C/G/T" bit per base ------------------------------
YSK_HD void ysk_pack16(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t& code, uint32_t& bad) {   // 16 sequence bytes
    const uint32_t x[4] = {x0, x1, x2, x3};
    code = 0; bad = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int q = 0; q < 4; q++)
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int b = 0; b < 4; b++) {
            const uint32_t c = ysk_code((uint8_t)(x[q] >> (8 * b)));
            code |= (c & 3u) << (2 * (4 * q + b));
            bad |= (c >> 2) << (4 * q + b);
        }
}
// What thread `tid` (16 window starts from base 16 * tid of the tile) needs: the codes of 48 bases and their "bad" bits.
YSK_HD void ysk_thread_span(const uint32_t* s_code, const uint32_t* s_bad, int tid, uint64_t& lo, uint64_t& hi, uint64_t& badbits) {
    lo = (uint64_t)s_code[tid] | ((uint64_t)s_code[tid + 1] << 32);                     // bases 16 tid .. 16 tid + 31
    hi = (uint64_t)s_code[tid + 2];                                                      // bases 16 tid + 32 .. 16 tid + 47
    badbits = ((uint64_t)s_bad[tid >> 1] | ((uint64_t)s_bad[(tid >> 1) + 1] << 32)) >> ((tid & 1) * 16);    // >= 48 bits from base 16 tid on
}
// Window i (0..15) of the span: false when it holds a base that is not A/C/G/T; else w = its codes (garbage above bit 2k).
YSK_HD bool ysk_span_window(uint64_t lo, uint64_t hi, uint64_t badbits, int i, int k, uint64_t& w) {
    const uint64_t kmask = (k >= 64) ? ~0ULL : ((1ULL << k) - 1ULL);
    if (((badbits >> i) & kmask) != 0) return false;
    w = i ? ((lo >> (2 * i)) | (hi << (64 - 2 * i))) : lo;
    return true;
}

// ---- 33 <= k <= 64: the window as two 64-bit words (w1:w0), base j in bits [2j+1 : 2j] of the 128-bit value ----------------
YSK_HD uint64_t ysk_rev2(uint64_t x) {                 // reverse the order of the 32 two-bit groups
    x = ysk_brev64(x);
    return ((x & 0x5555555555555555ULL) << 1) | ((x >> 1) & 0x5555555555555555ULL);
}
YSK_HD uint64_t ysk_ascii8w(uint64_t c0, uint64_t c1, int q) {      // bytes 8q .. 8q+7 of the k-mer held in (c1:c0), q = 0..7
    return q < 4 ? ysk_ascii8(c0, q) : ysk_ascii8(c1, q - 4);
}
YSK_HD uint64_t ysk_canonical_hash_packed2(uint64_t w0, uint64_t w1, int k, uint32_t seed) {
    const uint64_t mask1 = (k >= 64) ? ~0ULL : ((1ULL << (2 * k - 64)) - 1ULL);          // k >= 33: the low word is full
    const int sh = 128 - 2 * k;                                                         // 0 .. 62
    w1 &= mask1;
    const uint64_t f1 = ysk_rev2(w0), f0 = ysk_rev2(w1);                                // (f1:f0): base 0 in the top two bits, zero fill below
    const uint64_t n0 = ~w0, n1 = ~w1 & mask1;                                          // complement of every base
    const uint64_t r1 = sh ? ((n1 << sh) | (n0 >> (64 - sh))) : n1, r0 = n0 << sh;      // (r1:r0) = reverse complement, packed like (f1:f0)
    const bool fw = (f1 < r1) || (f1 == r1 && f0 <= r0);
    const uint64_t g0 = sh ? ((f0 >> sh) | (f1 << (64 - sh))) : f0, g1 = f1 >> sh;      // (f1:f0) >> sh: the window reversed, base 0 low
    const uint64_t c0 = fw ? w0 : ~g0, c1 = fw ? w1 : (~g1 & mask1);                    // canonical k-mer, byte i = base i
    const uint64_t m1 = 0x87c37b91114253d5ULL, m2 = 0x4cf5ad432745937fULL;
    uint64_t h1 = seed, h2 = seed;
    const int nblocks = k >> 4;
    for (int blk = 0; blk < nblocks; blk++) {
        uint64_t k1 = ysk_ascii8w(c0, c1, 2 * blk), k2 = ysk_ascii8w(c0, c1, 2 * blk + 1);
        k1 *= m1; k1 = ysk_rotl(k1, 31); k1 *= m2; h1 ^= k1;
        h1 = ysk_rotl(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        k2 *= m2; k2 = ysk_rotl(k2, 33); k2 *= m1; h2 ^= k2;
        h2 = ysk_rotl(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    const int t = k & 15;
    if (t > 8) {
        uint64_t k2 = ysk_ascii8w(c0, c1, 2 * nblocks + 1) & ysk_low_bytes(t - 8);
        k2 *= m2; k2 = ysk_rotl(k2, 33); k2 *= m1; h2 ^= k2;
    }
    if (t > 0) {
        uint64_t k1 = ysk_ascii8w(c0, c1, 2 * nblocks) & ysk_low_bytes(t);
        k1 *= m1; k1 = ysk_rotl(k1, 31); k1 *= m2; h1 ^= k1;
    }
    h1 ^= (uint64_t)k; h2 ^= (uint64_t)k;
    h1 += h2; h2 += h1;
    h1 = ysk_fmix(h1); h2 = ysk_fmix(h2);
    h1 += h2;
    return h1;
}
// Thread `tid`'s 80-base span (16 window starts, up to 63 bases behind the last one) and its "bad" bits.
YSK_HD void ysk_thread_span2(const uint32_t* s_code, const uint32_t* s_bad, int tid, uint64_t& s0, uint64_t& s1, uint64_t& s2,
                             uint64_t& b0, uint64_t& b1) {
    s0 = (uint64_t)s_code[tid] | ((uint64_t)s_code[tid + 1] << 32);
    s1 = (uint64_t)s_code[tid + 2] | ((uint64_t)s_code[tid + 3] << 32);
    s2 = (uint64_t)s_code[tid + 4];
    const int idx = tid >> 1, sh = (tid & 1) * 16;
    const uint64_t raw0 = (uint64_t)s_bad[idx] | ((uint64_t)s_bad[idx + 1] << 32);
    const uint64_t raw1 = (uint64_t)s_bad[idx + 2] | ((uint64_t)s_bad[idx + 3] << 32);
    b0 = sh ? ((raw0 >> 16) | (raw1 << 48)) : raw0;        // bases 16 tid .. 16 tid + 63
    b1 = raw1 >> sh;                                        // bases 16 tid + 64 .. (at least 48 of them; 15 are needed)
}
YSK_HD bool ysk_span_window2(uint64_t s0, uint64_t s1, uint64_t s2, uint64_t b0, uint64_t b1, int i, int k, uint64_t& w0, uint64_t& w1) {
    const uint64_t kmask = (k >= 64) ? ~0ULL : ((1ULL << k) - 1ULL);
    const uint64_t bad = i ? ((b0 >> i) | (b1 << (64 - i))) : b0;       // "bad" bits of bases i .. i + 63 of the span
    if ((bad & kmask) != 0) return false;
    w0 = i ? ((s0 >> (2 * i)) | (s1 << (64 - 2 * i))) : s0;
    w1 = i ? ((s1 >> (2 * i)) | (s2 << (64 - 2 * i))) : s1;
    return true;
}
