/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the FracMinHash sketching the reference delegates to sourmash.
 *
 * The reference's wrappers (src/yacht/sketch_ref_genomes.py:25,61; src/yacht/sketch_sample.py:32,49) shell out to
 *     sourmash sketch dna -p k=K,scaled=S,abund ...        (sourmash >= 4.8.3, < 5: env/yacht_env.yml:8)
 * sourmash is a third-party dependency that is absent from /root/reference and from this image, so its published
 * algorithm is restated here (sourmash docs "FracMinHash"; every signature the reference ships says
 * "hash_function": "0.murmur64", "seed": 42, "max_hash": 18446744073709552 for scaled = 1000):
 *   - per record, the sequence is upper-cased; every window of K characters made only of A, C, G, T is a k-mer, windows
 *     containing anything else are skipped (`sketch dna` adds sequences with force = True);
 *   - the canonical k-mer is the lexicographically smaller of the window and its reverse complement;
 *   - hash = first 64 bits of MurmurHash3_x64_128(canonical k-mer bytes, seed 42);
 *   - the hash is kept when hash <= max_hash, max_hash = round((2^64 - 1) / scaled) in double arithmetic
 *     (18446744073709552 at scaled = 1000); with `abund` the sketch counts how often each kept hash occurred.
 * PINNED (tests/test_sketch_oracle.py, tests/golden/make_sketch_golden.py):
 *   - MurmurHash3_x64_128 against the SMHasher verification value of the published algorithm (0x6384BA69);
 *   - end to end against the reference's own checked-in result workbook
 *     tests/testdata/standardize_output_testdata/results/result.xlsx, whose rows for five of the demo genomes
 *     (demo/ref_genomes/GCF_018918{045,095,125,185,235}.1, k = 31, scaled = 1000) record
 *     num_unique_kmers_in_genome_sketch / num_total_kmers_in_genome_sketch = 2452/2453, 3009/3035, 3495/3519, 2832/2844,
 *     2319/2323: this restatement reproduces all ten numbers.
 *   Hash VALUES have no golden vector in the reference tree (it ships no sequence together with its sketch); the pin is
 *   the ten counts (a different hash function, seed, canonical rule or max_hash changes them) plus the hash KAT.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use this file.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

static uint64_t fmix64(uint64_t k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}

/* MurmurHash3_x64_128 (Austin Appleby, public domain algorithm), both 64-bit halves. */
void so_murmur3_x64_128(const uint8_t* data, int len, uint32_t seed, uint64_t out[2]) {
    const int nblocks = len / 16;
    uint64_t h1 = seed, h2 = seed;
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    for (int i = 0; i < nblocks; i++) {
        uint64_t k1 = 0, k2 = 0;
        for (int b = 0; b < 8; b++) {            /* little-endian block words, byte by byte (no alignment assumptions) */
            k1 |= (uint64_t)data[16 * i + b] << (8 * b);
            k2 |= (uint64_t)data[16 * i + 8 + b] << (8 * b);
        }
        k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
        h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
        h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    const uint8_t* tail = data + nblocks * 16;
    const int t = len & 15;
    uint64_t k1 = 0, k2 = 0;
    for (int b = 8; b < t; b++) k2 |= (uint64_t)tail[b] << (8 * (b - 8));
    if (t > 8) { k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2; }
    for (int b = 0; b < (t < 8 ? t : 8); b++) k1 |= (uint64_t)tail[b] << (8 * b);
    if (t > 0) { k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1; }
    h1 ^= (uint64_t)len; h2 ^= (uint64_t)len;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2; h2 += h1;
    out[0] = h1; out[1] = h2;
}

/* SMHasher's VerificationTest for a 128-bit hash: keys {0}, {0,1}, ... of length 0..255 hashed with seed 256 - len, the
 * 256 results hashed again with seed 0; the first four bytes, little endian, are the verification value. */
uint32_t so_murmur3_verification(void) {
    uint8_t key[256], hashes[256 * 16], fin[16];
    uint64_t out[2];
    for (int i = 0; i < 256; i++) {
        key[i] = (uint8_t)i;
        so_murmur3_x64_128(key, i, (uint32_t)(256 - i), out);
        memcpy(hashes + 16 * i, out, 16);
    }
    so_murmur3_x64_128(hashes, 256 * 16, 0, out);
    memcpy(fin, out, 16);
    return (uint32_t)fin[0] | ((uint32_t)fin[1] << 8) | ((uint32_t)fin[2] << 16) | ((uint32_t)fin[3] << 24);
}

static int complement(int c) {
    switch (c) {
        case 'A': return 'T';
        case 'C': return 'G';
        case 'G': return 'C';
        case 'T': return 'A';
        default: return 0;
    }
}

/* Kept hashes of ONE record, in sequence order (duplicates kept: the caller counts abundances).
 * Returns the number of kept hashes; writes at most cap of them. */
uint64_t so_sketch_record(const uint8_t* seq, uint64_t n, int ksize, uint32_t seed, uint64_t max_hash, uint64_t* out, uint64_t cap) {
    if (ksize <= 0 || n < (uint64_t)ksize) return 0;
    uint8_t* fw = (uint8_t*)malloc((size_t)ksize);
    uint8_t* rc = (uint8_t*)malloc((size_t)ksize);
    uint64_t kept = 0;
    for (uint64_t p = 0; p + (uint64_t)ksize <= n; p++) {
        int ok = 1;
        for (int j = 0; j < ksize; j++) {
            int c = seq[p + j];
            if (c >= 'a' && c <= 'z') c -= 32;
            const int d = complement(c);
            if (!d) { ok = 0; break; }
            fw[j] = (uint8_t)c;
            rc[ksize - 1 - j] = (uint8_t)d;
        }
        if (!ok) continue;
        const uint8_t* canon = memcmp(fw, rc, (size_t)ksize) <= 0 ? fw : rc;
        uint64_t h[2];
        so_murmur3_x64_128(canon, ksize, seed, h);
        if (h[0] <= max_hash) {
            if (kept < cap) out[kept] = h[0];
            kept++;
        }
    }
    free(fw);
    free(rc);
    return kept;
}
