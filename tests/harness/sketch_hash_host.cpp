// TEST HARNESS: compiles yacht_b200/csrc/sketch_hash.cuh -- the very functions the sketching kernel calls on the device --
// for the host, so that the CPU suite can hold them against oracle/sketch_oracle.c without a GPU.  Not part of the product.
#include <stdint.h>
#include <vector>
#include "../../yacht_b200/csrc/sketch_hash.cuh"

extern "C" uint64_t hh_hash_windows(const uint8_t* seq, uint64_t n, int k, uint32_t seed, uint64_t* out) {
    std::vector<uint8_t> code(n);
    for (uint64_t i = 0; i < n; i++) code[i] = ysk_code(seq[i]);
    uint64_t kept = 0;
    int run = 0;
    for (uint64_t e = 0; e < n; e++) {                 // e = last base of the window, like the kernel's rolling counter
        run = code[e] < 4 ? run + 1 : 0;
        if (run < k) continue;
        out[kept++] = ysk_canonical_hash(code.data() + (e + 1 - k), k, seed);
    }
    return kept;
}

// the k <= 32 variant: windows handed over as one word of 2-bit codes, the way the packed kernel extracts them
extern "C" uint64_t hh_hash_windows_packed(const uint8_t* seq, uint64_t n, int k, uint32_t seed, uint64_t* out) {
    std::vector<uint8_t> code(n);
    for (uint64_t i = 0; i < n; i++) code[i] = ysk_code(seq[i]);
    uint64_t kept = 0;
    int run = 0;
    for (uint64_t e = 0; e < n; e++) {
        run = code[e] < 4 ? run + 1 : 0;
        if (run < k) continue;
        uint64_t w = 0;
        for (int j = 0; j < k; j++) w |= (uint64_t)code[e + 1 - k + j] << (2 * j);
        // bits above 2k are garbage in the kernel (the neighbouring bases): make sure they are ignored
        if (k < 32) w |= 0xA5A5A5A5A5A5A5A5ULL << (2 * k);
        out[kept++] = ysk_canonical_hash_packed(w, k, seed);
    }
    return kept;
}

// The packed kernel's tile walk, thread by thread, with the very helpers the kernel uses (ysk_pack16, ysk_thread_span,
// ysk_span_window, ysk_canonical_hash_packed); only the grid/CTA loops and the shuffle that pairs the "bad" half-words are
// re-stated here.  Emits every valid window's hash in position order, like hh_hash_windows.
extern "C" uint64_t hh_tile_walk_packed(const uint8_t* seq, uint64_t n, int k, uint32_t seed, uint64_t* out) {
    const int TILE = 4096, NT = 256, PER = 16;
    std::vector<uint8_t> padded(n + TILE + 256 + 16, 0);          // the device buffer is zero-padded the same way
    for (uint64_t i = 0; i < n; i++) padded[i] = seq[i];
    uint64_t kept = 0;
    const uint64_t n_tiles = (n + TILE - 1) / TILE;
    for (uint64_t tile = 0; tile < n_tiles; tile++) {
        const uint64_t base = tile * TILE;
        const uint32_t* g = (const uint32_t*)(padded.data() + base);          // base is a multiple of 4096: aligned
        uint32_t s_code[TILE / 16 + 2], s_bad[TILE / 32 + 2] = {0};
        uint32_t bad16[NT + 2];
        for (int t = 0; t < NT + 2; t++) ysk_pack16(g[4 * t], g[4 * t + 1], g[4 * t + 2], g[4 * t + 3], s_code[t], bad16[t]);
        for (int t = 0; t < NT + 2; t += 2) s_bad[t >> 1] = bad16[t] | (bad16[t + 1] << 16);
        for (int tid = 0; tid < NT; tid++) {
            uint64_t lo, hi, badbits;
            ysk_thread_span(s_code, s_bad, tid, lo, hi, badbits);
            for (int i = 0; i < PER; i++) {
                const uint64_t p = base + (uint64_t)tid * PER + i;
                uint64_t w;
                if (!ysk_span_window(lo, hi, badbits, i, k, w) || p + (uint64_t)k > n) continue;
                out[kept++] = ysk_canonical_hash_packed(w, k, seed);
            }
        }
    }
    return kept;
}

// 33 <= k <= 64: two-word windows
extern "C" uint64_t hh_hash_windows_packed2(const uint8_t* seq, uint64_t n, int k, uint32_t seed, uint64_t* out) {
    std::vector<uint8_t> code(n);
    for (uint64_t i = 0; i < n; i++) code[i] = ysk_code(seq[i]);
    uint64_t kept = 0;
    int run = 0;
    for (uint64_t e = 0; e < n; e++) {
        run = code[e] < 4 ? run + 1 : 0;
        if (run < k) continue;
        uint64_t w0 = 0, w1 = 0;
        for (int j = 0; j < k; j++) {
            const uint64_t c = code[e + 1 - k + j];
            if (j < 32) w0 |= c << (2 * j); else w1 |= c << (2 * (j - 32));
        }
        if (k < 64) w1 |= 0x5A5A5A5A5A5A5A5AULL << (2 * k - 64);        // garbage above the window must be ignored
        out[kept++] = ysk_canonical_hash_packed2(w0, w1, k, seed);
    }
    return kept;
}

extern "C" uint64_t hh_tile_walk_packed2(const uint8_t* seq, uint64_t n, int k, uint32_t seed, uint64_t* out) {
    const int TILE = 4096, NT = 256, PER = 16;
    std::vector<uint8_t> padded(n + TILE + 256 + 16, 0);
    for (uint64_t i = 0; i < n; i++) padded[i] = seq[i];
    uint64_t kept = 0;
    const uint64_t n_tiles = (n + TILE - 1) / TILE;
    for (uint64_t tile = 0; tile < n_tiles; tile++) {
        const uint64_t base = tile * TILE;
        const uint32_t* g = (const uint32_t*)(padded.data() + base);
        uint32_t s_code[TILE / 16 + 4], s_bad[TILE / 32 + 4];
        uint32_t bad16[NT + 4];
        for (int t = 0; t < NT + 4; t++) ysk_pack16(g[4 * t], g[4 * t + 1], g[4 * t + 2], g[4 * t + 3], s_code[t], bad16[t]);
        for (int t = 0; t < NT + 4; t += 2) s_bad[t >> 1] = bad16[t] | (bad16[t + 1] << 16);
        s_bad[TILE / 32 + 2] = 0xDEADBEEFu;            // read by the last threads, never used: any value must do
        s_bad[TILE / 32 + 3] = 0xDEADBEEFu;
        for (int tid = 0; tid < NT; tid++) {
            uint64_t s0, s1, s2, b0, b1;
            ysk_thread_span2(s_code, s_bad, tid, s0, s1, s2, b0, b1);
            for (int i = 0; i < PER; i++) {
                const uint64_t p = base + (uint64_t)tid * PER + i;
                uint64_t w0, w1;
                if (!ysk_span_window2(s0, s1, s2, b0, b1, i, k, w0, w1) || p + (uint64_t)k > n) continue;
                out[kept++] = ysk_canonical_hash_packed2(w0, w1, k, seed);
            }
        }
    }
    return kept;
}
