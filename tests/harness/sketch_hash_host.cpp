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
