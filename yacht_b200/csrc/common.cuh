// Shared declarations of libyachtgpu (internal; the public boundary is include/yacht_gpu.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <unordered_map>
#include "../../include/yacht_gpu.h"

struct ygpu_ctx {
    int device = 0;
    int num_sms = 148;
    int smem_optin = 0;          // max dynamic shared memory per CTA (opt-in), bytes
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {};
    std::string err;

    // ---- sketches (flat, device resident) ---------------------------------------------------
    uint32_t n = 0;              // genomes
    uint64_t T = 0;              // total hashes
    uint64_t* d_hashes = nullptr;   // [T]   sketch hashes, genome-major (as loaded)
    uint64_t* d_offsets = nullptr;  // [n+1]
    uint32_t* d_sizes = nullptr;    // [n]   sketch sizes
    uint32_t* d_gid = nullptr;      // [T]   genome id of every hash slot
    bool loaded = false;

    // ---- inverted index (K2 output) ---------------------------------------------------------
    bool indexed = false;
    bool sorted = false;            // d_skey / d_sgid valid
    uint64_t* d_skey = nullptr;     // [T]   hashes sorted ascending
    uint32_t* d_sgid = nullptr;     // [T]   genome id of each sorted hash (ascending inside a run)
    uint8_t* d_flag = nullptr;      // [T]   1 if the sorted slot belongs to a run of length >= 2
    uint32_t* d_cpos = nullptr;     // [T]   exclusive scan of d_flag (slot in the compacted postings)
    uint64_t P = 0;                 // postings kept
    uint32_t* d_post = nullptr;     // [P]   compacted posting array (genome ids), runs contiguous
    uint32_t* d_rem = nullptr;      // [P]   postings that follow slot c inside its run
    uint64_t n_items = 0;
    uint64_t* d_row_ptr = nullptr;  // [n+1] CSR over row_items
    uint64_t* d_row_items = nullptr;// [n_items] (first posting slot << 32) | count  -- per query genome
    uint64_t* d_row_work = nullptr; // [n]   increments row i performs (sum of counts)
    unsigned long long* d_row_cnt = nullptr;  // [n+1] build scratch
    ygpu_index_stats stats = {};

    // ---- scratch ------------------------------------------------------------------------------
    std::unordered_map<void*, size_t> caps;   // capacity (bytes) of every dev_alloc'ed buffer, keyed by member address
    void* d_temp = nullptr;         // CUB temp storage (grown on demand)
    size_t temp_bytes = 0;
    unsigned long long* d_scalars = nullptr;  // [16] device counters
    uint64_t* d_out_key = nullptr;  // pair compaction buffers
    uint32_t* d_out_cnt = nullptr;
    uint64_t* d_out_key2 = nullptr;
    uint32_t* d_out_cnt2 = nullptr;
    uint64_t out_cap = 0;
    ygpu_pair* d_pairs = nullptr;   // sorted flagged pairs of the last pairwise call
    uint64_t pairs_cap = 0;
    uint64_t n_pairs = 0;
    uint32_t force_tile_w = 0;      // test hook: cap the accumulator tile width (0 = automatic)

    void* run_scratch = nullptr;    // run-path buffers (run_kernels.cu)

    ygpu_timings tm = {};
};

int ygpu_fail(ygpu_ctx* ctx, int code, const char* fmt, ...);

#define YG_CUDA(ctx, call)                                                                        \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess)                                                                   \
            return ygpu_fail((ctx), YGPU_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call,     \
                             cudaGetErrorString(e__));                                            \
    } while (0)

#define YG_CHECK(expr)                 \
    do {                               \
        int rc__ = (expr);             \
        if (rc__ != 0) return rc__;    \
    } while (0)

int ygpu_temp_reserve(ygpu_ctx* ctx, size_t bytes);
int ygpu_sort_sketches(ygpu_ctx* ctx);   // K2a (yacht_gpu.cu)

// run path (run_kernels.cu)
void ygpu_run_release(ygpu_ctx* ctx);
