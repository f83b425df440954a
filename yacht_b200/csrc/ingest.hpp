// ingest.hpp -- file list -> flat (pinned) uint64 hash array + CSR offsets, on all host cores.
// Replaces read_sketches / read_sketches_one_chunk of the reference (src/cpp/main.cpp:89-124):
// same result (sketch i = mins of file i; unreadable file => message + empty sketch, :68-71),
// but the files are claimed dynamically in small blocks instead of one static chunk per thread,
// parsed by the purpose-built scanner of sig_scan.hpp, and assembled directly into one
// page-locked buffer that ygpu_load_sketches can DMA from.
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iostream>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#ifdef __linux__
#include <pthread.h>
#include <sched.h>
#endif

#include "../../include/yacht_gpu.h"
#include "sig_scan.hpp"

namespace yingest {

struct Ingest {
    std::vector<std::string> names;
    uint64_t* hashes = nullptr;      // pinned (ygpu_host_alloc) or malloc
    bool pinned = false;
    std::vector<uint64_t> offsets;   // n + 1
    std::vector<int> empty_ids;
    // when read_sketches(..., assemble_flat = false): the parsed blocks stay where the workers left them
    std::vector<std::vector<uint64_t>> blocks;     // block b = sketches of files [b * kFilesPerBlock, ...)
    std::vector<uint64_t> block_max;               // largest hash of block b
    std::vector<uint8_t> block_sorted;             // 1 if every sketch of block b is ascending (sourmash writes them so)
    // called by a parser thread as soon as block b is complete (blocks[b] is final from then on): lets the caller
    // ship blocks to the GPU while later files are still being parsed
    std::function<void(uint32_t)> on_block;
    bool fatal = false;
    bool quiet = false;              // library use: do not print the per-file message
    int ksize = 0;                   // > 0: the run side's rule (utils.py:31-51) -- exactly one sub-signature of this k-mer size per file,
                                     // searched in all records; 0: the train core's rule (main.cpp:78) -- [0]["signatures"][0]["mins"]
    std::atomic<uint32_t> n_unreadable{0};
    std::string fatal_msg;
};

constexpr uint32_t kFilesPerBlock = 32;

// Short-lived worker threads are not reliably spread over the allowed CPUs by the scheduler in
// containerised hosts (measured: 8 unpinned parser threads ran no faster than 1; pinned, 5x faster),
// so worker t pins itself to the t-th CPU of the process's affinity mask.
inline void pin_worker(int t) {
#ifdef __linux__
    cpu_set_t allowed;
    if (sched_getaffinity(0, sizeof allowed, &allowed) != 0) return;
    int cpus[CPU_SETSIZE], n = 0;
    for (int c = 0; c < CPU_SETSIZE; c++) if (CPU_ISSET(c, &allowed)) cpus[n++] = c;
    if (n <= 1) return;
    cpu_set_t one;
    CPU_ZERO(&one);
    CPU_SET(cpus[t % n], &one);
    pthread_setaffinity_np(pthread_self(), sizeof one, &one);
#else
    (void)t;
#endif
}

inline void read_sketches(Ingest& in, int threads, bool assemble_flat = true) {
    const uint32_t n = (uint32_t)in.names.size();
    const uint32_t nblocks = (n + kFilesPerBlock - 1) / kFilesPerBlock;
    in.blocks.assign(nblocks, std::vector<uint64_t>());
    in.block_max.assign(nblocks, 0);
    in.block_sorted.assign(nblocks, 1);
    std::vector<std::vector<uint64_t>>& block_hashes = in.blocks;
    std::vector<uint32_t> sizes(n, 0);
    std::atomic<uint32_t> next{0};
    std::mutex mu;
    auto worker = [&](int tid) {
        pin_worker(tid);
        std::vector<char> buf;
        for (;;) {
            const uint32_t b = next.fetch_add(1);
            if (b >= nblocks) break;
            // parse into a thread-local vector and hand it over once per block: the headers of
            // neighbouring block_hashes[] entries share cache lines, and push_back updates them
            std::vector<uint64_t> out;
            const uint32_t f0 = b * kFilesPerBlock, f1 = std::min(n, f0 + kFilesPerBlock);
            size_t guess = 0;
            for (uint32_t f = f0; f < f1; f++) guess += 6000;
            out.reserve(guess);
            for (uint32_t f = f0; f < f1; f++) {
                const size_t before = out.size();
                std::string why;
                int n_match = 1;
                sigscan::Status st = in.ksize > 0 ? sigscan::read_mins_ksize(in.names[f], in.ksize, buf, out, &why, &n_match)
                                                  : sigscan::read_mins(in.names[f], buf, out, &why);
                if (st == sigscan::OK && n_match != 1) {
                    st = sigscan::MALFORMED;
                    why = "Expected exactly one signature with ksize " + std::to_string(in.ksize) + ", found " + std::to_string(n_match);
                }
                if (st == sigscan::CANNOT_OPEN) {
                    if (!in.quiet) std::cerr << "Could not open the file!" << std::endl;  // main.cpp:69
                    in.n_unreadable.fetch_add(1);
                    out.resize(before);
                } else if (st == sigscan::MALFORMED) {
                    std::lock_guard<std::mutex> lk(mu);
                    if (!in.fatal) { in.fatal = true; in.fatal_msg = in.names[f] + ": " + why; }
                    out.resize(before);
                }
                sizes[f] = (uint32_t)(out.size() - before);
                // (while the sketch is still in cache) its largest hash and whether it is ascending: the multi-GPU host layer
                // cuts sorted sketches by hash range with two binary searches
                uint64_t mx = in.block_max[b];
                bool asc = true;
                for (size_t k = before; k < out.size(); k++) {
                    mx = std::max(mx, out[k]);
                    asc &= k == before || out[k - 1] <= out[k];
                }
                in.block_max[b] = mx;
                if (!asc) in.block_sorted[b] = 0;
            }
            block_hashes[b] = std::move(out);
            if (in.on_block) in.on_block(b);
        }
    };
    const auto t_begin = std::chrono::high_resolution_clock::now();
    std::vector<std::thread> pool;
    const int nt = std::max(1, std::min<int>(threads, (int)std::max<uint32_t>(nblocks, 1)));
    for (int t = 0; t < nt; t++) pool.emplace_back(worker, t);
    for (auto& t : pool) t.join();
    const auto t_parsed = std::chrono::high_resolution_clock::now();

    in.offsets.assign((size_t)n + 1, 0);
    for (uint32_t f = 0; f < n; f++) {
        in.offsets[f + 1] = in.offsets[f] + sizes[f];
        if (sizes[f] == 0) in.empty_ids.push_back((int)f);
    }
    const uint64_t T = in.offsets[n];
    if (!assemble_flat) {
        if (getenv("YACHT_INGEST_TIMING")) {
            auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
            std::cerr << "[ingest] " << n << " files, " << T << " hashes, " << nt << " threads: parse " << ms(t_begin, t_parsed)
                      << " ms (blocks kept in place)" << std::endl;
        }
        return;
    }
    in.hashes = (uint64_t*)ygpu_host_alloc(std::max<uint64_t>(T, 1) * sizeof(uint64_t));
    in.pinned = in.hashes != nullptr;
    if (!in.hashes) in.hashes = (uint64_t*)malloc(std::max<uint64_t>(T, 1) * sizeof(uint64_t));
    std::atomic<uint32_t> nextb{0};
    auto copier = [&](int tid) {
        pin_worker(tid);
        for (;;) {
            const uint32_t b = nextb.fetch_add(1);
            if (b >= nblocks) break;
            const uint64_t dst = in.offsets[(size_t)b * kFilesPerBlock];
            if (!block_hashes[b].empty())
                memcpy(in.hashes + dst, block_hashes[b].data(), block_hashes[b].size() * sizeof(uint64_t));
            std::vector<uint64_t>().swap(block_hashes[b]);
        }
    };
    const auto t_alloc = std::chrono::high_resolution_clock::now();
    pool.clear();
    for (int t = 0; t < nt; t++) pool.emplace_back(copier, t);
    for (auto& t : pool) t.join();
    in.blocks.clear();
    if (getenv("YACHT_INGEST_TIMING")) {
        const auto t_end = std::chrono::high_resolution_clock::now();
        auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        std::cerr << "[ingest] " << n << " files, " << T << " hashes, " << nt << " threads: parse " << ms(t_begin, t_parsed)
                  << " ms, staging alloc (" << (in.pinned ? "pinned" : "pageable") << ") " << ms(t_parsed, t_alloc)
                  << " ms, assemble " << ms(t_alloc, t_end) << " ms" << std::endl;
    }
}


}  // namespace yingest
