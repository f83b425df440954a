// libyachtgpu -- host-side entry points that sit next to the device path:
//   ygpu_read_signatures : multi-threaded signature ingest into a flat page-locked array
//                          (reference src/cpp/main.cpp:62-124 read_min_hashes / read_sketches)
#include "common.cuh"
#include "ingest.hpp"

#include <cstdio>
#include <cstring>

extern "C" int ygpu_read_signatures(const char* const* paths, uint32_t n, int threads, ygpu_sketch_set* out,
                                    char* errbuf, uint64_t errlen) {
    if (!out || (n && !paths)) return YGPU_ERR_ARG;
    memset(out, 0, sizeof(*out));
    yingest::Ingest in;
    in.quiet = true;
    in.names.reserve(n);
    for (uint32_t i = 0; i < n; i++) in.names.emplace_back(paths[i] ? paths[i] : "");
    yingest::read_sketches(in, threads < 1 ? 1 : threads);
    if (in.fatal) {
        if (errbuf && errlen) snprintf(errbuf, (size_t)errlen, "cannot parse signature %s", in.fatal_msg.c_str());
        if (in.hashes) { if (in.pinned) ygpu_host_free(in.hashes); else free(in.hashes); }
        return YGPU_ERR_ARG;
    }
    uint64_t* off = (uint64_t*)malloc(((size_t)n + 1) * sizeof(uint64_t));
    if (!off) {
        if (in.hashes) { if (in.pinned) ygpu_host_free(in.hashes); else free(in.hashes); }
        return YGPU_ERR_NOMEM;
    }
    memcpy(off, in.offsets.data(), ((size_t)n + 1) * sizeof(uint64_t));
    out->hashes = in.hashes;
    out->offsets = off;
    out->n_genomes = n;
    out->n_unreadable = (uint32_t)in.n_unreadable;
    out->pinned = in.pinned ? 1 : 0;
    return 0;
}

extern "C" void ygpu_sketch_set_free(ygpu_sketch_set* s) {
    if (!s) return;
    if (s->hashes) { if (s->pinned) ygpu_host_free(s->hashes); else free(s->hashes); }
    free(s->offsets);
    memset(s, 0, sizeof(*s));
}
