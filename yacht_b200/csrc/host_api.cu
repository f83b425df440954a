// libyachtgpu -- host-side entry points that sit next to the device path:
//   ygpu_read_signatures : multi-threaded signature ingest into a flat page-locked array
//                          (reference src/cpp/main.cpp:62-124 read_min_hashes / read_sketches)
//   ygpu_greedy_select   : the greedy near-duplicate removal over the flagged pairs
//                          (reference src/cpp/main.cpp:371-420 do_yacht_train; inherently sequential, milliseconds)
#include "common.cuh"
#include "ingest.hpp"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <utility>
#include <vector>

static int read_signatures_impl(const char* const* paths, uint32_t n, int threads, int ksize, ygpu_sketch_set* out, char* errbuf, uint64_t errlen);

extern "C" int ygpu_read_signatures(const char* const* paths, uint32_t n, int threads, ygpu_sketch_set* out,
                                    char* errbuf, uint64_t errlen) {
    return read_signatures_impl(paths, n, threads, 0, out, errbuf, errlen);
}

extern "C" int ygpu_read_signatures_ksize(const char* const* paths, uint32_t n, int threads, int ksize, ygpu_sketch_set* out,
                                          char* errbuf, uint64_t errlen) {
    if (ksize < 1) return YGPU_ERR_ARG;
    return read_signatures_impl(paths, n, threads, ksize, out, errbuf, errlen);
}

static int read_signatures_impl(const char* const* paths, uint32_t n, int threads, int ksize, ygpu_sketch_set* out, char* errbuf, uint64_t errlen) {
    if (!out || (n && !paths)) return YGPU_ERR_ARG;
    memset(out, 0, sizeof(*out));
    yingest::Ingest in;
    in.quiet = true;
    in.ksize = ksize;
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

// do_yacht_train (main.cpp:371-420): genomes are visited by ascending sketch size -- the SAME std::sort call on the same
// element type, initial order (genome id) and comparator as the reference, so genomes of equal size are visited in the
// same order (README: ties are "randomly selected"; here they are the reference's).  A genome is dropped iff one of the
// genomes it is contained in (pairs (i, j): i in j above the threshold) is still kept and not smaller.
extern "C" int ygpu_greedy_select(const uint64_t* offsets, uint32_t n, const ygpu_pair* pairs, uint64_t n_pairs, int32_t* selected,
                                  uint32_t* n_selected) {
    if (!offsets || (n_pairs && !pairs) || (n && !selected) || !n_selected) return YGPU_ERR_ARG;
    *n_selected = 0;
    std::vector<uint64_t> first((size_t)n + 1, 0);          // similars[i] = pairs[first[i] .. first[i+1]) after the stable bucket pass
    for (uint64_t k = 0; k < n_pairs; k++) {
        if (pairs[k].i < 0 || (uint32_t)pairs[k].i >= n || pairs[k].j < 0 || (uint32_t)pairs[k].j >= n) return YGPU_ERR_ARG;
        first[(size_t)pairs[k].i + 1]++;
    }
    for (uint32_t g = 0; g < n; g++) first[g + 1] += first[g];
    std::vector<int32_t> sim((size_t)n_pairs);
    {
        std::vector<uint64_t> fill(first.begin(), first.end() - 1);
        for (uint64_t k = 0; k < n_pairs; k++) sim[fill[pairs[k].i]++] = pairs[k].j;     // j ascending per i when pairs are (i, j)-sorted
    }
    std::vector<std::pair<int, int>> genome_id_size_pairs(n);
    for (uint32_t g = 0; g < n; g++) genome_id_size_pairs[g] = {(int)g, (int)(offsets[g + 1] - offsets[g])};
    std::sort(genome_id_size_pairs.begin(), genome_id_size_pairs.end(),
              [](const std::pair<int, int>& a, const std::pair<int, int>& b) { return a.second < b.second; });
    std::vector<char> excluded(n, 0);
    uint32_t ns = 0;
    for (uint32_t v = 0; v < n; v++) {
        const int g = genome_id_size_pairs[v].first;
        const int size_this = genome_id_size_pairs[v].second;
        bool keep = true;
        for (uint64_t k = first[g]; k < first[(size_t)g + 1]; k++) {
            const int o = sim[k];
            if (excluded[o]) continue;
            if ((int)(offsets[o + 1] - offsets[o]) >= size_this) { keep = false; break; }
        }
        if (keep) selected[ns++] = g;
        else excluded[g] = 1;
    }
    *n_selected = ns;
    return 0;
}
