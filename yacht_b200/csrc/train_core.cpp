// run_yacht_train_core -- B200 drop-in for the reference executable of the same name
// (KoslickiLab/YACHT src/cpp/main.cpp, launched by src/yacht/utils.py:143-147).
//
// Same command line, same files, same stdout banners:
//   run_yacht_train_core [-t T] [-c C] [-p P] file_list working_directory output_filename
//     (main.cpp:142-184; -t/-p keep their meaning for how the pair lines are split over
//      <working_directory>/<pass>_<thread:03d>.txt, :265-271,338-349)
//   output_filename: the selected sketch paths, one per line, greedy visit order (:412-418)
//
// This file is the HOST layer only: argument parsing, sketch ingest (sig_scan.hpp), the greedy
// near-duplicate removal (inherently sequential, main.cpp:371-420) and the file writers.  The
// inverted index, the pairwise shared-hash counts and the threshold/compaction run on the GPU(s)
// behind the C ABI of include/yacht_gpu.h; there is no CPU implementation of them here, and the
// program exits non-zero if no B200 is available.
//
// Multi-GPU: every visible device (or the first YACHT_NUM_GPUS of them) is one rank (a host thread) of the library's
// sharded train step: it holds, of every sketch, the hashes of its hash range (cut out of the parsed, sorted sketches
// on the host), builds the index of that range locally and sends the work items to the ranks that own the query rows
// over NVLink; the pair lists are gathered over NCCL (include/yacht_gpu.h: ygpu_comm_init /
// ygpu_load_sketches_hashrange / ygpu_train_step_sharded).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <fstream>
#include <iostream>
#include <mutex>
#include <string>
#include <thread>
#include <unistd.h>
#include <utility>
#include <vector>

#include "../../include/yacht_gpu.h"
#include "ingest.hpp"

namespace {

struct Arguments {
    std::string file_list, working_directory, output_filename;
    int number_of_threads = 1;
    int num_of_passes = 1;
    double containment_threshold = 0.9;
};

const char* kVersion = "1.0";

void usage(const char* prog) {
    std::cout << "Usage: " << prog << " [--help] [--version] [--threads VAR] [--passes VAR] "
              << "[--containment_threshold VAR] file_list working_directory output_filename\n\n"
              << "Positional arguments:\n"
              << "  file_list                    file containing list of files to be processed \n"
              << "  working_directory            working directory (where temp files are generated) \n"
              << "  output_filename              output filename (where the reduced ref filenames will be written) \n\n"
              << "Optional arguments:\n"
              << "  -h, --help                   shows help message and exits \n"
              << "  -v, --version                prints version information and exits \n"
              << "  -t, --threads                number of threads [nargs=0..1] [default: 1]\n"
              << "  -p, --passes                 number of passes [nargs=0..1] [default: 1]\n"
              << "  -c, --containment_threshold  containment threshold [nargs=0..1] [default: 0.9]\n";
}

bool parse_int(const char* s, int* out) {
    errno = 0;
    char* e = nullptr;
    const long v = strtol(s, &e, 10);
    if (e == s || *e != 0 || errno == ERANGE || v < INT32_MIN || v > INT32_MAX) return false;
    *out = (int)v;
    return true;
}

bool parse_double(const char* s, double* out) {
    errno = 0;
    char* e = nullptr;
    const double v = strtod(s, &e);  // the reference's argparse also ends in strtod (argparse.hpp:370-392)
    if (e == s || *e != 0) return false;
    *out = v;
    return true;
}

// returns 0 = go on, 1 = error (exit 1), 2 = help/version shown (exit 0)
int parse_arguments(int argc, char** argv, Arguments& a) {
    std::vector<std::string> pos;
    for (int i = 1; i < argc; i++) {
        const std::string s = argv[i];
        auto need = [&](const char* name) -> const char* {
            if (i + 1 >= argc) { std::cerr << "Too few arguments for '" << name << "'." << std::endl; return nullptr; }
            return argv[++i];
        };
        if (s == "-h" || s == "--help") { usage(argv[0]); return 2; }
        if (s == "-v" || s == "--version") { std::cout << kVersion << std::endl; return 2; }
        if (s == "-t" || s == "--threads") {
            const char* v = need("-t");
            if (!v || !parse_int(v, &a.number_of_threads)) { std::cerr << "invalid value for --threads" << std::endl; return 1; }
        } else if (s == "-p" || s == "--passes") {
            const char* v = need("-p");
            if (!v || !parse_int(v, &a.num_of_passes)) { std::cerr << "invalid value for --passes" << std::endl; return 1; }
        } else if (s == "-c" || s == "--containment_threshold") {
            const char* v = need("-c");
            if (!v || !parse_double(v, &a.containment_threshold)) { std::cerr << "invalid value for --containment_threshold" << std::endl; return 1; }
        } else if (s.size() > 1 && s[0] == '-' && !(s[1] >= '0' && s[1] <= '9') && s[1] != '.') {
            std::cerr << "Unknown argument: " << s << std::endl;
            return 1;
        } else {
            pos.push_back(s);
        }
    }
    if (pos.size() < 3) {
        static const char* names[3] = {"file_list", "working_directory", "output_filename"};
        std::cerr << names[pos.size()] << ": 1 argument(s) expected. 0 provided." << std::endl;
        return 1;
    }
    if (pos.size() > 3) { std::cerr << "Maximum number of positional arguments exceeded" << std::endl; return 1; }
    a.file_list = pos[0];
    a.working_directory = pos[1];
    a.output_filename = pos[2];
    // same validation and messages as main.cpp:172-182
    if (a.number_of_threads < 1) { std::cerr << "number of threads must be at least 1" << std::endl; return 1; }
    if (a.num_of_passes < 1) { std::cerr << "number of passes must be at least 1" << std::endl; return 1; }
    if (a.containment_threshold < 0.0 || a.containment_threshold > 1.0) {
        std::cerr << "containment threshold must be between 0.0 and 1.0" << std::endl;
        return 1;
    }
    return 0;
}

void show_arguments(const Arguments& a) {  // main.cpp:187-199
    using std::cout; using std::endl;
    cout << "Working with the following parameters:" << endl;
    cout << "**************************************" << endl;
    cout << "*" << endl;
    cout << "*    file_list: " << a.file_list << endl;
    cout << "*    working_directory: " << a.working_directory << endl;
    cout << "*    output_filename: " << a.output_filename << endl;
    cout << "*    number_of_threads: " << a.number_of_threads << endl;
    cout << "*    num_of_passes: " << a.num_of_passes << endl;
    cout << "*    containment_threshold: " << a.containment_threshold << endl;
    cout << "*" << endl;
    cout << "**************************************" << endl;
}

struct DeviceResult {
    int rc = 0;
    std::string err;
    ygpu_pair* pairs = nullptr;
    uint64_t n_pairs = 0;
    ygpu_index_stats stats{};
    ygpu_timings tm{};
};

int64_t ms_since(std::chrono::high_resolution_clock::time_point t0) {
    return std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::high_resolution_clock::now() - t0).count();
}

}  // namespace

int main(int argc, char** argv) {
    const auto t_main = std::chrono::high_resolution_clock::now();
    Arguments args;
    const int pa = parse_arguments(argc, argv, args);
    if (pa == 2) return 0;
    if (pa == 1) {
        std::cout << "Usage: " << argv[0] << " -h" << std::endl;  // main.cpp:434
        return 1;
    }
    show_arguments(args);

    // ---- read.  CUDA initialisation (~1 s per process on this class of host) and the creation of the GPU
    //      contexts run in the background from the very start; parsed blocks are shipped to the GPU(s) while later
    //      files are still being parsed: parser threads queue the ids of finished blocks, uploader threads --
    //      started as soon as the contexts exist -- feed ygpu_upload_block.
    auto t_read = std::chrono::high_resolution_clock::now();
    std::cout << "Reading all sketches in filelist using all " << args.number_of_threads << " threads..." << std::endl;
    yingest::Ingest in;
    {
        std::ifstream fl(args.file_list);
        if (!fl.is_open()) std::cerr << "Could not open the filelist: " << args.file_list << std::endl;  // main.cpp:131
        std::string line;
        while (std::getline(fl, line)) in.names.push_back(line);
    }
    const uint32_t n = (uint32_t)in.names.size();
    std::cout << "Total number of sketches to read: " << n << std::endl;
    int ndev = 0;
    std::vector<ygpu_ctx*> ctxs;
    std::vector<DeviceResult> res;
    const bool stream_upload = !getenv("YACHT_NO_STREAM_UPLOAD");
    std::mutex q_mu;
    std::condition_variable q_cv;
    std::deque<uint32_t> q_blocks;
    bool q_done = false;
    std::atomic<int> upload_rc{0};
    std::vector<std::thread> uploaders;
    // Multi-GPU: rank d (one thread per GPU) will hold, of every sketch, the hashes of its hash range (cut on the host once all
    // files are parsed); a single GPU receives the parsed blocks while later files are still being read.
    const uint32_t nblocks_total = (n + yingest::kFilesPerBlock - 1) / yingest::kFilesPerBlock;
    uint8_t comm_id[YGPU_COMM_ID_BYTES];
    std::atomic<bool> want_stream{false};
    if (stream_upload)
        in.on_block = [&](uint32_t b) {
            { std::lock_guard<std::mutex> lk(q_mu); q_blocks.push_back(b); }
            q_cv.notify_one();
        };
    std::thread gpu_init([&]() {
        int nd = ygpu_device_count();
        if (nd < 1) return;
        // How many GPUs: YACHT_NUM_GPUS when set; otherwise one GPU up to 200 000 sketches -- an 85 205-genome database is
        // indexed and compared in ~15 ms on one B200, while bringing up a communicator over several GPUs costs seconds --
        // and every visible GPU beyond that.
        if (const char* e = getenv("YACHT_NUM_GPUS")) { int v = atoi(e); if (v >= 1) nd = std::min(nd, v); }
        else if (n <= 200000u) nd = 1;
        nd = std::max(1, std::min<int>(nd, (int)std::max<uint32_t>(nblocks_total, 1)));
        ctxs.assign(nd, nullptr);
        res.resize(nd);
        if (nd > 1 && ygpu_comm_get_unique_id(comm_id)) {
            res[0].rc = -1; res[0].err = ygpu_last_error(nullptr);
            ndev = nd;
            return;
        }
        std::vector<std::thread> th;
        for (int d = 0; d < nd; d++)
            th.emplace_back([&, d]() {
                res[d].rc = ygpu_ctx_create(&ctxs[d], d);
                if (res[d].rc) { res[d].err = ygpu_last_error(nullptr); return; }
                if (nd > 1) {
                    res[d].rc = ygpu_comm_init(ctxs[d], d, nd, comm_id);
                    if (res[d].rc) { res[d].err = ygpu_last_error(ctxs[d]); return; }
                } else if (stream_upload) {
                    res[d].rc = ygpu_upload_begin(ctxs[d]);
                    if (res[d].rc) res[d].err = ygpu_last_error(ctxs[d]);
                }
            });
        for (auto& t : th) t.join();
        ndev = nd;
        if (!stream_upload || nd > 1) return;
        if (res[0].rc) return;
        want_stream = true;
        const int nu = std::max(2, std::min(4, args.number_of_threads / 4));
        for (int k = 0; k < nu; k++)
            uploaders.emplace_back([&, k, nu]() {
                yingest::pin_worker(k * std::max(1, args.number_of_threads / nu));   // unpinned helper threads starve here
                for (;;) {
                    uint32_t b;
                    {
                        std::unique_lock<std::mutex> lk(q_mu);
                        q_cv.wait(lk, [&] { return !q_blocks.empty() || q_done; });
                        if (q_blocks.empty()) return;
                        b = q_blocks.front();
                        q_blocks.pop_front();
                    }
                    const int rc = ygpu_upload_block(ctxs[0], b, in.blocks[b].data(), in.blocks[b].size());
                    if (rc) upload_rc = rc;
                }
            });
    });
    yingest::read_sketches(in, args.number_of_threads, /*assemble_flat=*/false);
    {
        { std::lock_guard<std::mutex> lk(q_mu); q_done = true; }
        q_cv.notify_all();
        gpu_init.join();
        for (auto& t : uploaders) t.join();
    }
    if (ndev < 1) {
        std::cerr << "run_yacht_train_core: no CUDA device available; this build has no CPU path" << std::endl;
        return 3;
    }
    if (in.fatal) {
        std::cerr << "run_yacht_train_core: cannot parse signature " << in.fatal_msg << std::endl;
        return 4;
    }
    std::cout << "All sketches read" << std::endl;
    std::cout << "Number of empty sketches: " << in.empty_ids.size() << std::endl;  // main.cpp:202-212
    if (!in.empty_ids.empty()) {
        std::cout << "Empty sketch ids: ";
        for (int i : in.empty_ids) std::cout << i << " ";
        std::cout << std::endl;
    }
    std::cout << "Time taken to read all sketches: " << ms_since(t_read) << " milliseconds" << std::endl;

    // ---- index + pairwise on the GPU(s) ----------------------------------------------------------
    auto t_index = std::chrono::high_resolution_clock::now();
    std::cout << "Building index from sketches..." << std::endl;
    for (int d = 0; d < ndev; d++)
        if (res[d].rc) {
            std::cerr << "run_yacht_train_core: GPU " << d << ": " << res[d].err << std::endl;
            return 5;
        }
    const size_t nb = in.blocks.size();
    std::vector<const uint64_t*> block_ptrs(nb);
    std::vector<uint64_t> block_lens(nb);
    for (size_t b = 0; b < nb; b++) { block_ptrs[b] = in.blocks[b].data(); block_lens[b] = in.blocks[b].size(); }
    // query rows of rank d: contiguous genome ranges of nearly equal hash counts; hash range of rank d: equal-width pieces
    // of [0, largest hash] (FracMinHash hashes are uniform below max_hash)
    std::vector<uint32_t> g_lo(ndev + 1, n);
    g_lo[0] = 0;
    {
        const uint64_t Tall = in.offsets[n];
        int d = 1;
        for (uint32_t g = 0; g < n && d < ndev; g++)
            while (d < ndev && in.offsets[g + 1] * (uint64_t)ndev >= Tall * (uint64_t)d) g_lo[d++] = g + 1;
    }
    uint64_t max_hash = 0;
    for (size_t b = 0; b < nb; b++) max_hash = std::max(max_hash, in.block_max[b]);
    std::vector<uint64_t> cuts(ndev + 1, ~0ull);
    for (int d = 0; d < ndev; d++) cuts[d] = (uint64_t)(((unsigned __int128)max_hash + 1) * (unsigned)d / (unsigned)ndev);
    std::vector<uint32_t> sizes32(n);
    for (uint32_t g = 0; g < n; g++) sizes32[g] = (uint32_t)(in.offsets[g + 1] - in.offsets[g]);
    // One thread per GPU from here on.  With several GPUs each rank cuts its share out of the parsed sketches, loads it
    // (collective) and runs the library's sharded train step; every rank ends up with the complete pair list and rank 0's is
    // written out.
    std::vector<ygpu_pair> pairs;
    ygpu_index_stats S{};
    bool sharded_refused = false;
    auto t_mat = t_index;
    {
        std::vector<std::thread> th;
        for (int d = 0; d < ndev; d++)
            th.emplace_back([&, d]() {
                DeviceResult& r = res[d];
                ygpu_ctx* c = ctxs[d];
                if (ndev == 1) {
                    std::vector<uint64_t> block_dst(nb, 0);
                    for (size_t b = 0; b < nb; b++) block_dst[b] = in.offsets[std::min<size_t>(b * yingest::kFilesPerBlock, n)];
                    if (want_stream)
                        r.rc = upload_rc ? (int)upload_rc : ygpu_upload_finish(c, block_dst.data(), (uint32_t)nb, in.offsets.data(), n);
                    else
                        r.rc = ygpu_load_sketch_blocks(c, block_ptrs.data(), block_lens.data(), (uint32_t)nb, in.offsets.data(), n);
                    if (!r.rc) r.rc = ygpu_build_index(c, &r.stats);
                    if (!r.rc) r.rc = ygpu_pairwise_flag(c, args.containment_threshold, 0, n, &r.pairs, &r.n_pairs);
                } else {
                    yingest::pin_worker(d * std::max(1, args.number_of_threads / ndev));
                    const uint64_t lo = cuts[d], hi = cuts[d + 1];
                    const bool last = d == ndev - 1;
                    std::vector<uint64_t> part_off((size_t)n + 1, 0), part;
                    part.reserve((size_t)(in.offsets[n] / ndev + in.offsets[n] / (16 * ndev) + 1024));
                    for (uint32_t g = 0; g < n; g++) {
                        const size_t b = g / yingest::kFilesPerBlock;
                        const uint64_t* sk = in.blocks[b].data() + (in.offsets[g] - in.offsets[b * yingest::kFilesPerBlock]);
                        const uint64_t* se = sk + sizes32[g];
                        if (in.block_sorted[b]) {
                            const uint64_t* a0 = std::lower_bound(sk, se, lo);
                            const uint64_t* a1 = last ? se : std::lower_bound(a0, se, hi);
                            part.insert(part.end(), a0, a1);
                        } else {
                            for (const uint64_t* x = sk; x < se; x++)
                                if (*x >= lo && (last || *x < hi)) part.push_back(*x);
                        }
                        part_off[g + 1] = part.size();
                    }
                    r.rc = ygpu_load_sketches_hashrange(c, part.data(), part_off.data(), sizes32.data(), n, g_lo[d], g_lo[d + 1]);
                    if (!r.rc) {
                        r.rc = ygpu_train_step_sharded(c, args.containment_threshold, &r.stats, &r.n_pairs);
                        if (!r.rc && d == 0 && r.n_pairs) {
                            r.pairs = (ygpu_pair*)malloc(r.n_pairs * sizeof(ygpu_pair));
                            r.rc = r.pairs ? ygpu_pairs_copy(c, r.pairs, 0) : -5;
                        }
                    }
                }
                if (r.rc) r.err = ygpu_last_error(c);
                ygpu_get_timings(c, &r.tm);
            });
        for (auto& t : th) t.join();
    }
    if (ndev > 1) {
        // the sharded step refuses databases it does not cover (extreme skew, tiny inputs): the same answer on every rank.
        // One GPU then takes the whole database through the general entry points.
        bool refused = false, failed = false;
        for (int d = 0; d < ndev; d++) { if (res[d].rc == -4) refused = true; else if (res[d].rc) failed = true; }
        if (refused && !failed) {
            // replicated index, rows split by measured work (ygpu_train_step_replicated): every GPU receives all sketches and
            // builds the full index; for such databases the pairwise count dominates, and it splits cleanly by rows
            sharded_refused = true;
            std::cout << "[multi-gpu] sharded step not applicable (" << res[0].err << "); replicated index, rows split over " << ndev << " GPUs" << std::endl;
            std::vector<std::thread> th;
            for (int d = 0; d < ndev; d++)
                th.emplace_back([&, d]() {
                    DeviceResult& r = res[d];
                    ygpu_ctx* c = ctxs[d];
                    r.n_pairs = 0;
                    r.rc = ygpu_load_sketch_blocks(c, block_ptrs.data(), block_lens.data(), (uint32_t)nb, in.offsets.data(), n);
                    if (!r.rc) r.rc = ygpu_train_step_replicated(c, args.containment_threshold, &r.stats, &r.n_pairs);
                    if (!r.rc && d == 0 && r.n_pairs) {
                        r.pairs = (ygpu_pair*)malloc(r.n_pairs * sizeof(ygpu_pair));
                        r.rc = r.pairs ? ygpu_pairs_copy(c, r.pairs, 0) : -5;
                    }
                    if (r.rc) r.err = ygpu_last_error(c);
                    ygpu_get_timings(c, &r.tm);
                });
            for (auto& t : th) t.join();
        }
    }
    for (int d = 0; d < ndev; d++)
        if (res[d].rc) {
            std::cerr << "run_yacht_train_core: GPU " << d << ": " << res[d].err << std::endl;
            return 5;
        }
    S = res[0].stats;
    std::cout << "Total number of distinct hashes: " << S.n_distinct << std::endl;                                       // main.cpp:242
    std::cout << "Total number of distinct hashes that appear in only one sketch: " << S.n_singleton << std::endl;      // :243
    std::cout << "Size of the index: " << S.n_index << std::endl;                                                       // :244
    std::cout << "Time taken to build index: " << ms_since(t_index) << " milliseconds" << std::endl;

    t_mat = std::chrono::high_resolution_clock::now();
    std::cout << "Computing intersection matrix..." << std::endl;
    pairs.assign(res[0].pairs, res[0].pairs + res[0].n_pairs);       // sorted by (i, j): the complete list
    if (res[0].pairs) { ygpu_free(res[0].pairs); res[0].pairs = nullptr; }
    (void)sharded_refused;

    // ---- pair files: same file partition as main.cpp:318,338-349, same line format as :305 --------
    {
        const int P = args.num_of_passes, Tn = args.number_of_threads;
        const int per_pass = (int)std::ceil(1.0 * n / P);
        size_t k = 0;
        for (int pass = 0; pass < P; pass++) {
            const int ps = pass * per_pass;
            const int pe = (pass == P - 1) ? (int)n : (pass + 1) * per_pass;
            const int n_this = pe - ps;
            const int chunk = n_this / Tn;
            for (int t = 0; t < Tn; t++) {
                const int re = (t == Tn - 1) ? pe : ps + (t + 1) * chunk;
                std::string id = std::to_string(t);
                while (id.size() < 3) id = "0" + id;
                const std::string fn = args.working_directory + "/" + std::to_string(pass) + "_" + id + ".txt";
                std::ofstream outfile(fn);
                while (k < pairs.size() && pairs[k].i < re) {
                    const ygpu_pair& pr = pairs[k++];
                    const size_t ni = in.offsets[pr.i + 1] - in.offsets[pr.i];
                    const size_t nj = in.offsets[pr.j + 1] - in.offsets[pr.j];
                    const int m = pr.count;
                    // the reference's expressions, operand types included (main.cpp:296-298)
                    const double jaccard = 1.0 * m / (ni + nj - m);
                    const double c_ij = 1.0 * m / ni;
                    const double c_ji = 1.0 * m / nj;
                    outfile << pr.i << "," << pr.j << "," << jaccard << "," << c_ij << "," << c_ji << "\n";
                }
                outfile.close();
            }
            std::cout << "Pass " << pass + 1 << "/" << P << " done." << std::endl;
        }
    }
    std::cout << "Time taken to compute intersection matrix: " << ms_since(t_mat) << " milliseconds" << std::endl;

    // ---- greedy near-duplicate removal (main.cpp:371-420), on the host -----------------------------
    auto t_train = std::chrono::high_resolution_clock::now();
    std::cout << "Starting yacht train..." << std::endl;
    std::cout << "Starting yacht train..." << std::endl;
    // ygpu_greedy_select (host code of the library): same element type, initial order (file-list order), comparator
    // and std::sort call as the reference, so genomes of equal size are visited in the same order.
    std::vector<int32_t> selected(std::max<uint32_t>(n, 1));
    uint32_t n_selected = 0;
    if (ygpu_greedy_select(in.offsets.data(), n, pairs.data(), pairs.size(), selected.data(), &n_selected)) {
        std::cerr << "run_yacht_train_core: greedy selection failed" << std::endl;
        return 5;
    }
    selected.resize(n_selected);
    std::cout << "Writing to output file.." << std::endl;
    {
        std::ofstream outfile(args.output_filename);
        for (int g : selected) outfile << in.names[g] << std::endl;
        outfile.close();
    }
    std::cout << "Time taken to do yacht train: " << ms_since(t_train) << " milliseconds" << std::endl;

    // extra (not in the reference): device-side timings, for apples-to-apples phase accounting
    for (int d = 0; d < ndev; d++) {
        const ygpu_timings& t = res[d].tm;
        std::cout << "[gpu " << d << "] rows [" << g_lo[d] << "," << g_lo[d + 1] << ") h2d " << t.ms_h2d << " ms, sort "
                  << t.ms_sort << " ms, index " << t.ms_index << " ms, count+flag " << t.ms_count << " ms, d2h " << t.ms_d2h
                  << " ms, pairs " << res[d].n_pairs << std::endl;
    }
    std::cout << "[wall] " << ms_since(t_main) << " ms from process start to results on disk" << std::endl;
    // Every output is on disk.  Tearing down ~20 GB of device buffers, the CUDA context and the parsed blocks costs
    // more than the whole GPU computation; a command-line process leaves that to the OS (YACHT_TRAIN_TEARDOWN=1 keeps
    // the orderly path, e.g. under compute-sanitizer).
    if (!getenv("YACHT_TRAIN_TEARDOWN")) {
        std::cout.flush();
        std::cerr.flush();
        fflush(nullptr);
        _exit(0);
    }
    for (auto* c : ctxs) ygpu_ctx_destroy(c);
    if (in.hashes) { if (in.pinned) ygpu_host_free(in.hashes); else free(in.hashes); }
    return 0;
}
