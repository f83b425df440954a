// libyachtgpu -- run path: sample membership + exclusive-hash reduction (K5) and the per-genome
// binomial statistics (K6).  sm_100a only.
//
// Reference behaviour being replaced (KoslickiLab/YACHT src/yacht/hypothesis_recovery_src.py):
//   get_organisms_with_nonzero_overlap :30-113   `sourmash scripts multisearch ... -t 0`: the genomes
//                                                whose sketch shares >= 1 hash with the sample
//   get_exclusive_hashes               :116-206  python sets: hashes occurring in exactly one of the
//                                                nontrivial genomes, and how many are in the sample
//   single_hyp_test / get_alt_mut_rate :209-306  scipy binom.ppf / binom.cdf / betaincinv
// Design: both reductions walk the (hash, genome) array that K2a already sorted -- an equal-hash run
// is "all genomes holding this hash".  The sample is sorted once and fronted by a bucket directory
// over its leading bits, so a membership probe is one directory read plus a search inside a bucket
// of a few hashes.  One thread per run head does the probe / the distinct-nontrivial-genome scan
// (with early exit at the second genome); per-genome counters are integer atomics in HBM.
#include "common.cuh"
#include "binom_stats.cuh"

#include <cub/cub.cuh>
#include <algorithm>
#include <cmath>
#include <vector>

struct RunScratch {
    uint64_t* d_sample = nullptr;      // sorted sample
    uint64_t* d_sample_in = nullptr;
    uint64_t cap = 0;
    uint32_t* d_dir = nullptr;         // bucket directory over the sorted sample
    uint64_t dir_cap = 0;
    uint8_t* d_head_in = nullptr;      // [T] at run heads: 1 if the hash is in the sample
    uint64_t head_cap = 0;
    ygpu_genome_counts* d_counts = nullptr;
    uint8_t* d_mask = nullptr;
    uint64_t n_cap = 0;
};

static RunScratch* scratch(ygpu_ctx* ctx) {
    if (!ctx->run_scratch) ctx->run_scratch = new RunScratch();
    return (RunScratch*)ctx->run_scratch;
}

void ygpu_run_release(ygpu_ctx* ctx) {
    RunScratch* r = (RunScratch*)ctx->run_scratch;
    if (!r) return;
    if (r->d_sample) cudaFree(r->d_sample);
    if (r->d_sample_in) cudaFree(r->d_sample_in);
    if (r->d_dir) cudaFree(r->d_dir);
    if (r->d_head_in) cudaFree(r->d_head_in);
    if (r->d_counts) cudaFree(r->d_counts);
    if (r->d_mask) cudaFree(r->d_mask);
    delete r;
    ctx->run_scratch = nullptr;
}

// dir[b] = first i with (sample[i] >> shift) >= b, b in [0, nb]
__global__ void __launch_bounds__(256) k5_sample_dir(const uint64_t* __restrict__ samp, uint64_t ns, int shift, uint32_t nb,
                                                      uint32_t* __restrict__ dir) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= ns; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t b_hi = (i < ns) ? (samp[i] >> shift) : (uint64_t)nb;          // buckets <= b_hi start at or before i
        const uint64_t b_lo = (i == 0) ? 0 : ((samp[i - 1] >> shift) + 1);           // buckets >= b_lo start at or after i
        for (uint64_t b = b_lo; b <= b_hi && b <= nb; b++) dir[b] = (uint32_t)i;
    }
}

__device__ __forceinline__ bool sample_has(const uint64_t* __restrict__ samp, const uint32_t* __restrict__ dir, int shift,
                                           uint32_t nb, uint64_t key) {
    const uint64_t b = key >> shift;
    if (b >= nb) return false;
    uint32_t lo = dir[b], hi = dir[b + 1];
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        const uint64_t v = samp[mid];
        if (v == key) return true;
        if (v < key) lo = mid + 1; else hi = mid;
    }
    return false;
}

// phase A: per distinct reference hash, is it in the sample?  If so every genome of the run (once
// per genome, whatever its multiplicity in the sketch: sets) gets one more overlapping hash.
__global__ void __launch_bounds__(256) k5_overlap(const uint64_t* __restrict__ key, const uint32_t* __restrict__ sgid, uint64_t T,
                                                   const uint64_t* __restrict__ samp, const uint32_t* __restrict__ dir, int shift,
                                                   uint32_t nb, uint8_t* __restrict__ head_in,
                                                   ygpu_genome_counts* __restrict__ counts) {
    for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < T; s += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t k = key[s];
        if (s > 0 && key[s - 1] == k) continue;       // not a run head
        const bool in = sample_has(samp, dir, shift, nb, k);
        head_in[s] = in ? 1 : 0;
        if (!in) continue;
        uint32_t prev = 0xffffffffu;
        for (uint64_t x = s; x < T && key[x] == k; x++) {
            const uint32_t g = sgid[x];
            if (g != prev) atomicAdd(&counts[g].n_overlap, 1u);
            prev = g;
        }
    }
}

__global__ void __launch_bounds__(256) k5_nontrivial(ygpu_genome_counts* __restrict__ counts, const uint8_t* __restrict__ mask,
                                                      uint32_t n) {
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < n; g += gridDim.x * blockDim.x)
        counts[g].nontrivial = mask ? (mask[g] ? 1u : 0u) : (counts[g].n_overlap > 0 ? 1u : 0u);
}

// phase B: a hash is exclusive iff exactly one nontrivial genome holds it
__global__ void __launch_bounds__(256) k5_exclusive(const uint64_t* __restrict__ key, const uint32_t* __restrict__ sgid, uint64_t T,
                                                     const uint8_t* __restrict__ head_in,
                                                     ygpu_genome_counts* __restrict__ counts) {
    for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < T; s += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t k = key[s];
        if (s > 0 && key[s - 1] == k) continue;
        uint32_t owner = 0xffffffffu;
        bool exclusive = true;
        for (uint64_t x = s; x < T && key[x] == k; x++) {
            const uint32_t g = sgid[x];
            if (g == owner || !counts[g].nontrivial) continue;
            if (owner != 0xffffffffu) { exclusive = false; break; }
            owner = g;
        }
        if (exclusive && owner != 0xffffffffu) {
            atomicAdd(&counts[owner].n_exclusive, 1u);
            if (head_in[s]) atomicAdd(&counts[owner].n_match, 1u);
        }
    }
}

extern "C" int ygpu_exclusive_hashes(ygpu_ctx* ctx, const uint64_t* sample, uint64_t n_sample, const uint8_t* mask,
                                     ygpu_genome_counts* counts) {
    if (!ctx || !counts) return YGPU_ERR_ARG;
    if (!ctx->loaded) return ygpu_fail(ctx, YGPU_ERR_STATE, "exclusive_hashes: no sketches loaded");
    if (n_sample && !sample) return ygpu_fail(ctx, YGPU_ERR_ARG, "exclusive_hashes: NULL sample");
    if (n_sample >= (1ull << 32)) return ygpu_fail(ctx, YGPU_ERR_ARG, "sample too large");
    YG_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint32_t n = ctx->n;
    const uint64_t T = ctx->T;
    if (n == 0) return 0;
    RunScratch* r = scratch(ctx);
    if (n_sample + 1 > r->cap) {
        if (r->d_sample) cudaFree(r->d_sample);
        if (r->d_sample_in) cudaFree(r->d_sample_in);
        r->d_sample = r->d_sample_in = nullptr; r->cap = 0;
        YG_CUDA(ctx, cudaMalloc(&r->d_sample, (n_sample + 1) * sizeof(uint64_t)));
        YG_CUDA(ctx, cudaMalloc(&r->d_sample_in, (n_sample + 1) * sizeof(uint64_t)));
        r->cap = n_sample + 1;
    }
    if (T + 1 > r->head_cap) {
        if (r->d_head_in) cudaFree(r->d_head_in);
        r->d_head_in = nullptr; r->head_cap = 0;
        YG_CUDA(ctx, cudaMalloc(&r->d_head_in, T + 1));
        r->head_cap = T + 1;
    }
    if (n > r->n_cap) {
        if (r->d_counts) cudaFree(r->d_counts);
        if (r->d_mask) cudaFree(r->d_mask);
        r->d_counts = nullptr; r->d_mask = nullptr; r->n_cap = 0;
        YG_CUDA(ctx, cudaMalloc(&r->d_counts, (size_t)n * sizeof(ygpu_genome_counts)));
        YG_CUDA(ctx, cudaMalloc(&r->d_mask, n));
        r->n_cap = n;
    }

    YG_CUDA(ctx, cudaEventRecord(ctx->ev[0], st));
    YG_CUDA(ctx, cudaMemsetAsync(r->d_counts, 0, (size_t)n * sizeof(ygpu_genome_counts), st));
    if (mask) YG_CUDA(ctx, cudaMemcpyAsync(r->d_mask, mask, n, cudaMemcpyHostToDevice, st));
    ctx->last_run_path = 0;
    // ---- preferred: probe the MSD-partitioned reference bucket by bucket (run_buckets.cuh): no sort of the reference or the sample
    if (T && ctx->run_path != 0) {
        if (n_sample) YG_CUDA(ctx, cudaMemcpyAsync(r->d_sample_in, sample, n_sample * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        YG_CUDA(ctx, cudaEventRecord(ctx->ev[2], st));
        int used = 0;
        YG_CHECK(ygpu_run_counts_buckets(ctx, r->d_sample_in, n_sample, mask ? r->d_mask : nullptr, r->d_counts, &used));
        if (used) {
            YG_CUDA(ctx, cudaMemcpyAsync(counts, r->d_counts, (size_t)n * sizeof(ygpu_genome_counts), cudaMemcpyDeviceToHost, st));
            YG_CUDA(ctx, cudaEventRecord(ctx->ev[1], st));
            YG_CUDA(ctx, cudaStreamSynchronize(st));
            float ms = 0.f, msk = 0.f;
            cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
            cudaEventElapsedTime(&msk, ctx->evp[0], ctx->evp[1]);
            ctx->tm.ms_sample += ms;
            ctx->tm.ms_sample_kernels += msk;
            return 0;
        }
        // not applicable: start over on the general path (counters may hold partial credits)
        YG_CUDA(ctx, cudaMemsetAsync(r->d_counts, 0, (size_t)n * sizeof(ygpu_genome_counts), st));
    }
    YG_CHECK(ygpu_sort_sketches(ctx));

    int shift = 0;
    uint32_t nb = 0;
    if (n_sample) {
        YG_CUDA(ctx, cudaMemcpyAsync(r->d_sample_in, sample, n_sample * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        size_t tb = 0;
        YG_CUDA(ctx, cub::DeviceRadixSort::SortKeys(nullptr, tb, r->d_sample_in, r->d_sample, (int64_t)n_sample, 0, 64, st));
        YG_CHECK(ygpu_temp_reserve(ctx, tb));
        tb = ctx->temp_bytes;
        YG_CUDA(ctx, cub::DeviceRadixSort::SortKeys(ctx->d_temp, tb, r->d_sample_in, r->d_sample, (int64_t)n_sample, 0, 64, st));
        ctx->tm.n_library_launches += 10;
        uint64_t smax = 0;
        YG_CUDA(ctx, cudaMemcpyAsync(&smax, r->d_sample + (n_sample - 1), sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        YG_CUDA(ctx, cudaStreamSynchronize(st));
        int bits = 1;
        while (bits < 64 && (smax >> bits) != 0) bits++;
        int want = 1;
        while (want < 26 && (1ull << want) < 2 * n_sample) want++;
        want = std::max(want, 8);
        const int dirbits = std::min(bits, want);
        shift = bits - dirbits;
        nb = (uint32_t)(smax >> shift) + 1;
        if ((uint64_t)nb + 2 > r->dir_cap) {
            if (r->d_dir) cudaFree(r->d_dir);
            r->d_dir = nullptr; r->dir_cap = 0;
            YG_CUDA(ctx, cudaMalloc(&r->d_dir, ((uint64_t)nb + 2) * sizeof(uint32_t)));
            r->dir_cap = (uint64_t)nb + 2;
        }
        k5_sample_dir<<<grid_for(ctx, n_sample + 1, 256), 256, 0, st>>>(r->d_sample, n_sample, shift, nb, r->d_dir);
        YG_CUDA(ctx, cudaGetLastError());
        ctx->tm.n_kernel_launches++;
    }
    if (T) {
        if (n_sample) {
            k5_overlap<<<grid_for(ctx, T, 256), 256, 0, st>>>(ctx->d_skey, ctx->d_sgid, T, r->d_sample, r->d_dir, shift, nb,
                                                             r->d_head_in, r->d_counts);
            YG_CUDA(ctx, cudaGetLastError());
        } else {
            YG_CUDA(ctx, cudaMemsetAsync(r->d_head_in, 0, T, st));
        }
        k5_nontrivial<<<grid_for(ctx, n, 256), 256, 0, st>>>(r->d_counts, mask ? r->d_mask : nullptr, n);
        YG_CUDA(ctx, cudaGetLastError());
        k5_exclusive<<<grid_for(ctx, T, 256), 256, 0, st>>>(ctx->d_skey, ctx->d_sgid, T, r->d_head_in, r->d_counts);
        YG_CUDA(ctx, cudaGetLastError());
        ctx->tm.n_kernel_launches += 3;
    } else {
        k5_nontrivial<<<grid_for(ctx, n, 256), 256, 0, st>>>(r->d_counts, mask ? r->d_mask : nullptr, n);
        YG_CUDA(ctx, cudaGetLastError());
        ctx->tm.n_kernel_launches++;
    }
    YG_CUDA(ctx, cudaMemcpyAsync(counts, r->d_counts, (size_t)n * sizeof(ygpu_genome_counts), cudaMemcpyDeviceToHost, st));
    YG_CUDA(ctx, cudaEventRecord(ctx->ev[1], st));
    YG_CUDA(ctx, cudaStreamSynchronize(st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    ctx->tm.ms_sample += ms;
    return 0;
}

// ============================================================================================
// K6: one thread per (coverage, genome) evaluation of single_hyp_test, all in fp64
// ============================================================================================
__global__ void __launch_bounds__(128) k6_hyp_test(const long long* __restrict__ n_excl, const long long* __restrict__ n_match,
                                                    uint64_t n, int ksize, double significance, double non_mut_p,
                                                    const double* __restrict__ cov, int n_cov, ygpu_hyp_row* __restrict__ rows) {
    const uint64_t total = n * (uint64_t)n_cov;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t c = t / n, r = t - c * n;
        const ystats::HypRow h = ystats::single_hyp_test(n_excl[r], n_match[r], ksize, significance, non_mut_p, cov[c]);
        ygpu_hyp_row o;
        o.in_sample_est = h.in_sample_est;
        o._pad = 0;
        o.p_val = h.p_val;
        o.num_exclusive_kmers = h.num_exclusive_kmers;
        o.num_exclusive_kmers_coverage = h.num_exclusive_kmers_coverage;
        o.num_matches = h.num_matches;
        o.acceptance_threshold_with_coverage = h.acceptance_threshold_with_coverage;
        o.actual_confidence_with_coverage = h.actual_confidence_with_coverage;
        o.alt_confidence_mut_rate_with_coverage = h.alt_confidence_mut_rate_with_coverage;
        rows[t] = o;
    }
}

extern "C" int ygpu_hyp_test(ygpu_ctx* ctx, const int64_t* n_exclusive, const int64_t* n_match, uint64_t n, int ksize,
                             double significance, double ani_thresh, const double* min_coverage, int n_cov,
                             ygpu_hyp_row* rows) {
    if (!ctx) return YGPU_ERR_ARG;
    if (n == 0 || n_cov <= 0) return 0;
    if (!n_exclusive || !n_match || !min_coverage || !rows) return ygpu_fail(ctx, YGPU_ERR_ARG, "hyp_test: NULL argument");
    if (ksize < 1) return ygpu_fail(ctx, YGPU_ERR_ARG, "hyp_test: ksize must be >= 1");
    YG_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    // non_mut_p = ani_thresh ** ksize (hypothesis_recovery_src.py:254), evaluated once on the host
    // with the C library pow -- the function CPython's float.__pow__ calls
    const double non_mut_p = std::pow(ani_thresh, (double)ksize);
    const uint64_t total = n * (uint64_t)n_cov;
    long long *d_ne = nullptr, *d_nm = nullptr;
    double* d_cov = nullptr;
    ygpu_hyp_row* d_rows = nullptr;
    auto cleanup = [&]() {
        if (d_ne) cudaFree(d_ne);
        if (d_nm) cudaFree(d_nm);
        if (d_cov) cudaFree(d_cov);
        if (d_rows) cudaFree(d_rows);
    };
    cudaError_t e;
    if ((e = cudaMalloc(&d_ne, n * sizeof(long long))) != cudaSuccess || (e = cudaMalloc(&d_nm, n * sizeof(long long))) != cudaSuccess ||
        (e = cudaMalloc(&d_cov, n_cov * sizeof(double))) != cudaSuccess || (e = cudaMalloc(&d_rows, total * sizeof(ygpu_hyp_row))) != cudaSuccess) {
        cleanup();
        return ygpu_fail(ctx, YGPU_ERR_NOMEM, "hyp_test: cudaMalloc: %s", cudaGetErrorString(e));
    }
    cudaEventRecord(ctx->ev[0], st);
    cudaMemcpyAsync(d_ne, n_exclusive, n * sizeof(long long), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_nm, n_match, n * sizeof(long long), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_cov, min_coverage, n_cov * sizeof(double), cudaMemcpyHostToDevice, st);
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((total + 127) / 128, (uint64_t)ctx->num_sms * 16));
    k6_hyp_test<<<grid, 128, 0, st>>>(d_ne, d_nm, n, ksize, significance, non_mut_p, d_cov, n_cov, d_rows);
    e = cudaGetLastError();
    ctx->tm.n_kernel_launches++;
    if (e == cudaSuccess) e = cudaMemcpyAsync(rows, d_rows, total * sizeof(ygpu_hyp_row), cudaMemcpyDeviceToHost, st);
    cudaEventRecord(ctx->ev[1], st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        cleanup();
        return ygpu_fail(ctx, YGPU_ERR_CUDA, "hyp_test: %s", cudaGetErrorString(e));
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    ctx->tm.ms_stats += ms;
    cleanup();
    return 0;
}

// get_alt_mut_rate on its own (the reference unit-tests it directly: tests/test_unit.py:11-20)
__global__ void __launch_bounds__(128) k6_alt_mut_rate(const long long* __restrict__ nu, const long long* __restrict__ thresh,
                                                        uint64_t n, int ksize, double significance, double* __restrict__ out) {
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (uint64_t)gridDim.x * blockDim.x) {
        const double x = ystats::betaincinv_int((double)nu[t], (double)thresh[t], significance);
        double mut = 1.0 - pow(1.0 - x, 1.0 / (double)ksize);
        if (isnan(mut)) mut = -1.0;
        out[t] = mut;
    }
}

extern "C" int ygpu_alt_mut_rate(ygpu_ctx* ctx, const int64_t* nu, const int64_t* thresh, uint64_t n, int ksize,
                                 double significance, double* out) {
    if (!ctx) return YGPU_ERR_ARG;
    if (n == 0) return 0;
    if (!nu || !thresh || !out) return ygpu_fail(ctx, YGPU_ERR_ARG, "alt_mut_rate: NULL argument");
    if (ksize < 1) return ygpu_fail(ctx, YGPU_ERR_ARG, "alt_mut_rate: ksize must be >= 1");
    YG_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    long long *d_nu = nullptr, *d_th = nullptr;
    double* d_out = nullptr;
    cudaError_t e;
    if ((e = cudaMalloc(&d_nu, n * 8)) != cudaSuccess || (e = cudaMalloc(&d_th, n * 8)) != cudaSuccess ||
        (e = cudaMalloc(&d_out, n * 8)) != cudaSuccess) {
        if (d_nu) cudaFree(d_nu);
        if (d_th) cudaFree(d_th);
        return ygpu_fail(ctx, YGPU_ERR_NOMEM, "alt_mut_rate: cudaMalloc: %s", cudaGetErrorString(e));
    }
    cudaMemcpyAsync(d_nu, nu, n * 8, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_th, thresh, n * 8, cudaMemcpyHostToDevice, st);
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((n + 127) / 128, (uint64_t)ctx->num_sms * 16));
    k6_alt_mut_rate<<<grid, 128, 0, st>>>(d_nu, d_th, n, ksize, significance, d_out);
    e = cudaGetLastError();
    ctx->tm.n_kernel_launches++;
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, n * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_nu); cudaFree(d_th); cudaFree(d_out);
    if (e != cudaSuccess) return ygpu_fail(ctx, YGPU_ERR_CUDA, "alt_mut_rate: %s", cudaGetErrorString(e));
    return 0;
}
