// TEST-ONLY harness: evaluates the fp64 formulas of yacht_b200/csrc/binom_stats.cuh with g++ so
// their numerics can be checked against scipy / the golden workbook rows on a box without a GPU.
// It is never built into, linked with or called by the product (which runs these formulas only
// inside the CUDA kernel k6_hyp_test).
// stdin:  lines "n_excl n_match ksize significance ani cov"      stdout: the 8 outputs per line
#include <cmath>
#include <cstdio>
#include "../../yacht_b200/csrc/binom_stats.cuh"

int main() {
    long long ne, nm; int k; double sig, ani, cov;
    while (scanf("%lld %lld %d %lf %lf %lf", &ne, &nm, &k, &sig, &ani, &cov) == 6) {
        const double p0 = std::pow(ani, (double)k);
        ystats::HypRow r = ystats::single_hyp_test(ne, nm, k, sig, p0, cov);
        printf("%d %.17g %lld %lld %lld %.17g %.17g %.17g\n", r.in_sample_est, r.p_val, r.num_exclusive_kmers,
               r.num_exclusive_kmers_coverage, r.num_matches, r.acceptance_threshold_with_coverage,
               r.actual_confidence_with_coverage, r.alt_confidence_mut_rate_with_coverage);
    }
    return 0;
}
