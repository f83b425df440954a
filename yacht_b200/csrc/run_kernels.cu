// placeholder, replaced below
#include "common.cuh"
void ygpu_run_release(ygpu_ctx* ctx) { (void)ctx; }
extern "C" int ygpu_exclusive_hashes(ygpu_ctx* ctx, const uint64_t*, uint64_t, const uint8_t*, ygpu_genome_counts*) { return ygpu_fail(ctx, YGPU_ERR_STATE, "not built yet"); }
extern "C" int ygpu_hyp_test(ygpu_ctx* ctx, const int64_t*, const int64_t*, uint64_t, int, double, double, const double*, int, ygpu_hyp_row*) { return ygpu_fail(ctx, YGPU_ERR_STATE, "not built yet"); }
