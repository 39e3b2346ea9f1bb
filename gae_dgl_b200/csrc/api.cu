// Library-level entry points: version, per-thread error string, tuning knobs, launch counter.
#include <atomic>
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace gae {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};
// defaults (B200 sweeps, profiles/r01_sweep_*.json): register-gather variant, 4 gathers in flight per
// lane at 40 registers (48 warps/SM), 64-thread CTAs, plain caching, warp per row; decoder dense pass with
// both GEMMs on the tensor cores (dec_mma = 1)
// PER-THREAD: a sweep on one host thread never changes what another thread's calls launch
static thread_local int32_t g_tuning[T_COUNT] = {0, 4, 64, 0, 1, 0, 2, 1, 2, 1, -1, 1, 4, 1, -1, 1};
static const char *const g_tuning_names[T_COUNT] = {"spmm_variant", "spmm_unroll", "spmm_block",
                                                     "spmm_cache", "spmm_rows_per_warp", "dec_splits",
                                                     "spmm_stages", "spmm_bins", "dec_rows",
                                                     "spmm_seg_order", "spmm_fused", "dec_mma", "push_unroll",
                                                     "push_stream_ld", "dec_tc", "gcn_fused"};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int32_t tuning(int idx) { return g_tuning[idx]; }

}  // namespace gae

extern "C" const char *gae_version(void) { return "gae_b200 0.1.0 (sm_100a)"; }
extern "C" const char *gae_last_error_string(void) { return gae::g_err; }
extern "C" int64_t gae_launch_count(void) { return gae::g_launches.load(); }

extern "C" int gae_set_tuning(const char *key, int32_t value) {
    if (key) {
        for (int i = 0; i < gae::T_COUNT; ++i)
            if (strcmp(key, gae::g_tuning_names[i]) == 0) {
                gae::g_tuning[i] = value;
                return GAE_OK;
            }
    }
    gae::set_error("unknown tuning key '%s'", key ? key : "(null)");
    return GAE_ERR_INVALID_ARG;
}
extern "C" int32_t gae_get_tuning(const char *key) {
    if (key)
        for (int i = 0; i < gae::T_COUNT; ++i)
            if (strcmp(key, gae::g_tuning_names[i]) == 0) return gae::g_tuning[i];
    return -1;
}
