// Shared helpers for libgae_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gae_b200.h"

namespace gae {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);
int32_t tuning(int idx);

enum TuningIdx {
    T_SPMM_VARIANT = 0,  // 0 = register gather (LDG.128), 1 = streaming + cp.async.bulk (TMA), 2 = streaming + LDGSTS
    T_SPMM_UNROLL,       // gathers in flight per lane (2 / 4 / 8); sets the register budget / occupancy
    T_SPMM_BLOCK,        // threads per CTA (32 / 64 / 128)
    T_SPMM_CACHE,        // 0 = plain ld.global.nc ; 1 = X gathers with an L2 evict_last policy + streaming Y stores
    T_SPMM_ROWS_PER_WARP,// 1 = warp per row, 2 = half-warp per row (d <= 64)
    T_DEC_SPLITS,        // 0 = auto
    T_SPMM_STAGES,       // streaming variant: batches of 32 rows in flight per warp (2/3/4)
    T_SPMM_BINS,         // 1 = use the plan's degree bins when present, 0 = single row pass
    T_DEC_ROWS,          // decoder dense pass: query rows per thread for d <= 16 (1 or 2)
    T_SPMM_SEG_ORDER,    // 1 = walk hub segments in the plan's seg_order (source-id order), 0 = row-major
    T_SPMM_FUSED,        // 1 = all row classes + hub segments of the binned forward in one launch (spmm_fused.cu), 0 = off, -1 = by size
    T_DEC_MMA,           // decoder dense pass, d <= 16: 1 = both GEMMs as split-precision TF32 MMAs (default), 0 = SIMT FFMA2
    T_PUSH_UNROLL,       // halo push kernel: row steps in flight per lane group (4 or 8; registers 55 / 106)
    T_PUSH_STREAM_LD,    // halo push kernel: 1 = read the local rows with evict-first loads (ld.global.cs)
    T_DEC_TC,            // decoder dense pass, d <= 16: tcgen05 / TMEM symmetric-half kernels -- -1 = by size (default: the fp16-split
                         // pipelined form from 5632 rows), 2 = that form (decoder_tc16.cu) from 512 rows, 1 = TF32 form
                         // (decoder_tc.cu), 0 = mma.sync / SIMT forms
    T_GCN_FUSED,         // gae_step_fwd_bwd_f32: 1 = layers with d_in, d_out <= 64 and no hub rows in ONE launch (gcn_layer.cu), 0 = SpMM + Linear
    T_COUNT
};

#define GAE_CHECK_ARG(cond, msg)                         \
    do {                                                 \
        if (!(cond)) {                                   \
            gae::set_error("invalid argument: %s", msg); \
            return GAE_ERR_INVALID_ARG;                  \
        }                                                \
    } while (0)

#define GAE_CUDA(call)                                                                   \
    do {                                                                                 \
        cudaError_t _e = (call);                                                         \
        if (_e != cudaSuccess) {                                                         \
            gae::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, \
                           __LINE__);                                                    \
            return (int)_e;                                                              \
        }                                                                                \
    } while (0)

#define GAE_LAUNCH_CHECK()                                                               \
    do {                                                                                 \
        cudaError_t _e = cudaGetLastError();                                             \
        if (_e != cudaSuccess) {                                                         \
            gae::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),   \
                           __FILE__, __LINE__);                                          \
            return (int)_e;                                                              \
        }                                                                                \
        gae::count_launch();                                                             \
    } while (0)

// arguments of the streaming SpMM variant (spmm_stream.cu)
struct StreamArgs {
    const int64_t *rowptr;
    const int32_t *col;
    const float *X;
    int64_t ldx;
    float *Y;          // rows pass: Y ; segment pass: partial buffer
    int64_t ldy;
    int64_t n_items;
    int32_t seg_len;   // rows pass: skip rows with deg > seg_len (0 = never)
    int32_t accumulate;
    const int32_t *long_row;
    const int64_t *long_seg_ptr;
    const int32_t *seg_row;
};

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- device helpers ---------------------------------------------------------------------
__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void f4_add(float4 &a, const float4 &b) {
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
}
__device__ __forceinline__ void f4_fma(float4 &a, float w, const float4 &b) {
    a.x = fmaf(w, b.x, a.x); a.y = fmaf(w, b.y, a.y); a.z = fmaf(w, b.z, a.z); a.w = fmaf(w, b.w, a.w);
}
__device__ __forceinline__ float4 f4_shfl_xor(const float4 &v, int m) {
    float4 r;
    r.x = __shfl_xor_sync(0xffffffffu, v.x, m);
    r.y = __shfl_xor_sync(0xffffffffu, v.y, m);
    r.z = __shfl_xor_sync(0xffffffffu, v.z, m);
    r.w = __shfl_xor_sync(0xffffffffu, v.w, m);
    return r;
}

__device__ __forceinline__ uint64_t make_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void st_stream_f4(float *p, const float4 &v) {
    __stcs(reinterpret_cast<float4 *>(p), v);
}

// Numerically stable pieces of BCE-with-logits, shared by decoder kernels.
//   e = exp(-|x|);  l = log(1+e);  softplus(x) = max(x,0) + l;  softplus(-x) = max(-x,0) + l;
//   sigmoid(x) = x>=0 ? 1/(1+e) : e/(1+e)
__device__ __forceinline__ void softplus_parts(float x, float &l, float &sg) {
    // raw MUFU ops (ex2 / rcp / lg2 .approx.ftz, each ~1 ulp): 1+e lies in (1,2], so none of the
    // IEEE slow paths that __frcp_rn / expf / logf carry (a conditional CALL per pair) is needed
    float e, r, l2;
    const float t = -fabsf(x) * 1.4426950408889634f;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
    const float one_e = 1.0f + e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(one_e));
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(one_e));
    l = l2 * 0.6931471805599453f;
    sg = (x >= 0.f) ? r : e * r;
}

}  // namespace gae
