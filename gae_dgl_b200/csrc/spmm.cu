// K1 / K2 -- CSR SpMM with sum aggregation:  Y[v,:] = sum_{e in row v} (w_e) X[col[e],:]
//
// Replaces DGL's update_all(copy_src, sum) (reference gae.py:18-19,28) and, with CSR(A^T),
// its adjoint.  This is an HBM/L2-bound gather-reduce (0.24 flop/B): no tensor cores.
//
// Mapping (vector path): a feature row of d floats is covered by LPR lanes x float4
// (d=64 -> 16 lanes, one 256-B row = two full 128-B lines per gather).  A warp holds
// G = 32/LPR lane groups; GPR of them cooperate on one dst row (edge e goes to group
// e mod GPR), so a warp owns RPW = G/GPR rows.  Every lane keeps U independent 128-bit
// gathers in flight (index loads first, then the U row loads, then the adds), the groups of
// a row are combined with a fixed xor-shuffle tree and one group stores the row with
// coalesced 128-bit stores.  The summation order is a pure function of (row range, GPR):
// no atomics, run-to-run deterministic.
//
// Degree skew (RMAT: max in-degree > 1e5, ~40 % empty rows): rows longer than
// plan->seg_len are skipped by the row kernel and handled by the SAME device routine
// launched over fixed-length segments that write partial rows, followed by a tiny ordered
// reduce -- deterministic two-stage reduction instead of atomics.  Empty rows are written
// as zeros by the row kernel.
#include "common.cuh"

namespace gae {

struct SpmmArgs {
    const int64_t *rowptr;
    const int32_t *col;
    const float *vals;
    const float *X;
    int64_t ldx;
    float *Y;        // row kernel: Y ; segment kernel: partial buffer
    int64_t ldy;     // row kernel: ldy ; segment kernel: partial stride
    int64_t n_items; // rows or segments
    int32_t d;
    int32_t seg_len;  // >0: row kernel skips rows with deg > seg_len
    int32_t accumulate;
    // segment kernel only
    const int32_t *long_row;
    const int64_t *long_seg_ptr;
    const int32_t *seg_row;
    // optional indirection: item -> row (degree-binned launches)
    const int32_t *row_list;
    // optional processing order of the segments
    const int32_t *seg_order;
};

// All gathers are `asm volatile` so that their program order (U index loads, then U row
// loads, then the adds) survives NVVM/ptxas scheduling: without it the compiler sinks each
// load next to its add and keeps only 2-3 requests in flight per lane.
template <int CACHE>
__device__ __forceinline__ float4 gather_f4(const float *p, uint64_t pol) {
    float4 r;
    if (CACHE == 1) {
        asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                     : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                     : "l"(p), "l"(pol));
    } else {
        asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                     : "l"(p));
    }
    return r;
}
__device__ __forceinline__ int ld_idx(const int32_t *p) {
    int r;
    asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

// LPR lanes per feature row, GPR groups per dst row, U gathers in flight per lane.
// MINB: resident 128-thread CTAs per SM the register allocation must allow (occupancy target).
template <int LPR, int GPR, int U, bool VALS, int CACHE, bool SEG, int MINB>
__global__ void __launch_bounds__(128, MINB) spmm_vec_kernel(const SpmmArgs a) {
    constexpr int G = 32 / LPR;
    constexpr int RPW = G / GPR;
    static_assert(G % GPR == 0, "GPR must divide G");
    const int lane = threadIdx.x & 31;
    const int sub = lane % LPR;
    const int grp = lane / LPR;
    const int phase = grp % GPR;  // which edges of the row this group takes
    const int rsel = grp / GPR;   // which of the warp's rows
    const int64_t warp_global = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t item = warp_global * RPW + rsel;

    bool active = item < a.n_items;
    int64_t start = 0, end = 0;
    float *out = nullptr;
    if (active) {
        if (!SEG) {
            const int64_t row = a.row_list ? (int64_t)__ldg(a.row_list + item) : item;
            start = __ldg(a.rowptr + row);
            end = __ldg(a.rowptr + row + 1);
            if (a.seg_len > 0 && end - start > (int64_t)a.seg_len) {  // hub row: segment pass
                active = false;
                start = end = 0;
            }
            out = a.Y + row * a.ldy;
        } else {
            const int64_t j = a.seg_order ? (int64_t)__ldg(a.seg_order + item) : item;   // segment id
            const int32_t k = __ldg(a.seg_row + j);
            const int64_t row = __ldg(a.long_row + k);
            const int64_t s = j - __ldg(a.long_seg_ptr + k);
            const int64_t r0 = __ldg(a.rowptr + row), r1 = __ldg(a.rowptr + row + 1);
            start = r0 + s * (int64_t)a.seg_len;
            end = min(start + (int64_t)a.seg_len, r1);
            out = a.Y + j * a.ldy;
        }
    }
    uint64_t pol = 0;
    if (CACHE == 1) pol = make_policy_evict_last();

    const int d = a.d;
    // LPR == 32: 128-float column chunks (wide features, e.g. d_in = 500 / 1433).
    // (wide rows: the 128-float column chunks are independent, so they are spread over gridDim.y
    // instead of being walked one after the other by the same warp)
    {
        const int ch = (LPR == 32) ? (int)blockIdx.y : 0;
        const int c4 = sub + ch * LPR;
        const bool colok = c4 * 4 < d;
        const float *xb = a.X + (int64_t)(colok ? c4 : 0) * 4;  // idle lanes re-read column 0
        float4 acc = f4_zero();
        // Loads are UNCONDITIONAL (edge index clamped to the row's last edge, which re-hits a
        // line already in flight) so the compiler keeps all U index loads and then all U row
        // gathers in flight; only the adds are predicated.
        // 32-bit edge offsets inside the row keep the loop control to a few integer ops
        const int len = (int)(end - start);
        const int32_t *cp = a.col + start;
        const float *wp = VALS ? a.vals + start : nullptr;
        const int ldx = (int)a.ldx;
        for (int i = phase; i < len; i += GPR * U) {
            int c[U];
            float w[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int ii = min(i + u * GPR, len - 1);
                c[u] = ld_idx(cp + ii);
                if (VALS) w[u] = __ldg(wp + ii);
            }
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = gather_f4<CACHE>(xb + (int64_t)c[u] * ldx, pol);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (i + u * GPR < len) {
                    if (VALS) f4_fma(acc, w[u], v[u]);
                    else f4_add(acc, v[u]);
                }
            }
        }
        if (GPR > 1) {
            __syncwarp();
#pragma unroll
            for (int off = LPR; off < LPR * GPR; off <<= 1) f4_add(acc, f4_shfl_xor(acc, off));
        }
        if (active && phase == 0 && colok) {
            float *o = out + (int64_t)c4 * 4;
            if (c4 * 4 + 4 <= d) {
                if (!SEG && a.accumulate) f4_add(acc, *reinterpret_cast<const float4 *>(o));
                if (CACHE != 0) st_stream_f4(o, acc);
                else *reinterpret_cast<float4 *>(o) = acc;
            } else {  // ragged tail (d % 4 != 0): scalar stores of the valid components
                const float t[4] = {acc.x, acc.y, acc.z, acc.w};
                for (int k = 0; c4 * 4 + k < d; ++k)
                    o[k] = (!SEG && a.accumulate) ? o[k] + t[k] : t[k];
            }
        }
    }
}

// Short rows (1..U in-edges): LPR lanes x CPL float4 cover a row, 32/LPR rows per warp, every
// gather of the row in flight at once (U x CPL 128-bit loads per lane) -- one latency per row
// instead of a warp and a dependent chain per row.
template <int LPR, int CPL, int U>
__global__ void __launch_bounds__(128, 8) spmm_short_rows_kernel(const SpmmArgs a) {
    constexpr int G = 32 / LPR;
    const int lane = threadIdx.x & 31;
    const int sub = lane % LPR, grp = lane / LPR;
    const int64_t item = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * G + grp;
    if (item >= a.n_items) return;
    const int64_t row = __ldg(a.row_list + item);
    const int64_t start = __ldg(a.rowptr + row);
    const int len = (int)(__ldg(a.rowptr + row + 1) - start);
    const int32_t *cp = a.col + start;
    const int ldx = (int)a.ldx;
    int c[U];
#pragma unroll
    for (int u = 0; u < U; ++u) c[u] = ld_idx(cp + min(u, len - 1));
    float4 v[U][CPL];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
            const int c4 = q * LPR + sub;
            v[u][q] = gather_f4<0>(a.X + (int64_t)c[u] * ldx + (c4 * 4 < a.d ? c4 : 0) * 4, 0);
        }
    float4 acc[CPL];
#pragma unroll
    for (int q = 0; q < CPL; ++q) acc[q] = v[0][q];
#pragma unroll
    for (int u = 1; u < U; ++u)
        if (u < len) {
#pragma unroll
            for (int q = 0; q < CPL; ++q) f4_add(acc[q], v[u][q]);
        }
    float *out = a.Y + row * a.ldy;
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
        const int c4 = q * LPR + sub;
        if (c4 * 4 + 4 <= a.d) {
            float *o = out + c4 * 4;
            if (a.accumulate) f4_add(acc[q], *reinterpret_cast<const float4 *>(o));
            *reinterpret_cast<float4 *>(o) = acc[q];
        }
    }
}

// Empty rows: streaming zero fill (Y = A X must still define them).
__global__ void spmm_zero_rows_kernel(const int32_t *__restrict__ rows, int64_t n, float *__restrict__ Y,
                                      int64_t ldy, int d4) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t r = t / d4;
    if (r >= n) return;
    st_stream_f4(Y + (int64_t)__ldg(rows + r) * ldy + (t % d4) * 4, f4_zero());
}

// Ordered reduce of the segment partials of each long row: Y[row] (+)= sum_s P[seg_s].
__global__ void spmm_hub_reduce_kernel(const float *__restrict__ P, int64_t ldp,
                                       const int32_t *__restrict__ long_row,
                                       const int64_t *__restrict__ long_seg_ptr, int64_t n_long,
                                       float *__restrict__ Y, int64_t ldy, int d, int accumulate) {
    const int d4 = (d + 3) >> 2;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t k = t / d4;
    const int c4 = (int)(t % d4);
    if (k >= n_long) return;
    const int64_t s0 = long_seg_ptr[k], s1 = long_seg_ptr[k + 1];
    float4 acc = f4_zero();
    for (int64_t s = s0; s < s1; ++s) f4_add(acc, *reinterpret_cast<const float4 *>(P + s * ldp + c4 * 4));
    float *o = Y + (int64_t)long_row[k] * ldy + c4 * 4;
    const float tt[4] = {acc.x, acc.y, acc.z, acc.w};
    for (int q = 0; q < 4 && c4 * 4 + q < d; ++q) o[q] = accumulate ? o[q] + tt[q] : tt[q];
}

// Scalar fallback (unaligned pointers / leading dimensions): warp per row, lane per column.
__global__ void spmm_scalar_kernel(const SpmmArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= a.n_items) return;
    const int64_t start = a.rowptr[row], end = a.rowptr[row + 1];
    for (int c = lane; c < a.d; c += 32) {
        float acc = 0.f;
        for (int64_t e = start; e < end; ++e) {
            const float x = __ldg(a.X + (int64_t)a.col[e] * a.ldx + c);
            acc = a.vals ? fmaf(a.vals[e], x, acc) : acc + x;
        }
        float *o = a.Y + row * a.ldy + c;
        *o = a.accumulate ? *o + acc : acc;
    }
}

template <int LPR, int GPR, int U, bool VALS, int CACHE, bool SEG>
static cudaError_t launch_one(const SpmmArgs &a, int block, cudaStream_t st) {
    constexpr int RPW = (32 / LPR) / GPR;
    // occupancy target: U=4 fits 40 registers (12 CTAs of 128 threads), U=8 needs 64 (8 CTAs)
    constexpr int MINB = (U <= 2) ? 16 : (U <= 4) ? 12 : 8;
    const int wpb = block / 32;
    const int64_t warps = cdiv(a.n_items, RPW);
    const int64_t blocks = cdiv(warps, wpb);
    if (blocks == 0) return cudaSuccess;
    const dim3 grid((unsigned)blocks, LPR == 32 ? (unsigned)((a.d + 127) >> 7) : 1u);
    spmm_vec_kernel<LPR, GPR, U, VALS, CACHE, SEG, MINB><<<grid, block, 0, st>>>(a);
    count_launch();
    return cudaGetLastError();
}

template <int LPR, int GPR, bool SEG>
static cudaError_t launch_shape(const SpmmArgs &a, int unroll, int cache, int block, cudaStream_t st) {
    if (a.vals) return launch_one<LPR, GPR, 4, true, 0, SEG>(a, block, st);
    if (cache == 1) {
        if (unroll >= 8) return launch_one<LPR, GPR, 8, false, 1, SEG>(a, block, st);
        return launch_one<LPR, GPR, 4, false, 1, SEG>(a, block, st);
    }
    if (unroll >= 8) return launch_one<LPR, GPR, 8, false, 0, SEG>(a, block, st);
    if (unroll <= 2) return launch_one<LPR, GPR, 2, false, 0, SEG>(a, block, st);
    return launch_one<LPR, GPR, 4, false, 0, SEG>(a, block, st);
}

template <bool SEG>
static cudaError_t launch_vec(const SpmmArgs &a, int rows_per_warp, int unroll, int cache, int block,
                              cudaStream_t st) {
    const int d4 = (a.d + 3) / 4;
    const bool split = (rows_per_warp <= 1) || SEG;  // all groups of the warp on one row
    if (d4 <= 4) return split ? launch_shape<4, 8, SEG>(a, unroll, cache, block, st)
                              : launch_shape<4, 1, SEG>(a, unroll, cache, block, st);
    if (d4 <= 8) return split ? launch_shape<8, 4, SEG>(a, unroll, cache, block, st)
                              : launch_shape<8, 1, SEG>(a, unroll, cache, block, st);
    if (d4 <= 16) return split ? launch_shape<16, 2, SEG>(a, unroll, cache, block, st)
                               : launch_shape<16, 1, SEG>(a, unroll, cache, block, st);
    return launch_shape<32, 1, SEG>(a, unroll, cache, block, st);
}

// rows up to which the automatic setting takes the single-launch form (profiles/r02_fused_sweep.log)
constexpr int64_t FUSED_AUTO_MAX_ROWS = (int64_t)1 << 40;

cudaError_t spmm_stream_launch(const StreamArgs &a, int d, bool seg, int stages, int mode, cudaStream_t st);
bool spmm_fused_launch(const int64_t *rowptr, const int32_t *col, const float *X, int64_t ldx, float *Y, int64_t ldy,
                       int32_t d, const gae_hub_plan_t *plan, float *partial_ws, int64_t ldp, bool seg_order,
                       cudaStream_t st, cudaError_t *err);

}  // namespace gae

using namespace gae;

extern "C" int gae_spmm_csr_f32(const int64_t *rowptr, const int32_t *col, const float *vals,
                                const float *X, int64_t ldx, float *Y, int64_t ldy, int64_t n_rows,
                                int32_t d, const gae_hub_plan_t *plan, float *partial_ws,
                                int32_t accumulate, void *stream) {
    GAE_CHECK_ARG(n_rows >= 0 && d >= 0, "n_rows, d must be >= 0");
    if (n_rows == 0 || d == 0) return GAE_OK;
    GAE_CHECK_ARG(rowptr && X && Y, "rowptr, X, Y must be non-null");
    GAE_CHECK_ARG(ldx >= d && ldy >= d, "leading dimensions must be >= d");
    GAE_CHECK_ARG(ldx < (int64_t)1 << 31, "ldx must fit in 31 bits");
    cudaStream_t st = (cudaStream_t)stream;
    const bool use_plan = plan && plan->n_seg > 0;
    if (use_plan) {
        GAE_CHECK_ARG(plan->seg_len > 0, "plan->seg_len must be > 0");
        GAE_CHECK_ARG(plan->long_row && plan->long_seg_ptr && plan->seg_row, "plan arrays must be non-null");
        GAE_CHECK_ARG(partial_ws != nullptr, "partial_ws required when the plan has segments");
    }
    SpmmArgs a{};
    a.rowptr = rowptr; a.col = col; a.vals = vals; a.X = X; a.ldx = ldx; a.Y = Y; a.ldy = ldy;
    a.n_items = n_rows; a.d = d; a.seg_len = use_plan ? plan->seg_len : 0; a.accumulate = accumulate;

    const bool vec = aligned16(X) && aligned16(Y) && (ldx % 4 == 0) && (ldy % 4 == 0) &&
                     (!use_plan || aligned16(partial_ws));
    int block = tuning(T_SPMM_BLOCK);
    if (block != 32 && block != 64 && block != 128) block = 64;
    const int unroll = tuning(T_SPMM_UNROLL);
    const int cache = tuning(T_SPMM_CACHE);
    const int rpw = tuning(T_SPMM_ROWS_PER_WARP);

    if (!vec) {
        // correctness-only path; hub rows are not split (one warp each)
        a.seg_len = 0;
        const int64_t blocks = cdiv(n_rows, block / 32);
        spmm_scalar_kernel<<<(unsigned)blocks, block, 0, st>>>(a);
        GAE_LAUNCH_CHECK();
        return GAE_OK;
    }
    const int variant = tuning(T_SPMM_VARIANT);
    const bool stream_ok = (variant == 1 || variant == 2) && !vals && (d == 32 || d == 64 || d == 128) &&
                           ldy % 4 == 0;
    if (stream_ok) {
        // streaming variant: bulk-async staged gather, sequential segmented sum (spmm_stream.cu)
        StreamArgs sa{};
        sa.rowptr = rowptr; sa.col = col; sa.X = X; sa.ldx = ldx; sa.Y = Y; sa.ldy = ldy; sa.n_items = n_rows;
        sa.seg_len = a.seg_len; sa.accumulate = accumulate;
        const int stages = tuning(T_SPMM_STAGES);
        GAE_CUDA(spmm_stream_launch(sa, d, false, stages, variant - 1, st));
        if (use_plan) {
            StreamArgs ss = sa;
            ss.Y = partial_ws; ss.ldy = d; ss.n_items = plan->n_seg; ss.seg_len = plan->seg_len; ss.accumulate = 0;
            ss.long_row = plan->long_row; ss.long_seg_ptr = plan->long_seg_ptr; ss.seg_row = plan->seg_row;
            GAE_CUDA(spmm_stream_launch(ss, d, true, stages, variant - 1, st));
            const int64_t threads = plan->n_long * (d / 4);
            spmm_hub_reduce_kernel<<<(unsigned)cdiv(threads, 256), 256, 0, st>>>(
                partial_ws, d, plan->long_row, plan->long_seg_ptr, plan->n_long, Y, ldy, d, accumulate);
            GAE_LAUNCH_CHECK();
        }
        return GAE_OK;
    }
    const bool binned = plan && plan->mid_rows && plan->short_rows && plan->empty_rows && d % 4 == 0 && d <= 64 &&
                        plan->short_max == 4 && !vals && tuning(T_SPMM_BINS) != 0;
    // single-launch form (spmm_fused.cu): 1 = on, 0 = off, -1 = by size (FUSED_AUTO_MAX_ROWS)
    const int fused_knob = tuning(T_SPMM_FUSED);
    const bool fused = fused_knob > 0 || (fused_knob < 0 && n_rows <= FUSED_AUTO_MAX_ROWS);
    if (binned && !accumulate && cache == 0 && fused) {
        const int64_t ldp = (int64_t)((d + 3) / 4) * 4;
        cudaError_t e = cudaSuccess;
        if (spmm_fused_launch(rowptr, col, X, ldx, Y, ldy, d, plan, partial_ws, ldp, tuning(T_SPMM_SEG_ORDER) != 0, st, &e)) {
            GAE_CUDA(e);
            if (use_plan) {
                const int64_t threads = plan->n_long * ((d + 3) / 4);
                spmm_hub_reduce_kernel<<<(unsigned)cdiv(threads, 256), 256, 0, st>>>(
                    partial_ws, ldp, plan->long_row, plan->long_seg_ptr, plan->n_long, Y, ldy, d, 0);
                GAE_LAUNCH_CHECK();
            }
            return GAE_OK;
        }
    }
    if (binned) {
        // degree-binned row pass: zero fill | short rows 4 per warp | a warp per remaining row
        if (plan->n_empty > 0 && !accumulate) {
            const int64_t threads = plan->n_empty * (d / 4);
            spmm_zero_rows_kernel<<<(unsigned)cdiv(threads, 256), 256, 0, st>>>(plan->empty_rows, plan->n_empty, Y, ldy, d / 4);
            GAE_LAUNCH_CHECK();
        }
        if (plan->n_short > 0) {
            SpmmArgs sh = a;
            sh.row_list = plan->short_rows; sh.n_items = plan->n_short;
            const int64_t warps = cdiv(plan->n_short, 4);
            spmm_short_rows_kernel<8, 2, 4><<<(unsigned)cdiv(warps, 4), 128, 0, st>>>(sh);
            GAE_LAUNCH_CHECK();
        }
        if (plan->n_mid > 0) {
            SpmmArgs md = a;
            md.row_list = plan->mid_rows; md.n_items = plan->n_mid;
            GAE_CUDA(launch_vec<false>(md, rpw, unroll, cache, block, st));
        }
    } else {
        GAE_CUDA(launch_vec<false>(a, rpw, unroll, cache, block, st));
    }
    if (use_plan) {
        const int64_t ldp = (int64_t)((d + 3) / 4) * 4;
        SpmmArgs s = a;
        s.Y = partial_ws; s.ldy = ldp; s.n_items = plan->n_seg; s.seg_len = plan->seg_len;
        s.long_row = plan->long_row; s.long_seg_ptr = plan->long_seg_ptr; s.seg_row = plan->seg_row;
        s.row_list = nullptr;
        s.seg_order = tuning(T_SPMM_SEG_ORDER) ? plan->seg_order : nullptr;
        GAE_CUDA(launch_vec<true>(s, 1, unroll, cache, block, st));
        const int64_t threads = plan->n_long * ((d + 3) / 4);
        spmm_hub_reduce_kernel<<<(unsigned)cdiv(threads, 256), 256, 0, st>>>(
            partial_ws, ldp, plan->long_row, plan->long_seg_ptr, plan->n_long, Y, ldy, d, accumulate);
        GAE_LAUNCH_CHECK();
    }
    return GAE_OK;
}

extern "C" int gae_spmm_csr_f32_host(const int64_t *rowptr, const int32_t *col, const float *X_host,
                                     int64_t n_src, int64_t ldx, float *Y_host, int64_t ldy,
                                     int64_t n_rows, int32_t d, const gae_hub_plan_t *plan,
                                     float *partial_ws, float *X_stage, float *Y_stage, void *stream) {
    GAE_CHECK_ARG(X_host && Y_host && X_stage && Y_stage, "host and staging buffers must be non-null");
    GAE_CHECK_ARG(n_src >= 0 && n_rows >= 0, "sizes must be >= 0");
    cudaStream_t st = (cudaStream_t)stream;
    GAE_CUDA(cudaMemcpyAsync(X_stage, X_host, sizeof(float) * (size_t)n_src * ldx, cudaMemcpyHostToDevice, st));
    int rc = gae_spmm_csr_f32(rowptr, col, nullptr, X_stage, ldx, Y_stage, ldy, n_rows, d, plan, partial_ws, 0, stream);
    if (rc != GAE_OK) return rc;
    GAE_CUDA(cudaMemcpyAsync(Y_host, Y_stage, sizeof(float) * (size_t)n_rows * ldy, cudaMemcpyDeviceToHost, st));
    return GAE_OK;
}

extern "C" int gae_hub_plan_count_host(const int64_t *rowptr, int64_t n_rows, int32_t seg_len,
                                       int64_t *n_long, int64_t *n_seg) {
    GAE_CHECK_ARG(rowptr && n_long && n_seg, "null pointer");
    GAE_CHECK_ARG(seg_len > 0 && n_rows >= 0, "seg_len must be > 0");
    GAE_CHECK_ARG(n_rows < ((int64_t)1 << 31), "row ids are int32: n_rows must be < 2^31");
    int64_t nl = 0, ns = 0;
    for (int64_t v = 0; v < n_rows; ++v) {
        const int64_t deg = rowptr[v + 1] - rowptr[v];
        if (deg > seg_len) { ++nl; ns += (deg + seg_len - 1) / seg_len; }
    }
    *n_long = nl; *n_seg = ns;
    return GAE_OK;
}

extern "C" int gae_row_bins_host(const int64_t *rowptr, int64_t n_rows, int32_t seg_len, int32_t short_max,
                                 int64_t counts[3], int32_t *empty_rows, int32_t *short_rows, int32_t *mid_rows) {
    GAE_CHECK_ARG(rowptr && counts, "null pointer");
    GAE_CHECK_ARG(seg_len > 0 && short_max > 0 && n_rows >= 0 && n_rows < ((int64_t)1 << 31), "bad sizes");
    int64_t ne = 0, ns = 0, nm = 0;
    for (int64_t v = 0; v < n_rows; ++v) {
        const int64_t deg = rowptr[v + 1] - rowptr[v];
        if (deg == 0) { if (empty_rows) empty_rows[ne] = (int32_t)v; ++ne; }
        else if (deg > seg_len) { /* hub row: owned by the segment pass, in no bin */ }
        else if (deg <= short_max) { if (short_rows) short_rows[ns] = (int32_t)v; ++ns; }
        else { if (mid_rows) mid_rows[nm] = (int32_t)v; ++nm; }
    }
    counts[0] = ne; counts[1] = ns; counts[2] = nm;
    return GAE_OK;
}

extern "C" int gae_hub_plan_fill_host(const int64_t *rowptr, int64_t n_rows, int32_t seg_len,
                                      int32_t *long_row, int64_t *long_seg_ptr, int32_t *seg_row) {
    GAE_CHECK_ARG(rowptr && long_seg_ptr, "null pointer");
    GAE_CHECK_ARG(seg_len > 0 && n_rows >= 0, "seg_len must be > 0");
    int64_t k = 0, s = 0;
    long_seg_ptr[0] = 0;
    for (int64_t v = 0; v < n_rows; ++v) {
        const int64_t deg = rowptr[v + 1] - rowptr[v];
        if (deg > seg_len) {
            const int64_t ns = (deg + seg_len - 1) / seg_len;
            long_row[k] = (int32_t)v;
            for (int64_t j = 0; j < ns; ++j) seg_row[s + j] = (int32_t)k;
            s += ns;
            long_seg_ptr[++k] = s;
        }
    }
    return GAE_OK;
}
