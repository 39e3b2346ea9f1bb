// Single-launch form of the degree-binned forward of spmm.cu (tuning "spmm_fused": 1 = on, 0 = off,
// -1 = automatic by size): the blocks of ONE grid take the four roles one after the other --
// hub segments (source order), mid rows (degree order), short rows, zero fill -- instead of four
// launches with their ramps and tails.  The per-row / per-segment summation is the routine of
// spmm_vec_kernel (edge e to lane group e mod GPR, U gathers in flight, xor-shuffle tree) and of
// spmm_short_rows_kernel: results are bit-identical to the separate launches (tested).  The ordered
// hub reduce stays a second (tiny) launch.
//
// Measured (B200, profiles/r02_round2_checks.log): an earlier form that INTERLEAVED hub-segment and
// mid-row blocks over blockIdx (so that every SM held both kinds at once) was 10 % slower than the
// separate launches on C4 (2.29 vs 2.09 ms; the random mid-row gathers evict the L2-resident source
// band the ordered segments live on) and 25 % faster on 1.5 M-edge graphs (0.056 vs 0.074 ms,
// launch-bound).  Roles are therefore sequential in blockIdx here.  The partitioned multi-GPU
// operator (halo.cu) runs one such launch per row block.
#include "common.cuh"

namespace gae {

struct FusedArgs {
    const int64_t *rowptr;
    const int32_t *col;
    const float *X;
    int64_t ldx;
    float *Y;
    int64_t ldy;
    float *P;                 // hub-segment partials [n_seg, ldp]
    int64_t ldp;
    int32_t d;
    int32_t seg_len;
    // work lists
    const int32_t *mid_rows;
    const int32_t *short_rows;
    const int32_t *empty_rows;
    const int32_t *long_row;
    const int64_t *long_seg_ptr;
    const int32_t *seg_row;
    const int32_t *seg_order;
    int64_t n_mid, n_short, n_empty, n_seg;
    // block ranges: nb_hub hub-segment blocks, then nb_mid, nb_short, nb_zero
    int64_t nb_hub, nb_mid, nb_short, nb_zero;
};

namespace fused {

__device__ __forceinline__ float4 gather(const float *p) {
    float4 r;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ int ldi(const int32_t *p) {
    int r;
    asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

constexpr int BLOCK = 64;          // two warps: one item (row / segment) each
constexpr int ZERO_ROWS = 16;      // rows per zero-fill block

// One row or one segment per warp; all lane groups of the warp on it (spmm_vec_kernel with RPW = 1).
template <int LPR, int U>
__device__ __forceinline__ void warp_item(const FusedArgs &a, const int32_t *cp, int len, float *out, int lane) {
    constexpr int GPR = 32 / LPR;
    const int sub = lane % LPR, phase = lane / LPR;
    const int d = a.d;
    const bool colok = sub * 4 < d;
    const float *xb = a.X + (colok ? sub : 0) * 4;
    const int ldx = (int)a.ldx;
    float4 acc = f4_zero();
    for (int i = phase; i < len; i += GPR * U) {
        int c[U];
#pragma unroll
        for (int u = 0; u < U; ++u) c[u] = ldi(cp + min(i + u * GPR, len - 1));
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = gather(xb + (int64_t)c[u] * ldx);
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (i + u * GPR < len) f4_add(acc, v[u]);
    }
    if (GPR > 1) {
        __syncwarp();
#pragma unroll
        for (int off = LPR; off < 32; off <<= 1) f4_add(acc, f4_shfl_xor(acc, off));
    }
    if (phase == 0 && colok) *reinterpret_cast<float4 *>(out + sub * 4) = acc;     // d % 4 == 0 on this path
}

}  // namespace fused

template <int LPR, int U>
__global__ void __launch_bounds__(fused::BLOCK, 24) spmm_fused_kernel(const FusedArgs a) {
    using namespace fused;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t b = blockIdx.x;
    const int64_t n_hm = a.nb_hub + a.nb_mid;
    if (b < n_hm) {
        const bool is_hub = b < a.nb_hub;
        const int64_t hub_before = is_hub ? b : a.nb_hub;
        if (is_hub) {
            const int64_t item = hub_before * 2 + warp;
            if (item >= a.n_seg) return;
            const int64_t j = a.seg_order ? (int64_t)__ldg(a.seg_order + item) : item;
            const int32_t k = __ldg(a.seg_row + j);
            const int64_t row = __ldg(a.long_row + k);
            const int64_t s = j - __ldg(a.long_seg_ptr + k);
            const int64_t r0 = __ldg(a.rowptr + row), r1 = __ldg(a.rowptr + row + 1);
            const int64_t start = r0 + s * (int64_t)a.seg_len;
            const int len = (int)(min(start + (int64_t)a.seg_len, r1) - start);
            warp_item<LPR, U>(a, a.col + start, len, a.P + j * a.ldp, lane);
        } else {
            const int64_t item = (b - hub_before) * 2 + warp;
            if (item >= a.n_mid) return;
            const int64_t row = __ldg(a.mid_rows + item);
            const int64_t start = __ldg(a.rowptr + row);
            const int len = (int)(__ldg(a.rowptr + row + 1) - start);
            warp_item<LPR, U>(a, a.col + start, len, a.Y + row * a.ldy, lane);
        }
        return;
    }
    if (b < n_hm + a.nb_short) {
        // short rows (1..4 edges): 8 lanes x 2 float4 per row, 4 rows per warp, all gathers in flight at once
        constexpr int SL = 8, CPL = 2, SU = 4;
        const int sub = lane % SL, grp = lane / SL;
        const int64_t item = ((b - n_hm) * 2 + warp) * (32 / SL) + grp;
        if (item >= a.n_short) return;
        const int64_t row = __ldg(a.short_rows + item);
        const int64_t start = __ldg(a.rowptr + row);
        const int len = (int)(__ldg(a.rowptr + row + 1) - start);
        const int32_t *cp = a.col + start;
        const int ldx = (int)a.ldx;
        int c[SU];
#pragma unroll
        for (int u = 0; u < SU; ++u) c[u] = ldi(cp + min(u, len - 1));
        float4 v[SU][CPL];
#pragma unroll
        for (int u = 0; u < SU; ++u)
#pragma unroll
            for (int q = 0; q < CPL; ++q) {
                const int c4 = q * SL + sub;
                v[u][q] = gather(a.X + (int64_t)c[u] * ldx + (c4 * 4 < a.d ? c4 : 0) * 4);
            }
        float4 acc[CPL];
#pragma unroll
        for (int q = 0; q < CPL; ++q) acc[q] = v[0][q];
#pragma unroll
        for (int u = 1; u < SU; ++u)
            if (u < len) {
#pragma unroll
                for (int q = 0; q < CPL; ++q) f4_add(acc[q], v[u][q]);
            }
        float *out = a.Y + row * a.ldy;
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
            const int c4 = q * SL + sub;
            if (c4 * 4 + 4 <= a.d) *reinterpret_cast<float4 *>(out + c4 * 4) = acc[q];
        }
        return;
    }
    // zero fill of the empty rows: ZERO_ROWS rows per block, streaming stores
    const int d4 = a.d >> 2;
    const int64_t first = (b - n_hm - a.nb_short) * ZERO_ROWS;
    const int total = ZERO_ROWS * d4;
    for (int t = threadIdx.x; t < total; t += BLOCK) {
        const int64_t r = first + t / d4;
        if (r < a.n_empty) st_stream_f4(a.Y + (int64_t)__ldg(a.empty_rows + r) * a.ldy + (t % d4) * 4, f4_zero());
    }
}

// Returns false when the shape is not covered (caller falls back to the separate launches).
bool spmm_fused_launch(const int64_t *rowptr, const int32_t *col, const float *X, int64_t ldx, float *Y, int64_t ldy,
                       int32_t d, const gae_hub_plan_t *plan, float *partial_ws, int64_t ldp, bool seg_order,
                       cudaStream_t st, cudaError_t *err) {
    *err = cudaSuccess;
    if (d % 4 != 0 || d > 64 || d <= 0) return false;
    FusedArgs a{};
    a.rowptr = rowptr; a.col = col; a.X = X; a.ldx = ldx; a.Y = Y; a.ldy = ldy; a.P = partial_ws; a.ldp = ldp;
    a.d = d; a.seg_len = plan->seg_len;
    a.mid_rows = plan->mid_rows; a.short_rows = plan->short_rows; a.empty_rows = plan->empty_rows;
    a.long_row = plan->long_row; a.long_seg_ptr = plan->long_seg_ptr; a.seg_row = plan->seg_row;
    a.seg_order = seg_order ? plan->seg_order : nullptr;
    a.n_mid = plan->n_mid; a.n_short = plan->n_short; a.n_empty = plan->n_empty; a.n_seg = plan->n_seg;
    a.nb_hub = cdiv(a.n_seg, 2);
    a.nb_mid = cdiv(a.n_mid, 2);
    a.nb_short = cdiv(a.n_short, 8);
    a.nb_zero = cdiv(a.n_empty, fused::ZERO_ROWS);
    const int64_t blocks = a.nb_hub + a.nb_mid + a.nb_short + a.nb_zero;
    if (blocks == 0) return true;
    if (blocks >= ((int64_t)1 << 31)) return false;
    const int d4 = d / 4;
    if (d4 <= 4) spmm_fused_kernel<4, 4><<<(unsigned)blocks, fused::BLOCK, 0, st>>>(a);
    else if (d4 <= 8) spmm_fused_kernel<8, 4><<<(unsigned)blocks, fused::BLOCK, 0, st>>>(a);
    else spmm_fused_kernel<16, 4><<<(unsigned)blocks, fused::BLOCK, 0, st>>>(a);
    count_launch();
    *err = cudaGetLastError();
    return true;
}

}  // namespace gae
