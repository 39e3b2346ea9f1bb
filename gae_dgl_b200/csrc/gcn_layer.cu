// One GCN layer in ONE launch: H' = act((A H) W^T + b)  --  gae.py:26-31 (update_all(copy_src, sum) followed by
// apply_nodes(Linear + activation)), SURVEY.md 8(b) `gae_gcn_layer_fwd/bwd_f32`.
//
// The two-kernel form writes Y = A H to HBM and reads it back in the sgemm kernel; for the hidden layers of the
// reference's models (39 -> 32 -> 16, 500 -> 32 -> 16) the aggregated row fits in registers, so a warp sums the
// neighbour rows of one destination row (LPR lanes x float4 per feature row, the warp's lane groups taking alternate
// edges, fixed xor tree: the summation order is a pure function of the row), parks the row in shared memory and
// multiplies it by W^T (shared, [k][o]: conflict-free across the output lanes) straight away.  Y reaches HBM only
// when the caller asks for it (training: dW = dPre^T Y needs it; encode / evaluation pass NULL).
// d_in, d_out <= 64.  No atomics, deterministic.
#include "common.cuh"

namespace gae {

constexpr int GL_THREADS = 128;
constexpr int GL_WARPS = GL_THREADS / 32;
constexpr int GL_MAXD = 64;

template <int LPR>
__global__ void __launch_bounds__(GL_THREADS)
gcn_layer_fwd_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col, const float *__restrict__ Hin,
                     int64_t ldh, const float *__restrict__ W, const float *__restrict__ b, float *__restrict__ Hout,
                     int64_t ldo, float *__restrict__ Y, int64_t ldy, int64_t n, int d_in, int d_out, int act) {
    constexpr int G = 32 / LPR;                           // edge groups per warp
    __shared__ float Wt[GL_MAXD][GL_MAXD + 1];            // Wt[k][o] = W[o][k]
    __shared__ __align__(16) float ybuf[GL_WARPS][GL_MAXD];
    __shared__ float bias[GL_MAXD];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < d_in * d_out; i += GL_THREADS) {
        const int o = i / d_in, k = i - o * d_in;
        Wt[k][o] = __ldg(W + i);
    }
    for (int i = tid; i < d_out; i += GL_THREADS) bias[i] = b ? __ldg(b + i) : 0.f;
    __syncthreads();
    const int sub = lane % LPR, grp = lane / LPR;
    const int c0 = 4 * sub;                               // my four feature columns
    const bool cols_on = c0 < d_in;
    const int kq = (d_in + 3) / 4;                        // float4 steps of the contraction
    for (int64_t r = (int64_t)blockIdx.x * GL_WARPS + warp; r < n; r += (int64_t)gridDim.x * GL_WARPS) {
        // ---- aggregate: y = sum of the source rows of r
        const int64_t e0 = rowptr[r], e1 = rowptr[r + 1];
        float4 acc = f4_zero(), acc2 = f4_zero();
        int64_t e = e0 + grp;
        for (; e + G < e1; e += 2 * G) {                  // two gathers in flight per lane
            const int32_t s0 = __ldg(col + e), s1 = __ldg(col + e + G);
            if (cols_on) {
                const float4 v0 = __ldg(reinterpret_cast<const float4 *>(Hin + (int64_t)s0 * ldh + c0));
                const float4 v1 = __ldg(reinterpret_cast<const float4 *>(Hin + (int64_t)s1 * ldh + c0));
                f4_add(acc, v0);
                f4_add(acc2, v1);
            }
        }
        if (e < e1 && cols_on) f4_add(acc, __ldg(reinterpret_cast<const float4 *>(Hin + (int64_t)__ldg(col + e) * ldh + c0)));
        f4_add(acc, acc2);
#pragma unroll
        for (int off = LPR; off < 32; off <<= 1) f4_add(acc, f4_shfl_xor(acc, off));
        // columns beyond d_in (row padding of the caller) do not take part
        if (c0 + 1 >= d_in) acc.y = 0.f;
        if (c0 + 2 >= d_in) acc.z = 0.f;
        if (c0 + 3 >= d_in) acc.w = 0.f;
        if (!cols_on) acc.x = 0.f;
        if (grp == 0 && c0 < GL_MAXD) {
            *reinterpret_cast<float4 *>(&ybuf[warp][c0]) = acc;
            if (Y && c0 < ldy) *reinterpret_cast<float4 *>(Y + r * ldy + c0) = acc;      // ldy is a multiple of 4
        }
        __syncwarp();
        // ---- apply: out[o] = act(b[o] + sum_k y[k] W[o][k]), lane = output column (two per lane above 32)
        float o0 = bias[lane < d_out ? lane : 0], o1 = bias[lane + 32 < d_out ? lane + 32 : 0];
        const int oa = lane < d_out ? lane : 0, ob = lane + 32 < d_out ? lane + 32 : 0;
        for (int k4 = 0; k4 < kq; ++k4) {
            const float4 y4 = *reinterpret_cast<const float4 *>(&ybuf[warp][4 * k4]);      // warp-wide broadcast
            const int k = 4 * k4;
            o0 = fmaf(y4.x, Wt[k][oa], o0);
            o1 = fmaf(y4.x, Wt[k][ob], o1);
            if (k + 1 < d_in) { o0 = fmaf(y4.y, Wt[k + 1][oa], o0); o1 = fmaf(y4.y, Wt[k + 1][ob], o1); }
            if (k + 2 < d_in) { o0 = fmaf(y4.z, Wt[k + 2][oa], o0); o1 = fmaf(y4.z, Wt[k + 2][ob], o1); }
            if (k + 3 < d_in) { o0 = fmaf(y4.w, Wt[k + 3][oa], o0); o1 = fmaf(y4.w, Wt[k + 3][ob], o1); }
        }
        if (act == GAE_ACT_RELU) {
            o0 = fmaxf(o0, 0.f);
            o1 = fmaxf(o1, 0.f);
        }
        if (lane < d_out) Hout[r * ldo + lane] = o0;
        if (lane + 32 < d_out) Hout[r * ldo + lane + 32] = o1;
        __syncwarp();                                     // ybuf is reused by the next row
    }
}

}  // namespace gae

using namespace gae;

extern "C" int gae_gcn_layer_fwd_f32(const int64_t *rowptr, const int32_t *col, const float *Hin, int64_t ldh,
                                     const float *W, const float *b, float *Hout, int64_t ldo, float *Y, int64_t ldy,
                                     int64_t n, int32_t d_in, int32_t d_out, int32_t act, void *stream) {
    GAE_CHECK_ARG(n >= 0 && d_in > 0 && d_out > 0, "bad sizes");
    if (n == 0) return GAE_OK;
    GAE_CHECK_ARG(rowptr && col && Hin && W && Hout, "null pointer");
    GAE_CHECK_ARG(act == GAE_ACT_IDENTITY || act == GAE_ACT_RELU, "unknown activation");
    GAE_CHECK_ARG(ldo >= d_out, "ldo too small");
    if (d_in > GL_MAXD || d_out > GL_MAXD) {
        set_error("gae_gcn_layer_fwd_f32 supports d_in, d_out <= 64 (got %d, %d): use gae_spmm_csr_f32 + gae_linear_fwd_f32", d_in, d_out);
        return GAE_ERR_UNSUPPORTED;
    }
    const int64_t dpad = (d_in + 3) / 4 * 4;
    if (ldh < dpad || ldh % 4 != 0 || !aligned16(Hin) || (Y && (ldy < dpad || ldy % 4 != 0 || !aligned16(Y)))) {
        set_error("gae_gcn_layer_fwd_f32 needs 16-byte aligned rows (row strides multiples of 4 floats, >= d_in rounded up to 4)");
        return GAE_ERR_UNSUPPORTED;
    }
    int64_t blocks = cdiv(n, GL_WARPS);
    if (blocks > 148 * 12) blocks = 148 * 12;
    cudaStream_t st = (cudaStream_t)stream;
    if (d_in <= 16)
        gcn_layer_fwd_kernel<4><<<(unsigned)blocks, GL_THREADS, 0, st>>>(rowptr, col, Hin, ldh, W, b, Hout, ldo, Y, ldy, n, d_in, d_out, act);
    else if (d_in <= 32)
        gcn_layer_fwd_kernel<8><<<(unsigned)blocks, GL_THREADS, 0, st>>>(rowptr, col, Hin, ldh, W, b, Hout, ldo, Y, ldy, n, d_in, d_out, act);
    else
        gcn_layer_fwd_kernel<16><<<(unsigned)blocks, GL_THREADS, 0, st>>>(rowptr, col, Hin, ldh, W, b, Hout, ldo, Y, ldy, n, d_in, d_out, act);
    GAE_LAUNCH_CHECK();
    return GAE_OK;
}

extern "C" int64_t gae_gcn_layer_bwd_ws_bytes(int64_t n, int32_t d_in, int32_t d_out) {
    return gae_linear_bwd_ws_bytes(n, d_in, d_out);
}

// Adjoint of the layer: dPre = dHout (.) act'(Hout); dW = dPre^T Y, db = sum dPre, dY = dPre W (gae_linear_bwd_f32), then
// dHin = A^T dY (gae_spmm_csr_f32 over CSR(A^T)) unless dHin is NULL (first layer: the input features are a leaf, gae.py:50).
extern "C" int gae_gcn_layer_bwd_f32(const int64_t *rowptr_t, const int32_t *col_t, const gae_hub_plan_t *plan_t,
                                     float *hub_ws_t, const float *Y, int64_t ldy, const float *W, const float *Hout,
                                     int64_t ldo, const float *dHout, int64_t ld_dh, float *dY, int64_t ld_dy, float *dHin,
                                     int64_t ld_dhin, float *dW, float *db, void *ws, int64_t ws_bytes, int64_t n,
                                     int32_t d_in, int32_t d_out, int32_t act, void *stream) {
    GAE_CHECK_ARG(!dHin || (rowptr_t && col_t && dY), "dHin needs CSR(A^T) and the dY scratch rows");
    int rc = gae_linear_bwd_f32(Y, ldy, W, Hout, ldo, dHout, ld_dh, dHin ? dY : nullptr, ld_dy, dW, db, ws, ws_bytes, n, d_in,
                                d_out, act, stream);
    if (rc || !dHin) return rc;
    return gae_spmm_csr_f32(rowptr_t, col_t, nullptr, dY, ld_dy, dHin, ld_dhin, n, d_in, plan_t, hub_ws_t, 0, stream);
}
