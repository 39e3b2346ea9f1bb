// K5 / K6 -- fused InnerProductDecoder + weighted BCE-with-logits and its adjoint.
//
// Reference: logits = mm(zd, zd.t()) (gae.py:71), adj = g.adjacency_matrix().to_dense()
// (train_inductive.py:44), loss = BCEWithLogits(logits, adj, pos_weight) (:48), backward
// (:51).  The reference materialises three N x N fp32 arrays per step (adj, logits, grad);
// here nothing N x N ever exists.  With y_ij = multiplicity of edge (j -> i):
//     N^2 L = sum_ij softplus(x_ij) + sum_{e=(i,j)} [pw softplus(-x_ij) - softplus(x_ij)]
//     dL/dx_ij = [ sigmoid(x_ij) - y_ij (pw (1 - sigmoid(x_ij)) + sigmoid(x_ij)) ] / N^2
//     dL/dZd   = (G + G^T) Zd
// The dense term is an attention-shaped pass: thread t owns R query rows z_i in registers,
// streams key rows z_j from shared memory (warp-broadcast LDS.128), and keeps the running
// loss and the d-wide gradient accumulator in registers -- one pass yields both the loss and
// the gradient (FP32 FFMA + MUFU bound: 2d FFMA + ex2 + lg2 + rcp per pair).  The j range
// is split across CTAs; split partials are combined in fixed order by the finalize kernel,
// which also adds the per-edge correction from CSR and CSR(A^T).  No atomics: deterministic.
#include <algorithm>

#include "common.cuh"

namespace gae {

constexpr int DEC_THREADS = 128;

struct DecConfig {
    int D;             // padded embedding width (16 / 32 / 64)
    int R;             // query rows per thread
    int JT;            // key rows per shared-memory tile
    int64_t nb;        // query row blocks
    int splits;        // key range splits
    int64_t j_chunk;   // key rows per split
    int64_t fin_blocks;
    bool mma;          // D = 16: both GEMMs on the tensor cores (dec_dense_mma_kernel)
    bool tc;           // D = 16: tcgen05 / TMEM symmetric-half pass (decoder_tc16.cu / decoder_tc.cu); nb = 128-row blocks then
    int tc_variant;    // 2 = fp16-split pipelined form, 1 = TF32 form
};

// decoder_tc.cu
int dec_tc_splits(int64_t n);
cudaError_t dec_tc_launch(const float *Zd, int64_t ldz, int64_t n, int d, int splits, float *dz_part, float *dzT_part,
                          double *loss_part, uint32_t *err, cudaStream_t st);
// decoder_tc16.cu (dec_tc = 2, the default): fp16-split operands, issuing warp, pipelined over two TMEM buffer sets; same outputs
int dec_tc16_splits(int64_t n);
cudaError_t dec_tc16_launch(const float *Zd, int64_t ldz, int64_t n, int d, int splits, float *dz_part, float *dzT_part,
                            double *loss_part, uint32_t *err, cudaStream_t st);
constexpr int64_t DEC_TC_MIN_ROWS = 512;      // an explicit dec_tc = 1 / 2 applies from here
constexpr int64_t DEC_TC_AUTO_ROWS = 5632;    // dec_tc = -1 (default): tcgen05 form from here (44 x 44 tiles), mma.sync form below -- a CTA's
                                              // TMEM / barrier set-up and drain cost about two tiles and the form needs three more small launches
                                              // (measured, whole decoder: 4096 rows 49 vs 40 us, 6000 rows 65 vs 69, 8192 rows 90 vs 106)

static bool dec_config(int64_t n, int32_t d, DecConfig *c) {
    if (d <= 16) { c->D = 16; c->R = tuning(T_DEC_ROWS) == 1 ? 1 : 2; }
    else if (d <= 32) { c->D = 32; c->R = 1; }
    else if (d <= 64) { c->D = 64; c->R = 1; }
    else return false;
    c->mma = d <= 16 && tuning(T_DEC_MMA) != 0;
    int tcv = tuning(T_DEC_TC);
    if (tcv < 0) tcv = n >= DEC_TC_AUTO_ROWS ? 2 : 0;
    c->tc = d <= 16 && tcv != 0 && n >= DEC_TC_MIN_ROWS;
    c->tc_variant = tcv;
    if (c->tc) {
        c->mma = false;
        c->JT = 128;
        c->nb = cdiv(n, 128);
        c->splits = tcv == 2 ? dec_tc16_splits(n) : dec_tc_splits(n);
        if (tuning(T_DEC_SPLITS) > 0) c->splits = (int)std::min<int64_t>(tuning(T_DEC_SPLITS), c->nb);
        c->j_chunk = 0;
        c->fin_blocks = cdiv(n, 256 / (c->D / 4));
        return true;
    }
    // query rows per CTA: the tensor-core kernel covers 4 warps x 2 m-tiles x 16 rows
    const int64_t rows_pb = c->mma ? 128 : (int64_t)DEC_THREADS * c->R;
    c->JT = 2048 / c->D;
    c->nb = cdiv(n, rows_pb);
    int64_t max_splits = cdiv(n, c->JT);
    int64_t want = tuning(T_DEC_SPLITS);
    if (want <= 0) {
        // ~3 waves of CTAs (4 resident 128-thread CTAs per SM at 128 registers), then nudge the
        // split count so the last wave is full: wave quantisation cost 20 % at Pubmed size
        // resident CTAs per SM by register budget: tensor-core kernel 128 registers -> 4; SIMT R = 1 at 64
        // registers -> 8, R = 2 at 143 -> 3; D = 32 / 64 (one row per thread, ~165 / ~250 registers) -> 3 / 2
        const int per_sm = c->mma ? 4 : (c->D == 16 && c->R == 1) ? 8 : (c->D == 64 ? 2 : 3);
        const int64_t slots = 148 * per_sm;
        want = cdiv(slots * 3, c->nb);
        double best = 1e30;
        int64_t best_s = want;
        for (int64_t s = (want + 1) / 2; s <= want * 2 && s <= max_splits; ++s) {
            const int64_t jc = cdiv(cdiv(n, s), c->JT) * c->JT;
            const int64_t real = cdiv(n, jc);
            const double ctas = (double)(c->nb * real);
            const double waves = (double)cdiv((int64_t)ctas, slots);
            const double cost = waves * (double)jc;     // time ~ waves x work per CTA
            if (cost < best - 1e-9) { best = cost; best_s = s; }
        }
        want = best_s;
    }
    if (want > max_splits) want = max_splits;
    if (want < 1) want = 1;
    c->j_chunk = cdiv(cdiv(n, want), c->JT) * c->JT;
    c->splits = (int)cdiv(n, c->j_chunk);
    const int rows_per_fin_block = 256 / (c->D / 4);
    c->fin_blocks = cdiv(n, rows_per_fin_block);
    return true;
}

template <int D, int R, bool LOSS, bool GRAD>
__global__ void __launch_bounds__(DEC_THREADS, (D == 16 && R == 1) ? 8 : 1)
dec_dense_kernel(const float *__restrict__ Zd, int64_t ldz, int64_t n, int d, int64_t j_chunk,
                 float *__restrict__ dz_part, double *__restrict__ loss_part) {
    constexpr int JT = 2048 / D;
    __shared__ __align__(16) float Zs[JT][D];
    __shared__ double red[DEC_THREADS / 32];
    const int tid = threadIdx.x;
    const int64_t i0 = (int64_t)blockIdx.x * DEC_THREADS * R;
    const int64_t jbeg = (int64_t)blockIdx.y * j_chunk;
    const int64_t jend = min(n, jbeg + j_chunk);

    // Packed fp32x2 FMAs (Blackwell FFMA2, __ffma2_rn): the dot product keeps an (even, odd)
    // pair of partial sums, the gradient accumulator is updated two columns per instruction --
    // half the issue slots of scalar FFMA for the 2d FMAs per pair.
    constexpr int D2 = D / 2;
    float2 zi[R][D2];
    float2 acc[R][D2];
    float lsum[R];
    bool valid[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int64_t i = i0 + (int64_t)r * DEC_THREADS + tid;
        valid[r] = i < n;
        lsum[r] = 0.f;
#pragma unroll
        for (int k = 0; k < D2; ++k) {
            zi[r][k].x = (valid[r] && 2 * k < d) ? __ldg(Zd + i * ldz + 2 * k) : 0.f;
            zi[r][k].y = (valid[r] && 2 * k + 1 < d) ? __ldg(Zd + i * ldz + 2 * k + 1) : 0.f;
            acc[r][k] = make_float2(0.f, 0.f);
        }
    }

    for (int64_t j0 = jbeg; j0 < jend; j0 += JT) {
        const int jcount = (int)min((int64_t)JT, jend - j0);
        __syncthreads();
        for (int t = tid; t < JT * D; t += DEC_THREADS) {
            const int jj = t / D, k = t % D;
            Zs[jj][k] = (jj < jcount && k < d) ? __ldg(Zd + (j0 + jj) * ldz + k) : 0.f;
        }
        __syncthreads();
#pragma unroll 2
        for (int jj = 0; jj < jcount; ++jj) {
            float2 zj[D2];
#pragma unroll
            for (int k4 = 0; k4 < D / 4; ++k4) {
                const float4 q = *reinterpret_cast<const float4 *>(&Zs[jj][k4 * 4]);
                zj[k4 * 2 + 0] = make_float2(q.x, q.y);
                zj[k4 * 2 + 1] = make_float2(q.z, q.w);
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float2 x2 = make_float2(0.f, 0.f);
#pragma unroll
                for (int k = 0; k < D2; ++k) x2 = __ffma2_rn(zi[r][k], zj[k], x2);
                const float x = x2.x + x2.y;
                float l, sg;
                softplus_parts(x, l, sg);
                if (LOSS) lsum[r] += fmaxf(x, 0.f) + l;
                if (GRAD) {
                    const float2 sg2 = make_float2(sg, sg);
#pragma unroll
                    for (int k = 0; k < D2; ++k) acc[r][k] = __ffma2_rn(sg2, zj[k], acc[r][k]);
                }
            }
        }
    }

    if (GRAD) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int64_t i = i0 + (int64_t)r * DEC_THREADS + tid;
            if (!valid[r]) continue;
            float4 *o = reinterpret_cast<float4 *>(dz_part + ((int64_t)blockIdx.y * n + i) * D);
#pragma unroll
            for (int k4 = 0; k4 < D / 4; ++k4)
                o[k4] = make_float4(acc[r][k4 * 2].x, acc[r][k4 * 2].y, acc[r][k4 * 2 + 1].x, acc[r][k4 * 2 + 1].y);
        }
    }
    if (LOSS) {
        double s = 0.0;
#pragma unroll
        for (int r = 0; r < R; ++r) s += valid[r] ? (double)lsum[r] : 0.0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if ((tid & 31) == 0) red[tid >> 5] = s;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < DEC_THREADS / 32; ++w) t += red[w];
            loss_part[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] = t;
        }
    }
}

// ---- tensor-core form of the dense pass (tuning "dec_mma", default for d <= 16) -----------------
// ncu on the SIMT kernel above (Pubmed shape, 0.70 ms): 16 FFMA2 per pair; one or two rows per
// thread, three or four CTAs per SM all take the same time -- the FP32 pipe sets it.  The pass is
// two GEMMs around an element-wise chain (attention-shaped): S = Z_I Z_J^T, then dZ_I = sigma(S) Z_J.
// Both run here as m16n8k8 TF32 MMAs in SPLIT PRECISION: every operand is hi + lo with hi its TF32
// rounding, the product is hi*hi + lo*hi + hi*lo with fp32 accumulation (~2^-21 relative, held to
// the same 1e-5 parity tolerance as the SIMT kernel by the tests).
//   * Lane (g, t) of a warp holds the pairs (rows g, g+8) x (keys 2t, 2t+1) of a 16 x 8 tile of S.
//     The contraction index of an MMA may be permuted freely as long as A and B agree, so with
//     slot t = key 2t and slot t+4 = key 2t+1 that accumulator fragment IS the A fragment of the
//     second GEMM: sigma goes from the softplus chain into the next MMA without a shuffle.  The same
//     trick on the embedding dimension makes a lane's four B values one LDS.128.
//   * The query rows' A fragments are split once and stay in registers for the whole launch; a key
//     tile is split once when it is staged in shared memory, key-major for S and dimension-major
//     for the gradient GEMM; sigma is split by masking its low 13 mantissa bits (lo is exact).
//   * Tensor cores round the fp32 accumulation towards zero: the gradient MMAs accumulate over one
//     128-key block only and are then added into ordinary fp32 registers.
//   * Loss: sum softplus(x) = sum max(x,0) + ln prod (1 + e^-|x|); a product of 64 factors in (1,2]
//     cannot overflow, so one lg2 per 64 pairs replaces one per pair.
// The dz_part / loss_part layouts are those of dec_dense_kernel; dec_finalize_kernel is shared.
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_m16n8k8_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <bool LOSS, bool GRAD>
__global__ void __launch_bounds__(DEC_THREADS, 4)
dec_dense_mma_kernel(const float *__restrict__ Zd, int64_t ldz, int64_t n, int d, int64_t j_chunk,
                      float *__restrict__ dz_part, double *__restrict__ loss_part) {
    constexpr int D = 16, MT = 2, JT = 128, JP = JT + 8, FOLD = 32;
    __shared__ __align__(16) uint32_t Zh[JT][D];    // key-major tf32 hi / lo: B operand of S = Z_I Z_J^T
    __shared__ __align__(16) uint32_t Zl[JT][D];
    __shared__ __align__(16) uint32_t ZhT[D][JP];   // dimension-major copies: B operand of P x Z_J
    __shared__ __align__(16) uint32_t ZlT[D][JP];   // (row stride 136 words: conflict-free 64-bit reads)
    __shared__ double red[DEC_THREADS / 32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int64_t i0 = (int64_t)blockIdx.x * (DEC_THREADS / 32 * 16 * MT) + warp * (16 * MT);
    const int64_t jbeg = (int64_t)blockIdx.y * j_chunk;
    const int64_t jend = min(n, jbeg + j_chunk);

    uint32_t ah[MT][2][4], al[MT][2][4];
    bool rv[MT][2];
#pragma unroll
    for (int m = 0; m < MT; ++m) {
        const int64_t r0 = i0 + m * 16 + g, r1 = r0 + 8;
        rv[m][0] = r0 < n;
        rv[m][1] = r1 < n;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const int c0 = 4 * t + 2 * ks, c1 = c0 + 1;
            const float x[4] = {(rv[m][0] && c0 < d) ? __ldg(Zd + r0 * ldz + c0) : 0.f,
                                (rv[m][1] && c0 < d) ? __ldg(Zd + r1 * ldz + c0) : 0.f,
                                (rv[m][0] && c1 < d) ? __ldg(Zd + r0 * ldz + c1) : 0.f,
                                (rv[m][1] && c1 < d) ? __ldg(Zd + r1 * ldz + c1) : 0.f};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                ah[m][ks][q] = to_tf32(x[q]);
                al[m][ks][q] = to_tf32(x[q] - __uint_as_float(ah[m][ks][q]));
            }
        }
    }
    // gradient: fp32 accumulators [m-tile][n-tile][fragment]; fragment e: (row g + 8*(e>>1), dim 8*nt + 2t + (e&1))
    float facc[MT][2][4];
    float msum[MT][2], lsum[MT][2], prod[MT][2];
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            msum[m][h] = 0.f;
            lsum[m][h] = 0.f;
            prod[m][h] = 1.f;
#pragma unroll
            for (int e = 0; e < 4; ++e) facc[m][h][e] = 0.f;
        }
    int tiles = 0, pads = 0;

    for (int64_t j0 = jbeg; j0 < jend; j0 += JT) {
        const int jcount = (int)min((int64_t)JT, jend - j0);
        __syncthreads();
        for (int idx = tid; idx < JT * D; idx += DEC_THREADS) {
            const int jj = idx / D, k = idx % D;
            const float x = (jj < jcount && k < d) ? __ldg(Zd + (j0 + jj) * ldz + k) : 0.f;
            const uint32_t hi = to_tf32(x);
            const uint32_t lo = to_tf32(x - __uint_as_float(hi));
            Zh[jj][k] = hi;
            Zl[jj][k] = lo;
            ZhT[k][jj] = hi;
            ZlT[k][jj] = lo;
        }
        __syncthreads();
        if (LOSS) pads += min(2, max(0, ((jcount + 7) & ~7) - jcount - (6 - 2 * t)));
        float ga[MT][2][4];
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) ga[m][nt][e] = 0.f;
        for (int jt = 0; jt < jcount; jt += 8) {
            const uint4 vh = *reinterpret_cast<const uint4 *>(&Zh[jt + g][4 * t]);
            const uint4 vl = *reinterpret_cast<const uint4 *>(&Zl[jt + g][4 * t]);
            float cs[MT][4], cb[MT][4];
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int e = 0; e < 4; ++e) cs[m][e] = cb[m][e] = 0.f;
#pragma unroll
            for (int m = 0; m < MT; ++m) mma_m16n8k8_tf32(cs[m], al[m][0], vh.x, vh.y);
#pragma unroll
            for (int m = 0; m < MT; ++m) mma_m16n8k8_tf32(cb[m], ah[m][0], vh.x, vh.y);
#pragma unroll
            for (int m = 0; m < MT; ++m) mma_m16n8k8_tf32(cs[m], ah[m][0], vl.x, vl.y);
#pragma unroll
            for (int m = 0; m < MT; ++m) mma_m16n8k8_tf32(cb[m], ah[m][1], vh.z, vh.w);
#pragma unroll
            for (int m = 0; m < MT; ++m) mma_m16n8k8_tf32(cs[m], al[m][1], vh.z, vh.w);
#pragma unroll
            for (int m = 0; m < MT; ++m) mma_m16n8k8_tf32(cs[m], ah[m][1], vl.z, vl.w);
            // B fragments of P x Z_J: (k-slot t, n = g) = Z_J[key 2t][dim 8 nt + g], slot t+4 = key 2t+1
            uint2 gh[2], gl[2];
            if (GRAD) {
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    gh[nt] = *reinterpret_cast<const uint2 *>(&ZhT[nt * 8 + g][jt + 2 * t]);
                    gl[nt] = *reinterpret_cast<const uint2 *>(&ZlT[nt * 8 + g][jt + 2 * t]);
                }
            }
            uint32_t ph[MT][4], pl[MT][4];
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                // e = 0: (row g, key 2t)  1: (row g, key 2t+1)  2: (row g+8, key 2t)  3: (row g+8, key 2t+1)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int h = e >> 1;
                    const float x = cs[m][e] + cb[m][e];
                    float ex, r;
                    const float a = -fabsf(x) * 1.4426950408889634f;
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(a));
                    const float one_e = 1.0f + ex;
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(one_e));
                    if (LOSS) {
                        msum[m][h] += fmaxf(x, 0.f);
                        prod[m][h] *= one_e;
                    }
                    if (GRAD) {
                        const float sg = (x >= 0.f) ? r : ex * r;
                        // A fragment order: a0 = e0, a1 = e2, a2 = e1, a3 = e3
                        const int ai = ((e & 1) << 1) | (e >> 1);
                        const uint32_t hi = __float_as_uint(sg) & 0xffffe000u;
                        ph[m][ai] = hi;
                        pl[m][ai] = __float_as_uint(sg - __uint_as_float(hi));
                    }
                }
            }
            if (GRAD) {
#pragma unroll
                for (int m = 0; m < MT; ++m)
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt) mma_m16n8k8_tf32(ga[m][nt], pl[m], gh[nt].x, gh[nt].y);
#pragma unroll
                for (int m = 0; m < MT; ++m)
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt) mma_m16n8k8_tf32(ga[m][nt], ph[m], gl[nt].x, gl[nt].y);
#pragma unroll
                for (int m = 0; m < MT; ++m)
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt) mma_m16n8k8_tf32(ga[m][nt], ph[m], gh[nt].x, gh[nt].y);
            }
            if (LOSS && (++tiles & (FOLD - 1)) == 0) {
#pragma unroll
                for (int m = 0; m < MT; ++m)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        float l2;
                        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(prod[m][h]));
                        lsum[m][h] += l2;
                        prod[m][h] = 1.f;
                    }
            }
        }
        if (GRAD) {
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                    for (int e = 0; e < 4; ++e) facc[m][nt][e] += ga[m][nt][e];
        }
    }

    if (GRAD) {
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (!rv[m][h]) continue;
                const int64_t i = i0 + m * 16 + h * 8 + g;
                float *o = dz_part + ((int64_t)blockIdx.y * n + i) * D + 2 * t;
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
                    *reinterpret_cast<float2 *>(o + nt * 8) = make_float2(facc[m][nt][2 * h], facc[m][nt][2 * h + 1]);
            }
    }
    if (LOSS) {
        double s = 0.0;
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float l2;
                asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(prod[m][h]));
                const double row = (double)msum[m][h] + ((double)(lsum[m][h] + l2) - (double)pads) * 0.6931471805599453;
                s += rv[m][h] ? row : 0.0;
            }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) red[warp] = s;
        __syncthreads();
        if (tid == 0) {
            double tot = 0.0;
            for (int w = 0; w < DEC_THREADS / 32; ++w) tot += red[w];
            loss_part[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] = tot;
        }
    }
}

// Tensor-core (symmetric-half) paths: the partial sums of one row -- `splits` G_I slots and one G_J slot per row block
// above the row's own -- folded into slot 0 of dz_part by ONE WARP PER ROW: lane = (slot group 0..7, float4 0..3), each
// lane adds every 8th slot in slot order with four loads in flight, the eight groups meet in a fixed xor tree.
// (dec_finalize_kernel used to walk the up to 155 slots from 4 lanes per row: 99 us at the Pubmed shape, all of it
// load latency.)  Deterministic; the row's slot 0 is read and written by the same lanes.
__global__ void __launch_bounds__(256) dec_slot_reduce_kernel(float *__restrict__ dz_part, int splits,
                                                              const float *__restrict__ dzT_part, int64_t n) {
    const int lane = threadIdx.x & 31, sub = lane & 3, sg = lane >> 2;
    const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= n) return;
    const int64_t step = n * 4;                                   // float4 per slot
    float4 g = f4_zero();
    const float4 *p = reinterpret_cast<const float4 *>(dz_part) + i * 4 + sub;
    for (int s = sg; s < splits; s += 8) f4_add(g, __ldcs(p + s * step));
    const float4 *q = reinterpret_cast<const float4 *>(dzT_part) + i * 4 + sub;
    const int nI = (int)(i / 128);
    int I = sg;
    for (; I + 24 < nI; I += 32) {
        float4 t[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) t[u] = __ldcs(q + (int64_t)(I + 8 * u) * step);
#pragma unroll
        for (int u = 0; u < 4; ++u) f4_add(g, t[u]);
    }
    for (; I < nI; I += 8) f4_add(g, __ldcs(q + (int64_t)I * step));
#pragma unroll
    for (int off = 4; off < 32; off <<= 1) f4_add(g, f4_shfl_xor(g, off));
    if (sg == 0) reinterpret_cast<float4 *>(dz_part)[i * 4 + sub] = g;
}

// One lane group (D/4 lanes) per row: ordered split reduce + per-edge corrections.
template <int D>
__global__ void __launch_bounds__(256)
dec_finalize_kernel(const float *__restrict__ Zd, int64_t ldz, int64_t n, int d,
                    const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                    const int64_t *__restrict__ rowptr_t, const int32_t *__restrict__ col_t, float pw,
                    const float *__restrict__ dz_part, int splits, int mode, float inv_n2,
                    float *__restrict__ dZ, int64_t ld_dz, double *__restrict__ loss_part,
                    const int64_t *__restrict__ blk_lo, const int64_t *__restrict__ blk_hi) {
    constexpr int LPR = D / 4;
    constexpr int RPB = 256 / LPR;
    __shared__ double red[8];
    const int tid = threadIdx.x;
    const int sub = tid % LPR;
    const int64_t i = (int64_t)blockIdx.x * RPB + tid / LPR;
    const bool valid = i < n;
    const bool want_loss = mode & GAE_DEC_LOSS, want_grad = mode & GAE_DEC_GRAD;

    auto load_row = [&](int64_t r) -> float4 {
        float4 v = f4_zero();
        const float *p = Zd + r * ldz + sub * 4;
        if (sub * 4 + 0 < d) v.x = __ldg(p + 0);
        if (sub * 4 + 1 < d) v.y = __ldg(p + 1);
        if (sub * 4 + 2 < d) v.z = __ldg(p + 2);
        if (sub * 4 + 3 < d) v.w = __ldg(p + 3);
        return v;
    };
    auto group_dot = [&](const float4 &a, const float4 &b) -> float {
        float x = a.x * b.x;
        x = fmaf(a.y, b.y, x); x = fmaf(a.z, b.z, x); x = fmaf(a.w, b.w, x);
#pragma unroll
        for (int off = 1; off < LPR; off <<= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
        return x;
    };

    const float4 zi = valid ? load_row(i) : f4_zero();
    float4 g = f4_zero();
    if (want_grad && valid && !blk_lo) {
        for (int s = 0; s < splits; ++s)
            f4_add(g, *reinterpret_cast<const float4 *>(dz_part + ((int64_t)s * n + i) * D + sub * 4));
        // x_ij = x_ji: (G + G^T) Zd doubles the dense term
        g.x *= 2.f; g.y *= 2.f; g.z *= 2.f; g.w *= 2.f;
    }
    float lsum = 0.f;
    if (blk_lo) {
        // block-diagonal variant (per-graph decoder): the dense term runs
        // over the row's own graph only, j in [blk_lo[i], blk_hi[i]); no dense pass was launched
        const int64_t j0 = valid ? blk_lo[i] : 0, j1 = valid ? blk_hi[i] : 0;
        int64_t len = j1 - j0, maxlen = len;
#pragma unroll
        for (int off = LPR; off < 32; off <<= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, off));
        for (int64_t t = 0; t < maxlen; ++t) {
            const bool on = t < len;
            const float4 zj = on ? load_row(j0 + t) : f4_zero();
            const float x = group_dot(zi, zj);
            float l, sg;
            softplus_parts(x, l, sg);
            if (on) {
                lsum += fmaxf(x, 0.f) + l;
                f4_fma(g, 2.f * sg, zj);            // (G + G^T) Zd with x_ij = x_ji
            }
        }
    }
    // groups of one warp walk different rows: loop to the warp-wide max degree so the
    // shuffles inside group_dot stay convergent
    {
        const int64_t e0 = valid ? rowptr[i] : 0, e1 = valid ? rowptr[i + 1] : 0;
        int64_t len = e1 - e0, maxlen = len;
#pragma unroll
        for (int off = LPR; off < 32; off <<= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, off));
        for (int64_t t = 0; t < maxlen; ++t) {
            const bool on = t < len;
            const float4 zj = on ? load_row(col[e0 + t]) : f4_zero();
            const float x = group_dot(zi, zj);
            float l, sg;
            softplus_parts(x, l, sg);
            if (on) {
                lsum += pw * (fmaxf(-x, 0.f) + l) - (fmaxf(x, 0.f) + l);
                const float c = -(pw * (1.f - sg) + sg);
                f4_fma(g, c, zj);
            }
        }
    }
    if (want_grad) {
        const int64_t e0 = valid ? rowptr_t[i] : 0, e1 = valid ? rowptr_t[i + 1] : 0;
        int64_t len = e1 - e0, maxlen = len;
#pragma unroll
        for (int off = LPR; off < 32; off <<= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, off));
        for (int64_t t = 0; t < maxlen; ++t) {
            const bool on = t < len;
            const float4 zj = on ? load_row(col_t[e0 + t]) : f4_zero();
            const float x = group_dot(zi, zj);
            float l, sg;
            softplus_parts(x, l, sg);
            if (on) {
                const float c = -(pw * (1.f - sg) + sg);
                f4_fma(g, c, zj);
            }
        }
        if (valid) {
            float *o = dZ + i * ld_dz + sub * 4;
            if (sub * 4 + 0 < d) o[0] = g.x * inv_n2;
            if (sub * 4 + 1 < d) o[1] = g.y * inv_n2;
            if (sub * 4 + 2 < d) o[2] = g.z * inv_n2;
            if (sub * 4 + 3 < d) o[3] = g.w * inv_n2;
        }
    }
    if (want_loss) {
        double s = (valid && sub == 0) ? (double)lsum : 0.0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if ((tid & 31) == 0) red[tid >> 5] = s;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < 8; ++w) t += red[w];
            loss_part[blockIdx.x] = t;
        }
    }
}

// The same finalize step with ONE WARP PER ROW (the full-matrix decoder; the kernel above stays for the block-diagonal
// variant): the warp's 32 / LPR lane groups take every G-th partial slot and every G-th edge of the row, and meet in a
// fixed xor tree at the end -- deterministic, and a row's dependent chain (col -> Zd row -> dot -> MUFU) is deg / G
// long instead of deg, with 8x (D = 16) the warps in flight.  Measured at the Pubmed shape: 82 us -> see profiles/.
template <int D>
__global__ void __launch_bounds__(256)
dec_finalize_rows_kernel(const float *__restrict__ Zd, int64_t ldz, int64_t n, int d,
                         const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                         const int64_t *__restrict__ rowptr_t, const int32_t *__restrict__ col_t, float pw,
                         const float *__restrict__ dz_part, int splits, int mode, float inv_n2,
                         float *__restrict__ dZ, int64_t ld_dz, double *__restrict__ loss_part) {
    constexpr int LPR = D / 4;
    constexpr int G = 32 / LPR;
    __shared__ double red[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sub = lane % LPR, grp = lane / LPR;
    const int64_t i = (int64_t)blockIdx.x * 8 + warp;
    const bool valid = i < n;                                     // warp-uniform
    const bool want_loss = mode & GAE_DEC_LOSS, want_grad = mode & GAE_DEC_GRAD;

    auto load_row = [&](int64_t r) -> float4 {
        float4 v = f4_zero();
        const float *p = Zd + r * ldz + sub * 4;
        if (sub * 4 + 0 < d) v.x = __ldg(p + 0);
        if (sub * 4 + 1 < d) v.y = __ldg(p + 1);
        if (sub * 4 + 2 < d) v.z = __ldg(p + 2);
        if (sub * 4 + 3 < d) v.w = __ldg(p + 3);
        return v;
    };
    auto group_dot = [&](const float4 &a, const float4 &b) -> float {
        float x = a.x * b.x;
        x = fmaf(a.y, b.y, x); x = fmaf(a.z, b.z, x); x = fmaf(a.w, b.w, x);
#pragma unroll
        for (int off = 1; off < LPR; off <<= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
        return x;
    };

    float4 g = f4_zero();
    float lsum = 0.f;
    if (valid) {
        const float4 zi = load_row(i);
        if (want_grad) {
            for (int s = grp; s < splits; s += G)
                f4_add(g, *reinterpret_cast<const float4 *>(dz_part + ((int64_t)s * n + i) * D + sub * 4));
            // x_ij = x_ji: (G + G^T) Zd doubles the dense term
            g.x *= 2.f; g.y *= 2.f; g.z *= 2.f; g.w *= 2.f;
        }
        {
            const int64_t e0 = rowptr[i], len = rowptr[i + 1] - e0;
            for (int64_t t = grp; t < len + grp; t += G) {      // same trip count for every group: the shuffles stay convergent
                const bool on = t < len;
                const float4 zj = on ? load_row(col[e0 + t]) : f4_zero();
                const float x = group_dot(zi, zj);
                float l, sg;
                softplus_parts(x, l, sg);
                if (on) {
                    lsum += pw * (fmaxf(-x, 0.f) + l) - (fmaxf(x, 0.f) + l);
                    f4_fma(g, -(pw * (1.f - sg) + sg), zj);
                }
            }
        }
        if (want_grad) {
            const int64_t e0 = rowptr_t[i], len = rowptr_t[i + 1] - e0;
            for (int64_t t = grp; t < len + grp; t += G) {
                const bool on = t < len;
                const float4 zj = on ? load_row(col_t[e0 + t]) : f4_zero();
                const float x = group_dot(zi, zj);
                float l, sg;
                softplus_parts(x, l, sg);
                if (on) f4_fma(g, -(pw * (1.f - sg) + sg), zj);
            }
#pragma unroll
            for (int off = LPR; off < 32; off <<= 1) f4_add(g, f4_shfl_xor(g, off));
            if (grp == 0) {
                float *o = dZ + i * ld_dz + sub * 4;
                if (sub * 4 + 0 < d) o[0] = g.x * inv_n2;
                if (sub * 4 + 1 < d) o[1] = g.y * inv_n2;
                if (sub * 4 + 2 < d) o[2] = g.z * inv_n2;
                if (sub * 4 + 3 < d) o[3] = g.w * inv_n2;
            }
        }
    }
    if (want_loss) {
        double s = (valid && sub == 0) ? (double)lsum : 0.0;     // one lane per group holds the group's edges
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) red[warp] = s;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < 8; ++w) t += red[w];
            loss_part[blockIdx.x] = t;
        }
    }
}

// loss = (sum dense partials + sum edge partials) / N^2, fixed order, fp64 accumulate
__global__ void dec_loss_reduce_kernel(const double *__restrict__ part, int64_t count, double inv_n2,
                                       float *__restrict__ loss) {
    __shared__ double red[256];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < count; i += 256) s += part[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) *loss = (float)(red[0] * inv_n2);
}

// Materialised logits X = Zd Zd^T (compatibility path: GAE.forward returns [N,N]).
__global__ void __launch_bounds__(256)
dec_logits_kernel(const float *__restrict__ Zd, int64_t ldz, int64_t n, int d, float *__restrict__ X, int64_t ldx) {
    constexpr int T = 64, KC = 16;
    __shared__ float Zi[KC][T + 1], Zj[KC][T + 1];
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int64_t i0 = (int64_t)blockIdx.y * T, j0 = (int64_t)blockIdx.x * T;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < d; k0 += KC) {
        __syncthreads();
        for (int t = tid; t < T * KC; t += 256) {
            const int r = t / KC, k = t % KC;
            Zi[k][r] = (i0 + r < n && k0 + k < d) ? __ldg(Zd + (i0 + r) * ldz + k0 + k) : 0.f;
            Zj[k][r] = (j0 + r < n && k0 + k < d) ? __ldg(Zd + (j0 + r) * ldz + k0 + k) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { a[q] = Zi[k][ty * 4 + q]; b[q] = Zj[k][tx + 16 * q]; }
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[p][q] = fmaf(a[p], b[q], acc[p][q]);
        }
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int64_t i = i0 + ty * 4 + p;
        if (i >= n) continue;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int64_t j = j0 + tx + 16 * q;
            if (j < n) X[i * ldx + j] = acc[p][q];
        }
    }
}

static cudaError_t launch_dense_mma(const DecConfig &c, int mode, const float *Zd, int64_t ldz, int64_t n, int d,
                                    float *dz_part, double *loss_part, cudaStream_t st) {
    dim3 grid((unsigned)c.nb, (unsigned)c.splits);
    const bool L = mode & GAE_DEC_LOSS, G = mode & GAE_DEC_GRAD;
    if (L && G) dec_dense_mma_kernel<true, true><<<grid, DEC_THREADS, 0, st>>>(Zd, ldz, n, d, c.j_chunk, dz_part, loss_part);
    else if (L) dec_dense_mma_kernel<true, false><<<grid, DEC_THREADS, 0, st>>>(Zd, ldz, n, d, c.j_chunk, dz_part, loss_part);
    else dec_dense_mma_kernel<false, true><<<grid, DEC_THREADS, 0, st>>>(Zd, ldz, n, d, c.j_chunk, dz_part, loss_part);
    count_launch();
    return cudaGetLastError();
}

template <int D, int R>
static cudaError_t launch_dense(const DecConfig &c, int mode, const float *Zd, int64_t ldz, int64_t n, int d,
                                float *dz_part, double *loss_part, cudaStream_t st) {
    dim3 grid((unsigned)c.nb, (unsigned)c.splits);
    const bool L = mode & GAE_DEC_LOSS, G = mode & GAE_DEC_GRAD;
    if (L && G) dec_dense_kernel<D, R, true, true><<<grid, DEC_THREADS, 0, st>>>(Zd, ldz, n, d, c.j_chunk, dz_part, loss_part);
    else if (L) dec_dense_kernel<D, R, true, false><<<grid, DEC_THREADS, 0, st>>>(Zd, ldz, n, d, c.j_chunk, dz_part, loss_part);
    else dec_dense_kernel<D, R, false, true><<<grid, DEC_THREADS, 0, st>>>(Zd, ldz, n, d, c.j_chunk, dz_part, loss_part);
    count_launch();
    return cudaGetLastError();
}

}  // namespace gae

using namespace gae;

static int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

extern "C" int64_t gae_decoder_ws_bytes(int64_t n, int32_t d) {
    DecConfig c;
    if (n <= 0 || d <= 0 || !dec_config(n, d, &c)) return 0;
    const int64_t dz = align_up((int64_t)sizeof(float) * c.splits * n * c.D, 256);
    const int64_t lp = (int64_t)sizeof(double) * (c.nb * c.splits + cdiv(n, 8));     // dense partials + one per finalize block
    const int64_t tc = c.tc ? align_up((int64_t)sizeof(float) * c.nb * n * c.D, 256) + 256 : 0;
    return dz + align_up(lp, 256) + tc;
}

extern "C" int gae_decoder_bce_f32(const float *Zd, int64_t ldz, int64_t n, int32_t d,
                                   const int64_t *rowptr, const int32_t *col, const int64_t *rowptr_t,
                                   const int32_t *col_t, float pos_weight, int32_t mode, float *loss,
                                   float *dZd_unit, int64_t ld_dz, void *ws, int64_t ws_bytes, void *stream) {
    GAE_CHECK_ARG(n > 0 && d > 0, "n, d must be > 0");
    GAE_CHECK_ARG(Zd && rowptr, "null pointer");
    GAE_CHECK_ARG(ldz >= d, "ldz too small");
    GAE_CHECK_ARG((mode & ~3) == 0 && mode != 0, "mode must be a combination of GAE_DEC_LOSS|GAE_DEC_GRAD");
    const bool want_loss = mode & GAE_DEC_LOSS, want_grad = mode & GAE_DEC_GRAD;
    GAE_CHECK_ARG(!want_loss || loss, "loss pointer required");
    GAE_CHECK_ARG(!want_grad || (dZd_unit && rowptr_t && ld_dz >= d), "gradient needs dZd_unit and CSR(A^T)");
    DecConfig c;
    if (!dec_config(n, d, &c)) {
        set_error("decoder supports embedding width d <= 64 (got %d)", d);
        return GAE_ERR_UNSUPPORTED;
    }
    if (!ws || ws_bytes < gae_decoder_ws_bytes(n, d)) {
        set_error("decoder workspace too small: have %lld need %lld", (long long)ws_bytes,
                  (long long)gae_decoder_ws_bytes(n, d));
        return GAE_ERR_WORKSPACE;
    }
    GAE_CHECK_ARG(aligned16(ws), "workspace must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    float *dz_part = (float *)ws;
    double *loss_part = (double *)((char *)ws + align_up((int64_t)sizeof(float) * c.splits * n * c.D, 256));
    double *loss_part_edges = loss_part + c.nb * c.splits;
    float *dzT_part = nullptr;
    const int64_t fin_rows_blocks = cdiv(n, 8);          // dec_finalize_rows_kernel: a warp per row, 8 rows per block
    if (c.tc) {
        char *p = (char *)loss_part + align_up((int64_t)sizeof(double) * (c.nb * c.splits + fin_rows_blocks), 256);
        uint32_t *err = (uint32_t *)p;
        dzT_part = (float *)(p + 256);
        GAE_CUDA(cudaMemsetAsync(err, 0, 2 * sizeof(uint32_t), st));      // expiry counter, max |Zd| bits
        if (c.tc_variant == 2) GAE_CUDA(dec_tc16_launch(Zd, ldz, n, d, c.splits, dz_part, dzT_part, loss_part, err, st));
        else GAE_CUDA(dec_tc_launch(Zd, ldz, n, d, c.splits, dz_part, dzT_part, loss_part, err, st));
    } else
    if (c.mma) GAE_CUDA(launch_dense_mma(c, mode, Zd, ldz, n, d, dz_part, loss_part, st));
    else if (c.D == 16 && c.R == 1) GAE_CUDA((launch_dense<16, 1>(c, mode, Zd, ldz, n, d, dz_part, loss_part, st)));
    else if (c.D == 16) GAE_CUDA((launch_dense<16, 2>(c, mode, Zd, ldz, n, d, dz_part, loss_part, st)));
    else if (c.D == 32) GAE_CUDA((launch_dense<32, 1>(c, mode, Zd, ldz, n, d, dz_part, loss_part, st)));
    else GAE_CUDA((launch_dense<64, 1>(c, mode, Zd, ldz, n, d, dz_part, loss_part, st)));

    const float inv_n2 = (float)(1.0 / ((double)n * (double)n));
    int fin_splits = c.splits;
    if (c.tc && want_grad) {      // fold the partial slots first (one warp per row); finalize then reads slot 0
        dec_slot_reduce_kernel<<<(unsigned)cdiv(n, 8), 256, 0, st>>>(dz_part, c.splits, dzT_part, n);
        GAE_LAUNCH_CHECK();
        fin_splits = 1;
    }
    if (c.D == 16)
        dec_finalize_rows_kernel<16><<<(unsigned)fin_rows_blocks, 256, 0, st>>>(Zd, ldz, n, d, rowptr, col, rowptr_t, col_t, pos_weight, dz_part,
                                                                                fin_splits, mode, inv_n2, dZd_unit, ld_dz, loss_part_edges);
    else if (c.D == 32)
        dec_finalize_rows_kernel<32><<<(unsigned)fin_rows_blocks, 256, 0, st>>>(Zd, ldz, n, d, rowptr, col, rowptr_t, col_t, pos_weight, dz_part,
                                                                                fin_splits, mode, inv_n2, dZd_unit, ld_dz, loss_part_edges);
    else
        dec_finalize_rows_kernel<64><<<(unsigned)fin_rows_blocks, 256, 0, st>>>(Zd, ldz, n, d, rowptr, col, rowptr_t, col_t, pos_weight, dz_part,
                                                                                fin_splits, mode, inv_n2, dZd_unit, ld_dz, loss_part_edges);
    GAE_LAUNCH_CHECK();
    if (want_loss) {
        dec_loss_reduce_kernel<<<1, 256, 0, st>>>(loss_part, c.nb * c.splits + fin_rows_blocks,
                                                  1.0 / ((double)n * (double)n), loss);
        GAE_LAUNCH_CHECK();
    }
    return GAE_OK;
}

extern "C" int64_t gae_decoder_blockdiag_ws_bytes(int64_t n, int32_t d) {
    DecConfig c;
    if (n <= 0 || d <= 0 || !dec_config(n, d, &c)) return 0;
    return align_up((int64_t)sizeof(double) * c.fin_blocks, 256);
}

extern "C" int gae_decoder_bce_blockdiag_f32(const float *Zd, int64_t ldz, int64_t n, int32_t d, const int64_t *rowptr,
                                             const int32_t *col, const int64_t *rowptr_t, const int32_t *col_t,
                                             const int64_t *blk_lo, const int64_t *blk_hi, double n_pairs,
                                             float pos_weight, int32_t mode, float *loss, float *dZd_unit,
                                             int64_t ld_dz, void *ws, int64_t ws_bytes, void *stream) {
    GAE_CHECK_ARG(n > 0 && d > 0 && n_pairs > 0, "n, d, n_pairs must be > 0");
    GAE_CHECK_ARG(Zd && rowptr && blk_lo && blk_hi, "null pointer");
    GAE_CHECK_ARG(ldz >= d, "ldz too small");
    GAE_CHECK_ARG((mode & ~3) == 0 && mode != 0, "mode must be a combination of GAE_DEC_LOSS|GAE_DEC_GRAD");
    const bool want_loss = mode & GAE_DEC_LOSS, want_grad = mode & GAE_DEC_GRAD;
    GAE_CHECK_ARG(!want_loss || loss, "loss pointer required");
    GAE_CHECK_ARG(!want_grad || (dZd_unit && rowptr_t && ld_dz >= d), "gradient needs dZd_unit and CSR(A^T)");
    DecConfig c;
    if (!dec_config(n, d, &c)) {
        set_error("decoder supports embedding width d <= 64 (got %d)", d);
        return GAE_ERR_UNSUPPORTED;
    }
    if (!ws || ws_bytes < gae_decoder_blockdiag_ws_bytes(n, d)) {
        set_error("decoder workspace too small");
        return GAE_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    double *loss_part = (double *)ws;
    const float inv = (float)(1.0 / n_pairs);
    if (c.D == 16)
        dec_finalize_kernel<16><<<(unsigned)c.fin_blocks, 256, 0, st>>>(Zd, ldz, n, d, rowptr, col, rowptr_t, col_t, pos_weight,
                                                                        nullptr, 0, mode, inv, dZd_unit, ld_dz, loss_part, blk_lo, blk_hi);
    else if (c.D == 32)
        dec_finalize_kernel<32><<<(unsigned)c.fin_blocks, 256, 0, st>>>(Zd, ldz, n, d, rowptr, col, rowptr_t, col_t, pos_weight,
                                                                        nullptr, 0, mode, inv, dZd_unit, ld_dz, loss_part, blk_lo, blk_hi);
    else
        dec_finalize_kernel<64><<<(unsigned)c.fin_blocks, 256, 0, st>>>(Zd, ldz, n, d, rowptr, col, rowptr_t, col_t, pos_weight,
                                                                        nullptr, 0, mode, inv, dZd_unit, ld_dz, loss_part, blk_lo, blk_hi);
    GAE_LAUNCH_CHECK();
    if (want_loss) {
        dec_loss_reduce_kernel<<<1, 256, 0, st>>>(loss_part, c.fin_blocks, 1.0 / n_pairs, loss);
        GAE_LAUNCH_CHECK();
    }
    return GAE_OK;
}

extern "C" int gae_decoder_logits_f32(const float *Zd, int64_t ldz, int64_t n, int32_t d, float *X, int64_t ldx,
                                      void *stream) {
    GAE_CHECK_ARG(n >= 0 && d > 0, "bad sizes");
    if (n == 0) return GAE_OK;
    GAE_CHECK_ARG(Zd && X && ldz >= d && ldx >= n, "bad pointers / leading dimensions");
    dim3 grid((unsigned)cdiv(n, 64), (unsigned)cdiv(n, 64));
    dec_logits_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(Zd, ldz, n, d, X, ldx);
    GAE_LAUNCH_CHECK();
    return GAE_OK;
}
