// K3 -- NodeApplyModule: H = act(Yin W^T + b) and its adjoint (reference gae.py:7-16).
//
// The dense products here are skinny (d_out = 32 / 16, K = d_in) fp32 GEMMs that must match
// torch's fp32 nn.Linear to 1e-5, so they run as SIMT FFMA tiles (TF32 tensor cores would
// break the tolerance; the work is <1 % of a step).  One strided,
// shared-memory-tiled kernel serves the three products:
//   fwd : H[n,do]    = Yin[n,di]   . W^T          (+ bias, activation)
//   dY  : dYin[n,di] = dPre[n,do]  . W
//   dW  : dW[do,di]  = dPre^T      . Yin           (split over n, ordered reduce)
// with dPre = dH * (H > 0) applied on the fly when the layer has a ReLU.
#include "common.cuh"

namespace gae {

struct GemmArgs {
    const float *A; int64_t a_rs, a_cs;      // A(m,k) = A[m*a_rs + k*a_cs]
    const float *Mask; int64_t m_rs, m_cs;   // optional: A(m,k) *= (Mask(m,k) > 0)
    const float *B; int64_t b_rs, b_cs;      // B(k,n) = B[k*b_rs + n*b_cs]
    float *C; int64_t ldc;                   // C[m*ldc + n] ; split z adds z*c_split_stride
    int64_t c_split_stride;
    const float *bias;                       // optional [N]
    int64_t M, N, K;
    int64_t k_chunk;                         // K range per blockIdx.z
    int act;
};

constexpr int BK = 16, GEMM_THREADS = 128;

template <int BM, int BN, bool A_KC, bool B_KC>
__global__ void __launch_bounds__(GEMM_THREADS) sgemm_kernel(const GemmArgs g) {
    constexpr int TM = BM / 16, TN = BN / 8;
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int ty = tid / 8, tx = tid % 8;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int64_t n0 = (int64_t)blockIdx.y * BN;
    const int64_t kbeg = (int64_t)blockIdx.z * g.k_chunk;
    const int64_t kend = min(g.K, kbeg + g.k_chunk);

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    // register double buffering: the global loads of tile t+1 are issued before the FMAs of tile t
    constexpr int NA = BM * BK / GEMM_THREADS, NB = BN * BK / GEMM_THREADS;
    float ra[NA], rb[NB];
    auto fetch = [&](int64_t k0) {
#pragma unroll
        for (int q = 0; q < NA; ++q) {
            const int i = tid + q * GEMM_THREADS;
            int mm, kk;
            if (A_KC) { kk = i % BK; mm = i / BK; } else { mm = i % BM; kk = i / BM; }
            const int64_t m = m0 + mm, k = k0 + kk;
            float v = 0.f;
            if (m < g.M && k < kend) {
                v = __ldg(g.A + m * g.a_rs + k * g.a_cs);
                if (g.Mask && !(__ldg(g.Mask + m * g.m_rs + k * g.m_cs) > 0.f)) v = 0.f;
            }
            ra[q] = v;
        }
#pragma unroll
        for (int q = 0; q < NB; ++q) {
            const int i = tid + q * GEMM_THREADS;
            int nn, kk;
            if (B_KC) { kk = i % BK; nn = i / BK; } else { nn = i % BN; kk = i / BN; }
            const int64_t n = n0 + nn, k = k0 + kk;
            rb[q] = (n < g.N && k < kend) ? __ldg(g.B + k * g.b_rs + n * g.b_cs) : 0.f;
        }
    };
    auto stash = [&]() {
#pragma unroll
        for (int q = 0; q < NA; ++q) {
            const int i = tid + q * GEMM_THREADS;
            int mm, kk;
            if (A_KC) { kk = i % BK; mm = i / BK; } else { mm = i % BM; kk = i / BM; }
            As[kk][mm] = ra[q];
        }
#pragma unroll
        for (int q = 0; q < NB; ++q) {
            const int i = tid + q * GEMM_THREADS;
            int nn, kk;
            if (B_KC) { kk = i % BK; nn = i / BK; } else { nn = i % BN; kk = i / BN; }
            Bs[kk][nn] = rb[q];
        }
    };
    if (kbeg < kend) fetch(kbeg);
    for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
        stash();
        __syncthreads();
        if (k0 + BK < kend) fetch(k0 + BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    float *C = g.C + (int64_t)blockIdx.z * g.c_split_stride;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int64_t m = m0 + ty * TM + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int64_t n = n0 + tx * TN + j;
            if (n >= g.N) continue;
            float v = acc[i][j];
            if (g.bias) v += __ldg(g.bias + n);
            if (g.act == GAE_ACT_RELU) v = fmaxf(v, 0.f);
            C[m * g.ldc + n] = v;
        }
    }
}

template <int BM, bool A_KC, bool B_KC>
static cudaError_t launch_gemm_bm(const GemmArgs &g, int splits, cudaStream_t st) {
    if (g.N <= 16) {
        dim3 grid((unsigned)cdiv(g.M, BM), (unsigned)cdiv(g.N, 16), splits);
        sgemm_kernel<BM, 16, A_KC, B_KC><<<grid, GEMM_THREADS, 0, st>>>(g);
    } else if (g.N <= 32) {
        dim3 grid((unsigned)cdiv(g.M, BM), (unsigned)cdiv(g.N, 32), splits);
        sgemm_kernel<BM, 32, A_KC, B_KC><<<grid, GEMM_THREADS, 0, st>>>(g);
    } else {
        dim3 grid((unsigned)cdiv(g.M, BM), (unsigned)cdiv(g.N, 64), splits);
        sgemm_kernel<BM, 64, A_KC, B_KC><<<grid, GEMM_THREADS, 0, st>>>(g);
    }
    count_launch();
    return cudaGetLastError();
}

// ---- fast paths for the two products that read the [n, d_in] activation matrix (round 2) ----------------------------
// At the Pubmed shape (n = 19 717, d_in = 500, d_out = 32) the generic kernel above took 88 us for the forward product
// and 88 + 13 + 13 us for dW / db -- a third of the train step once the decoder had shrunk -- at 0.45 TB/s and 7 TFLOP/s:
// scalar global loads, 8 scalar LDS per 8 FMAs.  These two kernels keep the FP32 FFMA arithmetic (nn.Linear parity)
// but move 128-bit everywhere: float4 global loads along the contiguous dimension, k-major shared tiles read back as
// one LDS.128 per operand per k for a 4 x TN register tile (2 LDS per 16 FMAs), an XOR swizzle on the transposed store
// of the activation tile (conflict-free both ways).  Requirements (else the generic kernel runs): 16-byte aligned
// rows, row strides multiples of 4 floats.

constexpr int FBM = 64, FBK = 32, F_THREADS = 128;
constexpr int FW_THREADS = 256;                        // forward kernel: two groups of 128 threads split each k tile between them

// H[M, N] = act(A[M, K] W[N, K]^T + b), A rows K-contiguous (lda), W rows K-contiguous (stride K).  BN = 16 / 32 / 64.
template <int BN>
__global__ void __launch_bounds__(FW_THREADS) linear_fwd_tn_kernel(const float *__restrict__ A, int64_t lda, const float *__restrict__ W,
                                                                    const float *__restrict__ bias, float *__restrict__ C, int64_t ldc,
                                                                    int64_t M, int N, int K, int act) {
    constexpr int TN = BN / 8;                          // 2 / 4 / 8 output columns per thread, 4 rows
    __shared__ __align__(16) float As[FBK][FBM];        // [k][row], float4 groups of 4 rows XOR-swizzled by (k / 4) % 8
    __shared__ __align__(16) float Bs[FBK][BN];         // [k][n], float4 groups of 4 columns XOR-swizzled by (k / 4) % (BN / 4)
    __shared__ float red[F_THREADS][4 * TN + 1];        // second group's accumulators on their way to the first
    // Two groups of 128 threads own the same 64 x BN output tile and take k 0..15 / 16..31 of every k tile: twice the
    // warps per tile to cover the latency of the next tile's global loads (the grid is only ~2 CTAs per SM at n = 2e4).
    const int tid = threadIdx.x, grp = tid >> 7, t2 = tid & 127, ty = t2 >> 3, tx = t2 & 7;
    const int64_t m0 = (int64_t)blockIdx.x * FBM;
    const int n0 = blockIdx.y * BN;
    float acc[4][TN];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    constexpr int NB = BN * FBK / FW_THREADS;           // W elements per thread per tile
    float4 ra[2];
    float rb[NB];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int f = tid + FW_THREADS * j, row = f >> 3, k = k0 + 4 * (f & 7);
            const int64_t m = m0 + row;
            float4 v = f4_zero();
            if (m < M) {
                const float *p = A + m * lda + k;
                if (k + 3 < K) v = __ldg(reinterpret_cast<const float4 *>(p));
                else {
                    if (k < K) v.x = __ldg(p);
                    if (k + 1 < K) v.y = __ldg(p + 1);
                    if (k + 2 < K) v.z = __ldg(p + 2);
                }
            }
            ra[j] = v;
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) {              // a warp reads 32 consecutive k of one W row: one line per load
            const int f = tid + FW_THREADS * j, k = k0 + (f & 31), n = f >> 5;
            rb[j] = (n0 + n < N && k < K) ? __ldg(W + (int64_t)(n0 + n) * K + k) : 0.f;
        }
    };
    auto stash = [&]() {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int f = tid + FW_THREADS * j, row = f >> 3, kq = f & 7;
            const int col = 4 * ((row >> 2) ^ kq) + (row & 3);
            As[4 * kq + 0][col] = ra[j].x;
            As[4 * kq + 1][col] = ra[j].y;
            As[4 * kq + 2][col] = ra[j].z;
            As[4 * kq + 3][col] = ra[j].w;
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            const int f = tid + FW_THREADS * j, kk = f & 31, n = f >> 5;
            Bs[kk][4 * ((n >> 2) ^ ((kk >> 2) & (BN / 4 - 1))) + (n & 3)] = rb[j];
        }
    };
    fetch(0);
    for (int k0 = 0; k0 < K; k0 += FBK) {
        stash();
        __syncthreads();
        if (k0 + FBK < K) fetch(k0 + FBK);
#pragma unroll
        for (int kh = 0; kh < FBK / 2; ++kh) {
            const int kk = kh + (FBK / 2) * grp;
            const float4 a = *reinterpret_cast<const float4 *>(&As[kk][4 * (ty ^ ((kk >> 2) & 7))]);
            float b[TN];
            constexpr int SW = BN / 4 - 1;
            if (TN == 2) {                              // columns 2 tx, 2 tx + 1: half of float4 group tx / 2
                const float2 t = *reinterpret_cast<const float2 *>(&Bs[kk][4 * ((tx >> 1) ^ ((kk >> 2) & SW)) + 2 * (tx & 1)]);
                b[0] = t.x; b[1] = t.y;
            } else {
#pragma unroll
                for (int j4 = 0; j4 < TN / 4; ++j4) {
                    const float4 t = *reinterpret_cast<const float4 *>(&Bs[kk][4 * ((tx * (TN / 4) + j4) ^ ((kk >> 2) & SW))]);
                    b[4 * j4] = t.x; b[4 * j4 + 1] = t.y; b[4 * j4 + 2] = t.z; b[4 * j4 + 3] = t.w;
                }
            }
            const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    if (grp == 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) red[t2][i * TN + j] = acc[i][j];
    }
    __syncthreads();
    if (grp == 1) return;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t m = m0 + 4 * ty + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + tx * TN + j;
            if (n >= N) continue;
            float v = acc[i][j] + red[t2][i * TN + j];
            if (bias) v += __ldg(bias + n);
            if (act == GAE_ACT_RELU) v = fmaxf(v, 0.f);
            C[m * ldc + n] = v;
        }
    }
}

// Partial dW and db over one chunk of rows:  part_w[z][o][c] = sum_{rows of chunk z} dPre[row][o] Y[row][c],
// part_b[z][o] = sum dPre[row][o], dPre = dH (.) (H > 0) when H is given.  Both operands are read row by row with the
// contracted index (the row) outermost, so the shared tiles are filled without a transpose.  BM = 32 / 64 (d_out).
template <int BM>
__global__ void __launch_bounds__(F_THREADS) linear_dw_kernel(const float *__restrict__ dH, int64_t ld_dh, const float *__restrict__ H, int64_t ld_h,
                                                               const float *__restrict__ Y, int64_t ldy, float *__restrict__ part_w,
                                                               float *__restrict__ part_b, int64_t n, int d_in, int d_out, int64_t chunk) {
    constexpr int BN = 64, TM = BM / 8;                 // thread tile TM x 4
    __shared__ __align__(16) float As[FBK][BM];         // [row][o]
    __shared__ __align__(16) float Bs[FBK][BN];         // [row][c]
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int c0 = blockIdx.x * BN;
    const int64_t r_begin = (int64_t)blockIdx.y * chunk, r_end = min(n, r_begin + chunk);
    float acc[TM][4];
    float bsum[TM];
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        bsum[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    }
    constexpr int NA = FBK * BM / 4 / F_THREADS;        // float4 per thread per tile: 2 (BM 32) / 4 (BM 64)
    constexpr int NBQ = FBK * BN / 4 / F_THREADS;       // 4
    float4 ra[NA], rb[NBQ];
    auto fetch = [&](int64_t r0) {
#pragma unroll
        for (int j = 0; j < NA; ++j) {
            const int f = tid + F_THREADS * j, rr = f / (BM / 4), o = 4 * (f % (BM / 4));
            const int64_t row = r0 + rr;
            float4 v = f4_zero();
            if (row < r_end && o < d_out) {
                v = __ldg(reinterpret_cast<const float4 *>(dH + row * ld_dh + o));       // ld_dh >= d_out rounded up to 4
                if (H) {
                    const float4 h = __ldg(reinterpret_cast<const float4 *>(H + row * ld_h + o));
                    if (!(h.x > 0.f)) v.x = 0.f;
                    if (!(h.y > 0.f)) v.y = 0.f;
                    if (!(h.z > 0.f)) v.z = 0.f;
                    if (!(h.w > 0.f)) v.w = 0.f;
                }
                if (o + 1 >= d_out) v.y = 0.f;
                if (o + 2 >= d_out) v.z = 0.f;
                if (o + 3 >= d_out) v.w = 0.f;
            }
            ra[j] = v;
        }
#pragma unroll
        for (int j = 0; j < NBQ; ++j) {
            const int f = tid + F_THREADS * j, rr = f / (BN / 4), c = c0 + 4 * (f % (BN / 4));
            const int64_t row = r0 + rr;
            float4 v = f4_zero();
            if (row < r_end && c < d_in) {
                v = __ldg(reinterpret_cast<const float4 *>(Y + row * ldy + c));          // ldy >= d_in rounded up to 4
                if (c + 1 >= d_in) v.y = 0.f;
                if (c + 2 >= d_in) v.z = 0.f;
                if (c + 3 >= d_in) v.w = 0.f;
            }
            rb[j] = v;
        }
    };
    auto stash = [&]() {
#pragma unroll
        for (int j = 0; j < NA; ++j) {
            const int f = tid + F_THREADS * j;
            *reinterpret_cast<float4 *>(&As[f / (BM / 4)][4 * (f % (BM / 4))]) = ra[j];
        }
#pragma unroll
        for (int j = 0; j < NBQ; ++j) {
            const int f = tid + F_THREADS * j;
            *reinterpret_cast<float4 *>(&Bs[f / (BN / 4)][4 * (f % (BN / 4))]) = rb[j];
        }
    };
    if (r_begin < r_end) fetch(r_begin);
    for (int64_t r0 = r_begin; r0 < r_end; r0 += FBK) {
        stash();
        __syncthreads();
        if (r0 + FBK < r_end) fetch(r0 + FBK);
#pragma unroll
        for (int kk = 0; kk < FBK; ++kk) {
            float a[TM];
#pragma unroll
            for (int i4 = 0; i4 < TM / 4; ++i4) {
                const float4 t = *reinterpret_cast<const float4 *>(&As[kk][ty * TM + 4 * i4]);
                a[4 * i4] = t.x; a[4 * i4 + 1] = t.y; a[4 * i4 + 2] = t.z; a[4 * i4 + 3] = t.w;
            }
            const float4 b = *reinterpret_cast<const float4 *>(&Bs[kk][4 * tx]);
#pragma unroll
            for (int i = 0; i < TM; ++i) {
                acc[i][0] = fmaf(a[i], b.x, acc[i][0]);
                acc[i][1] = fmaf(a[i], b.y, acc[i][1]);
                acc[i][2] = fmaf(a[i], b.z, acc[i][2]);
                acc[i][3] = fmaf(a[i], b.w, acc[i][3]);
                bsum[i] += a[i];
            }
        }
        __syncthreads();
    }
    float *pw = part_w + (int64_t)blockIdx.y * d_out * d_in;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int o = ty * TM + i;
        if (o >= d_out) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + 4 * tx + j;
            if (c < d_in) pw[(int64_t)o * d_in + c] = acc[i][j];
        }
        if (blockIdx.x == 0 && tx == 0) part_b[(int64_t)blockIdx.y * d_out + o] = bsum[i];
    }
}

template <bool A_KC, bool B_KC>
static cudaError_t launch_gemm(const GemmArgs &g, int splits, cudaStream_t st) {
    if (g.M == 0 || g.N == 0) return cudaSuccess;
    // 32-row tiles when M is small (dW: M = d_out) or when 64-row tiles would not fill the chip
    const int64_t ctas64 = cdiv(g.M, 64) * cdiv(g.N, g.N <= 16 ? 16 : g.N <= 32 ? 32 : 64) * splits;
    if (g.M <= 32 || ctas64 < 148 * 4) return launch_gemm_bm<32, A_KC, B_KC>(g, splits, st);
    return launch_gemm_bm<64, A_KC, B_KC>(g, splits, st);
}

// out[i] = sum_z part[z*stride + i], deterministic split-K reduce.  A block covers 32 outputs; its 8 warps take every
// 8th partial each (lane = output: coalesced), then meet in shared memory in warp order.  (One thread per output
// walking all ~100 partials took 13 us per call at the Pubmed shape -- four calls per step, all of it load latency.)
__global__ void __launch_bounds__(256) split_reduce_kernel(const float *__restrict__ part, int64_t stride, int splits,
                                                          float *__restrict__ out, int64_t count) {
    __shared__ float red[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * 32 + lane;
    float s = 0.f;
    if (i < count) {
        int z = warp;
        for (; z + 24 < splits; z += 32) {              // four loads in flight
            const float a = part[(int64_t)z * stride + i], b = part[(int64_t)(z + 8) * stride + i];
            const float c = part[(int64_t)(z + 16) * stride + i], d = part[(int64_t)(z + 24) * stride + i];
            s += a; s += b; s += c; s += d;
        }
        for (; z < splits; z += 8) s += part[(int64_t)z * stride + i];
    }
    red[warp][lane] = s;
    __syncthreads();
    if (warp == 0 && i < count) {
        float t = red[0][lane];
#pragma unroll
        for (int w = 1; w < 8; ++w) t += red[w][lane];
        out[i] = t;
    }
}

// partial column sums of dPre = dH * (H > 0): part[z, j] over row chunk z
__global__ void colsum_masked_kernel(const float *__restrict__ dH, int64_t ld_dh,
                                     const float *__restrict__ H, int64_t ld_h, int relu, int64_t n,
                                     int d_out, int64_t rows_per_block, float *__restrict__ part) {
    // blockDim = (32, 8): x over columns, y over rows
    __shared__ float red[8][33];
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r1 = min(n, r0 + rows_per_block);
    for (int j0 = 0; j0 < d_out; j0 += 32) {
        const int j = j0 + threadIdx.x;
        float s = 0.f;
        if (j < d_out)
            for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) {
                float v = dH[r * ld_dh + j];
                if (relu && !(H[r * ld_h + j] > 0.f)) v = 0.f;
                s += v;
            }
        red[threadIdx.y][threadIdx.x] = s;
        __syncthreads();
        if (threadIdx.y == 0 && j < d_out) {
            float t = 0.f;
            for (int y = 0; y < 8; ++y) t += red[y][threadIdx.x];
            part[(int64_t)blockIdx.x * d_out + j] = t;
        }
        __syncthreads();
    }
}

static inline int64_t db_rows_per_block(int64_t n) { return n / 2048 > 256 ? n / 2048 : 256; }

static void bwd_split_config(int64_t n, int32_t d_in, int32_t d_out, int *splits, int64_t *k_chunk) {
    // enough CTAs to fill 148 SMs a few times, chunks a multiple of BK, at most 1024 splits
    const int64_t tiles = cdiv(d_out, d_out <= 32 ? 32 : 64) * cdiv(d_in, d_in <= 16 ? 16 : d_in <= 32 ? 32 : 64);
    int64_t want = cdiv(148 * 6, tiles);          // (148 * 3: the dW kernel of the Pubmed input layer went from ~40 to 57 us -- it needs the CTAs)
    int64_t chunk = cdiv(cdiv(n, want), BK) * BK;
    if (chunk < 64) chunk = 64;
    int64_t s = cdiv(n, chunk);
    if (s > 1024) { chunk = cdiv(cdiv(n, 1024), BK) * BK; s = cdiv(n, chunk); }
    if (s < 1) s = 1;
    *splits = (int)s;
    *k_chunk = chunk;
}

}  // namespace gae

using namespace gae;

extern "C" int gae_linear_fwd_f32(const float *Yin, int64_t ld_in, const float *W, const float *b,
                                  float *H, int64_t ld_out, int64_t n, int32_t d_in, int32_t d_out,
                                  int32_t act, void *stream) {
    GAE_CHECK_ARG(n >= 0 && d_in > 0 && d_out > 0, "bad sizes");
    if (n == 0) return GAE_OK;
    GAE_CHECK_ARG(Yin && W && H, "null pointer");
    GAE_CHECK_ARG(ld_in >= d_in && ld_out >= d_out, "leading dimension too small");
    GAE_CHECK_ARG(act == GAE_ACT_IDENTITY || act == GAE_ACT_RELU, "unknown activation");
    cudaStream_t st = (cudaStream_t)stream;
    if (aligned16(Yin) && ld_in % 4 == 0 && ld_in >= (d_in + 3) / 4 * 4 && d_out <= 64) {      // 128-bit path
        dim3 grid((unsigned)cdiv(n, FBM), 1);
        if (d_out <= 16) linear_fwd_tn_kernel<16><<<grid, FW_THREADS, 0, st>>>(Yin, ld_in, W, b, H, ld_out, n, d_out, d_in, act);
        else if (d_out <= 32) linear_fwd_tn_kernel<32><<<grid, FW_THREADS, 0, st>>>(Yin, ld_in, W, b, H, ld_out, n, d_out, d_in, act);
        else linear_fwd_tn_kernel<64><<<grid, FW_THREADS, 0, st>>>(Yin, ld_in, W, b, H, ld_out, n, d_out, d_in, act);
        GAE_LAUNCH_CHECK();
        return GAE_OK;
    }
    GemmArgs g{};
    g.A = Yin; g.a_rs = ld_in; g.a_cs = 1;
    g.B = W; g.b_rs = 1; g.b_cs = d_in;   // B(k, j) = W[j, k]
    g.C = H; g.ldc = ld_out; g.bias = b;
    g.M = n; g.N = d_out; g.K = d_in; g.k_chunk = d_in; g.act = act;
    GAE_CUDA((launch_gemm<true, true>(g, 1, st)));
    return GAE_OK;
}

extern "C" int64_t gae_linear_bwd_ws_bytes(int64_t n, int32_t d_in, int32_t d_out) {
    if (n <= 0 || d_in <= 0 || d_out <= 0) return 0;
    int splits; int64_t chunk;
    bwd_split_config(n, d_in, d_out, &splits, &chunk);
    int64_t db_blocks = cdiv(n, db_rows_per_block(n));
    if (db_blocks < splits) db_blocks = splits;          // the 128-bit dW kernel writes one db partial per row chunk
    return (int64_t)sizeof(float) * ((int64_t)splits * d_out * d_in + db_blocks * d_out);
}

extern "C" int gae_linear_bwd_f32(const float *Yin, int64_t ld_in, const float *W, const float *H,
                                  int64_t ld_out, const float *dH, int64_t ld_dh, float *dYin,
                                  int64_t ld_dyin, float *dW, float *db, void *ws, int64_t ws_bytes,
                                  int64_t n, int32_t d_in, int32_t d_out, int32_t act, void *stream) {
    GAE_CHECK_ARG(n >= 0 && d_in > 0 && d_out > 0, "bad sizes");
    GAE_CHECK_ARG(Yin && W && dH && dW && db, "null pointer");
    GAE_CHECK_ARG(act == GAE_ACT_IDENTITY || (act == GAE_ACT_RELU && H), "ReLU adjoint needs H");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        GAE_CUDA(cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)d_out * d_in, st));
        GAE_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * (size_t)d_out, st));
        return GAE_OK;
    }
    if (ws_bytes < gae_linear_bwd_ws_bytes(n, d_in, d_out) || !ws) {
        set_error("linear_bwd workspace too small: have %lld need %lld", (long long)ws_bytes,
                  (long long)gae_linear_bwd_ws_bytes(n, d_in, d_out));
        return GAE_ERR_WORKSPACE;
    }
    const float *mask = (act == GAE_ACT_RELU) ? H : nullptr;
    int splits; int64_t chunk;
    bwd_split_config(n, d_in, d_out, &splits, &chunk);
    float *part_w = (float *)ws;
    float *part_b = part_w + (int64_t)splits * d_out * d_in;

    const int64_t do4 = (d_out + 3) / 4 * 4, di4 = (d_in + 3) / 4 * 4;
    const bool fast = d_out <= 64 && aligned16(dH) && ld_dh % 4 == 0 && ld_dh >= do4 && aligned16(Yin) && ld_in % 4 == 0 && ld_in >= di4 &&
                      (!mask || (aligned16(mask) && ld_out % 4 == 0 && ld_out >= do4));
    if (fast) {
        // dW and db partials per row chunk in one launch (128-bit loads, no transpose), then the two ordered reductions
        dim3 grid((unsigned)cdiv(d_in, 64), (unsigned)splits);
        if (d_out <= 32) linear_dw_kernel<32><<<grid, F_THREADS, 0, st>>>(dH, ld_dh, mask, ld_out, Yin, ld_in, part_w, part_b, n, d_in, d_out, chunk);
        else linear_dw_kernel<64><<<grid, F_THREADS, 0, st>>>(dH, ld_dh, mask, ld_out, Yin, ld_in, part_w, part_b, n, d_in, d_out, chunk);
        GAE_LAUNCH_CHECK();
        const int64_t cnt = (int64_t)d_out * d_in;
        split_reduce_kernel<<<(unsigned)cdiv(cnt, 32), 256, 0, st>>>(part_w, cnt, splits, dW, cnt);
        GAE_LAUNCH_CHECK();
        split_reduce_kernel<<<(unsigned)cdiv(d_out, 32), 256, 0, st>>>(part_b, d_out, splits, db, d_out);
        GAE_LAUNCH_CHECK();
    } else {
    // dW[j,k] = sum_n dPre[n,j] Yin[n,k]:  A(m=j,k=n) = dH[n*ld + j], B(k=n, n=k) = Yin[n*ld + k]
    {
        GemmArgs g{};
        g.A = dH; g.a_rs = 1; g.a_cs = ld_dh;
        g.Mask = mask; g.m_rs = 1; g.m_cs = ld_out;
        g.B = Yin; g.b_rs = ld_in; g.b_cs = 1;
        g.C = part_w; g.ldc = d_in; g.c_split_stride = (int64_t)d_out * d_in;
        g.M = d_out; g.N = d_in; g.K = n; g.k_chunk = chunk; g.act = GAE_ACT_IDENTITY;
        GAE_CUDA((launch_gemm<false, false>(g, splits, st)));
        const int64_t cnt = (int64_t)d_out * d_in;
        split_reduce_kernel<<<(unsigned)cdiv(cnt, 32), 256, 0, st>>>(part_w, cnt, splits, dW, cnt);
        GAE_LAUNCH_CHECK();
    }
    // db[j] = sum_n dPre[n,j]
    {
        const int64_t rows_per_block = db_rows_per_block(n);
        const int64_t blocks = cdiv(n, rows_per_block);
        colsum_masked_kernel<<<(unsigned)blocks, dim3(32, 8), 0, st>>>(dH, ld_dh, mask, ld_out, mask != nullptr, n,
                                                                       d_out, rows_per_block, part_b);
        GAE_LAUNCH_CHECK();
        split_reduce_kernel<<<(unsigned)cdiv(d_out, 32), 256, 0, st>>>(part_b, d_out, (int)blocks, db, d_out);
        GAE_LAUNCH_CHECK();
    }
    }
    // dYin[n,k] = sum_j dPre[n,j] W[j,k]
    if (dYin) {
        GAE_CHECK_ARG(ld_dyin >= d_in, "ld_dyin too small");
        GemmArgs g{};
        g.A = dH; g.a_rs = ld_dh; g.a_cs = 1;
        g.Mask = mask; g.m_rs = ld_out; g.m_cs = 1;
        g.B = W; g.b_rs = d_in; g.b_cs = 1;
        g.C = dYin; g.ldc = ld_dyin;
        g.M = n; g.N = d_in; g.K = d_out; g.k_chunk = d_out; g.act = GAE_ACT_IDENTITY;
        GAE_CUDA((launch_gemm<true, false>(g, 1, st)));
    }
    return GAE_OK;
}
