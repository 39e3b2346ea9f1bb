// K5 / K6 on the 5th-generation tensor cores, pipelined form: the dense pass of the fused decoder for d <= 16 over
// the UPPER TRIANGLE of 128 x 128 tiles (see decoder_tc.cu for the algebra: S = Z_I Z_J^T once per tile pair,
// G_I += sigma Z_J, G_J = sigma^T Z_I).  Reference: gae.py:71 + train_inductive.py:44-51.
//
// What differs from decoder_tc.cu (the TF32 form it supersedes; that kernel measured 359 us at the Pubmed shape with
// the MUFU chain busy 23 % of the time -- everything else was un-overlapped latency; this one: 235 us, MUFU pipe 42 % busy):
//
//  * Operands are split into TWO FP16 TERMS (hi = fp16(x), lo = fp16(x - hi): 11 + 11 significand bits, products
//    exact in the fp32 accumulators, hi*hi + lo*hi + hi*lo as before -> the same ~2^-21 relative accuracy and the same
//    1e-5 parity tests).  kind::f16 contracts K = 16 per instruction where kind::tf32 contracts 8, so a tile needs
//    35 MMAs instead of 70 (an M = 128 MMA costs 74 - 110 cycles whatever N is, tools/mma_bench.cu); sigma_hi and
//    sigma_lo of a row's 32 keys pack into the 32 TMEM columns its logits came from (in place: 128 columns per tile),
//    and sigma^T in shared memory shrinks from 144 KB to 72 KB -- room to double-buffer it.  FP16's range is handled by
//    a power-of-two scale taken from max|Zd| (a tiny pre-pass) that puts the largest operand just below 2^14; it folds
//    into constants the chain multiplies by anyway.
//  * WARP ROLES.  16 compute warps do nothing but S -> sigma (tcgen05.ld, the ex2 / rcp chain, sigma back to TMEM in place
//    and transposed to shared memory).  4 service warps (one per TMEM lane quarter; thread r = row r of a block) bring
//    the Z_J operand tiles in (fp16 hi / lo [row][dim] tiles for S two tiles ahead, their transposes for sigma Z_J) and
//    take the gradient accumulators out (G_I into registers, G_J to its global slot).  One issuing warp owns every
//    tcgen05.mma, each batch under elect.sync (under `if (lane == 0)` ptxas wraps every MMA in an elect / vote loop).
//    No CTA barrier inside the loop: mbarriers only ("S landed", "gradients landed", "sigma stored"), every lane
//    polling (one polling lane + __syncwarp sees a flip ~370 cycles later, tools/bar_bench.cu).
//  * TWO BUFFER SETS by tile parity (S / sigma in TMEM 2 x 128 columns, gradient accumulators 2 x 64, sigma^T, Z_J^T and
//    the [row][dim] tiles in shared memory): at "sigma(k) stored" the issuing warp queues G(k) and then S(k + 2) -- into
//    the buffer sigma(k) sits in; the pipe runs it after G_I(k) -- so S is a whole tile ahead of its use and the gradient
//    MMAs of tile k run under the chain of tile k + 1; the service warps read them out one tile behind.
//
// Shared-memory operands: K-major, no swizzle, canonical 8-row x 16-byte core matrices (8 fp16 along K):
//     off(m, k) = (m / 8) * SBO + (k / 8) * LBO + (m % 8) * 16 + (k % 8) * 2
// TMEM A operand of sigma Z_J: lane = query row, one 32-bit column = two consecutive keys (low half first).
// Deterministic: fixed slots per (I, split) and per tile, summed in fixed order by dec_slot_reduce_kernel.  Waits are
// bounded (%globaltimer): on expiry an error word is set and the loss becomes NaN -- never a hung GPU.
// GAE_TC_PROF=1 prints per-warp cycles per phase after every launch (debug; profiles/r02_tc16_phase_profile.log).
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace gae {

constexpr int H_COMPUTE = 512;              // 16 compute warps: S -> sigma, nothing else
constexpr int H_SERVICE = 128;              // 4 service warps (one per TMEM lane quarter): operand tiles in, gradient tiles out
constexpr int H_THREADS = H_COMPUTE + H_SERVICE + 32;       // + the issuing warp
constexpr int H_TILE = 128;
constexpr int H_D = 16;
constexpr uint32_t H_TMEM_COLS = 512;
// TMEM: S buffers at 0 / 128 (tile k in buffer k & 1); two sets of gradient accumulators (tile k in set k & 1) at 256 /
// 384, each: G_I = sigma_hi [Z_hi | Z_lo] + sigma_lo Z_hi on its first half (32 columns), G_J = sigma^T Z_I (32).
// (Measured on this part, tools/mma_bench.cu: an M = 128 kind::f16 MMA occupies the pipe for ~74 cycles with A in TMEM
// and ~98 with A in shared memory whatever N <= 64 is, 110 at N = 128, and separate accumulators change nothing -- the 35
// MMAs of a tile cost ~3 100 cycles however they are arranged, so the schedule must keep them off the critical path.)
constexpr uint32_t H_COL_ACC = 256, H_ACC_STRIDE = 128;
constexpr uint32_t H_GI = 0, H_GJ = 32;
// shared memory map (bytes)
constexpr int H_Z_BYTES = H_TILE * H_D * 2;                   // [row][dim] fp16 tile, K = dim: LBO 128, SBO 256
constexpr int H_OFF_ZI_HI = 0, H_OFF_ZI_LO = H_Z_BYTES;
constexpr int H_OFF_ZJ = 2 * H_Z_BYTES;                       // [buffer 0 / 1][hi | lo]
constexpr int H_OFF_ZIT = H_OFF_ZJ + 4 * H_Z_BYTES;           // Z_I^T: [n = 32][k' = 2 row + h], LBO 128, SBO 4096
constexpr int H_ZIT_SBO = 4096, H_ZIT_BYTES = 4 * H_ZIT_SBO;
constexpr int H_OFF_ZJT = H_OFF_ZIT + H_ZIT_BYTES;            // Z_J^T x 2: [n = hi dims | lo dims][key]; LBO 144: the 4 key groups of a warp on different banks
constexpr int H_ZJT_LBO = 144, H_ZJT_SBO = 16 * H_ZJT_LBO, H_ZJT_BYTES = 4 * H_ZJT_SBO;
constexpr int H_OFF_SGT = H_OFF_ZJT + 2 * H_ZJT_BYTES;        // sigma^T x 2: [key][k' = 2 row + h]; LBO 144 keeps the 32 rows of a warp on 32 banks
constexpr int H_SGT_LBO = 144, H_SGT_SBO = 32 * H_SGT_LBO, H_SGT_BYTES = 16 * H_SGT_SBO;
constexpr int H_OFF_BAR = H_OFF_SGT + 2 * H_SGT_BYTES;
constexpr int H_SMEM_BYTES = H_OFF_BAR + 64;
static_assert(H_SMEM_BYTES + 1024 <= 227 * 1024, "shared memory budget");

struct HArgs {
    const float *Zd;
    int64_t ldz, n;
    int32_t d, T, splits;
    float *dz_part;            // [splits][n][16]: G_I partial of CTA (I, s)
    float *dzT_part;           // [T][n][16]: G_J partial of tile (I, J) at slot I, rows of block J
    double *loss_part;         // [T * splits]
    uint32_t *err;             // bounded-wait expiry counter
    const uint32_t *absmax;    // bits of max |Zd| (dec_absmax_kernel)
    float *probe_S, *probe_GI, *probe_GJ;     // probe (one tile, one CTA): raw S, G_I, G_J
    int32_t probe_I, probe_J;
    unsigned long long *prof;  // GAE_TC_PROF=1: per-phase cycle totals of warp 0 (compute) and of the issuing warp, all CTAs
};

// phase stamps (debug): P_* index prof[]
enum { P_SETUP = 0, P_WAIT_S, P_LD, P_CHAIN, P_WAIT_G, P_READOUT, P_STORE, P_ARRIVE, P_TAIL, P_TILES, P_ISS_WAIT, P_ISS_ISSUE, P_CTAS, P_COUNT };
#define H_STAMP(idx)                                                                  \
    if (prof_on) {                                                                    \
        const long long now_ = clock64();                                             \
        atomicAdd(a.prof + warp * P_COUNT + (idx), (unsigned long long)(now_ - t_prev)); \
        t_prev = now_;                                                                \
    }

// max |Zd| as float bits (non-negative floats order like their bit patterns; NaN sorts above inf and poisons the scale,
// which is what it should do)
__global__ void dec_absmax_kernel(const float *__restrict__ Zd, int64_t ldz, int64_t n, int d, uint32_t *out) {
    uint32_t m = 0;
    const int64_t total = n * d;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / d;
        const int c = (int)(i - r * d);
        m = max(m, __float_as_uint(__ldg(Zd + r * ldz + c)) & 0x7fffffffu);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

// two values at once through the packed converter (F2FP on the ALU pipe; the scalar cvt is an XU-pipe F2F)
__device__ __forceinline__ void h_split2(float x0, float x1, uint32_t &hi, uint32_t &lo) {
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
    const float2 hf = __half22float2(*reinterpret_cast<const __half2 *>(&hi));
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(x1 - hf.y), "f"(x0 - hf.x));
}

// Service thread r (one per row of a 128-row block) loads its whole row of Zd; the power-of-two scale zs is applied when
// a store routine consumes the values, not at the load (the loads stay in flight meanwhile).
__device__ __forceinline__ void h_load_row(const HArgs &a, int64_t row, float (&x)[H_D]) {
    const bool rv = row < a.n;
    const float *src = a.Zd + row * a.ldz;
#pragma unroll
    for (int k = 0; k < H_D; ++k) x[k] = (rv && k < a.d) ? __ldg(src + k) : 0.f;
}
// hi / lo fp16 pairs of a row: h[p] = (x[2p], x[2p+1]) hi halves, l[p] the remainders
__device__ __forceinline__ void h_split_row(const float (&x)[H_D], float zs, uint32_t (&h)[H_D / 2], uint32_t (&l)[H_D / 2]) {
#pragma unroll
    for (int p2 = 0; p2 < H_D / 2; ++p2) h_split2(x[2 * p2] * zs, x[2 * p2 + 1] * zs, h[p2], l[p2]);
}
// [row][dim] tiles (K = dim): row r owns two 16-byte core-matrix rows in each of the hi and lo tiles
__device__ __forceinline__ void h_store_z_row(int r, const uint32_t (&h)[8], const uint32_t (&l)[8], unsigned char *hi_tile, unsigned char *lo_tile) {
    const int off = (r >> 3) * 256 + (r & 7) * 16;
    *reinterpret_cast<uint4 *>(hi_tile + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4 *>(hi_tile + off + 128) = make_uint4(h[4], h[5], h[6], h[7]);
    *reinterpret_cast<uint4 *>(lo_tile + off) = make_uint4(l[0], l[1], l[2], l[3]);
    *reinterpret_cast<uint4 *>(lo_tile + off + 128) = make_uint4(l[4], l[5], l[6], l[7]);
}
// the row's packed halves back from the [row][dim] tiles (the service thread wrote them itself two tiles ago)
__device__ __forceinline__ void h_load_z_row(int r, const unsigned char *hi_tile, const unsigned char *lo_tile, uint32_t (&h)[8], uint32_t (&l)[8]) {
    const int off = (r >> 3) * 256 + (r & 7) * 16;
    const uint4 h0 = *reinterpret_cast<const uint4 *>(hi_tile + off), h1 = *reinterpret_cast<const uint4 *>(hi_tile + off + 128);
    const uint4 l0 = *reinterpret_cast<const uint4 *>(lo_tile + off), l1 = *reinterpret_cast<const uint4 *>(lo_tile + off + 128);
    h[0] = h0.x; h[1] = h0.y; h[2] = h0.z; h[3] = h0.w; h[4] = h1.x; h[5] = h1.y; h[6] = h1.z; h[7] = h1.w;
    l[0] = l0.x; l[1] = l0.y; l[2] = l0.z; l[3] = l0.w; l[4] = l1.x; l[5] = l1.y; l[6] = l1.z; l[7] = l1.w;
}
// Z_J^T: [n][key] with n = dim (hi) or 16 + dim (lo), K = key = r
__device__ __forceinline__ void h_store_zt_row(int r, const uint32_t (&h)[8], const uint32_t (&l)[8], unsigned char *t_tile) {
    const int koff = (r >> 3) * H_ZJT_LBO + (r & 7) * 2;
#pragma unroll
    for (int dim = 0; dim < H_D; ++dim) {
        const uint32_t hw = h[dim >> 1], lw = l[dim >> 1];
        *reinterpret_cast<unsigned short *>(t_tile + (dim >> 3) * H_ZJT_SBO + (dim & 7) * 16 + koff) = (unsigned short)((dim & 1) ? hw >> 16 : hw & 0xffffu);
        *reinterpret_cast<unsigned short *>(t_tile + (2 + (dim >> 3)) * H_ZJT_SBO + (dim & 7) * 16 + koff) = (unsigned short)((dim & 1) ? lw >> 16 : lw & 0xffffu);
    }
}
// Z_I^T for sigma^T Z_I with hi / lo of sigma interleaved along K (k' = 2 row + h):
//   n < 16 : Z_hi[row][n] at h = 0 and h = 1        (sigma_hi Z_hi + sigma_lo Z_hi)
//   n >= 16: Z_lo[row][n - 16] at h = 0, 0 at h = 1 (sigma_hi Z_lo)
__device__ __forceinline__ void h_store_zit_row(int r, const uint32_t (&h)[8], const uint32_t (&l)[8], unsigned char *t_tile) {
    const int koff = (r >> 2) * 128 + (r & 3) * 4;          // k' = 2 r: group (2 r) / 8, position 2 ((2 r) % 8)
#pragma unroll
    for (int dim = 0; dim < H_D; ++dim) {
        const uint32_t hv = (dim & 1) ? h[dim >> 1] >> 16 : h[dim >> 1] & 0xffffu;
        const uint32_t lv = (dim & 1) ? l[dim >> 1] >> 16 : l[dim >> 1] & 0xffffu;
        *reinterpret_cast<uint32_t *>(t_tile + (dim >> 3) * H_ZIT_SBO + (dim & 7) * 16 + koff) = hv | (hv << 16);
        *reinterpret_cast<uint32_t *>(t_tile + (2 + (dim >> 3)) * H_ZIT_SBO + (dim & 7) * 16 + koff) = lv;
    }
}

// The element-wise chain on 32 logits of one row.  v[e] (the raw accumulator, logit / s2) becomes the shared-memory
// word of sigma: fp16 sigma_hi in the low half, fp16 sigma_lo in the high half.  msum / prod collect the loss.
// c1 = -log2(e) * s2.
template <bool RAGGED>
__device__ __forceinline__ void h_chain(uint32_t (&v)[32], float c1, float &msum, float &prod, bool row_ok, int64_t keys_left) {
#pragma unroll
    for (int e = 0; e < 32; ++e) {
        const float x = __uint_as_float(v[e]);
        float ex, rc;
        const float t = fabsf(x) * c1;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(t));
        float one_e = 1.0f + ex;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(one_e));
        float sg = (x >= 0.f) ? rc : ex * rc;
        if (RAGGED) {
            const bool ok = row_ok && (e < keys_left);
            sg = ok ? sg : 0.f;
            one_e = ok ? one_e : 1.0f;
        }
        msum += fmaxf(x, 0.f);
        prod *= one_e;
        // hi = sigma truncated to 11 significand bits (exact in fp16 for sigma >= 2^-14; below that the conversion rounds
        // to a multiple of 2^-24, an absolute error of 3e-8 on a weight of that size), lo = the exact remainder
        const float hb = __uint_as_float(__float_as_uint(sg) & 0xffffe000u);
        const float lo = sg - hb;
        uint32_t w;
        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(lo), "f"(hb));      // upper half <- lo, lower half <- hb
        v[e] = w;
    }
}

template <bool PROBE>
__global__ void __launch_bounds__(H_THREADS, 1) dec_dense_tc16_kernel(const HArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ double red[H_COMPUTE / 32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool compute = warp < H_COMPUTE / 32, service = !compute && warp < (H_COMPUTE + H_SERVICE) / 32, issuer = !compute && !service;
    const int q = warp & 3, cq = (warp >> 2) & 3;      // TMEM lane quarter (16 = 0 mod 4: also right for the service warps), key quarter
    // grid = (splits, T): row block 0 -- the longest runs -- is scheduled first, the short tail rows last
    const int I = PROBE ? a.probe_I : (int)blockIdx.y;
    const int split = PROBE ? 0 : (int)blockIdx.x;
    int j_begin, j_end;
    if (PROBE) {
        j_begin = a.probe_J;
        j_end = a.probe_J + 1;
    } else {
        const int len = a.T - I, per = (len + a.splits - 1) / a.splits;
        j_begin = I + split * per;
        j_end = min(a.T, j_begin + per);
    }
    const int count = max(0, j_end - j_begin);
    const long long t_start = clock64();
    const uint32_t bar0 = tc_smem_u32(smem + H_OFF_BAR);
    // bar_s[b] = bar0 + 8 b, bar_g[b] = bar0 + 16 + 8 b, bar_sig[b] = bar0 + 32 + 8 b  (b = tile & 1).  "sigma stored" needs
    // two barriers: nothing in iteration k + 1 makes a fast warp wait for the issuing warp's round k (S(k + 1) and G(k - 1)
    // were issued a round earlier), so it can arrive for tile k + 1 while a slow warp has not arrived for tile k -- on a
    // single barrier that arrival would complete the wrong phase.  It cannot get two tiles ahead: S(k + 2) is issued only
    // after every thread has arrived for tile k.
    const uint32_t bar_sig = bar0 + 32;

    // power-of-two operand scale from max |Zd|: scaled magnitudes lie in [2^13, 2^14) at the top (fp16 overflows at
    // 65504, and its subnormals start 2^-14 -- 27 binades below the largest operand)
    float zs, gs, s2;
    {
        const uint32_t mb = __ldg(a.absmax);
        int sh = mb ? (int)(mb >> 23) - 127 - 13 : 0;               // both ways: small embeddings keep their low-order bits too
        sh = sh < -60 ? -60 : (sh > 60 ? 60 : sh);
        zs = __uint_as_float((uint32_t)(127 - sh) << 23);          // 2^-sh  : Zd -> operands
        gs = __uint_as_float((uint32_t)(127 + sh) << 23);          // 2^sh   : gradient products -> true scale
        s2 = __uint_as_float((uint32_t)(127 + 2 * sh) << 23);      // 4^sh   : accumulator -> logit
    }

    // ---- set-up: barriers, TMEM, the stationary row block, the first two key blocks ------------------------
    if (tid == 0) {
        for (int b = 0; b < 4; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8 * b) : "memory");   // S x 2, gradients x 2
        for (int b = 0; b < 2; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_sig + 8 * b), "r"(H_COMPUTE + H_SERVICE) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&tmem_slot)),
                     "r"(H_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    const int r = 32 * q + lane;                                  // my row inside a tile (compute: query row; service: query and key row)
    if (service) {
        float x[H_D], y0[H_D], y1[H_D];
        h_load_row(a, (int64_t)I * H_TILE + r, x);                 // all three loads in flight before the first use
        if (count > 0) h_load_row(a, (int64_t)j_begin * H_TILE + r, y0);
        if (count > 1) h_load_row(a, (int64_t)(j_begin + 1) * H_TILE + r, y1);
        uint32_t h[8], l[8];
        h_split_row(x, zs, h, l);
        h_store_z_row(r, h, l, smem + H_OFF_ZI_HI, smem + H_OFF_ZI_LO);
        h_store_zit_row(r, h, l, smem + H_OFF_ZIT);
        if (count > 0) {
            h_split_row(y0, zs, h, l);
            h_store_z_row(r, h, l, smem + H_OFF_ZJ, smem + H_OFF_ZJ + H_Z_BYTES);
        }
        if (count > 1) {
            h_split_row(y1, zs, h, l);
            h_store_z_row(r, h, l, smem + H_OFF_ZJ + 2 * H_Z_BYTES, smem + H_OFF_ZJ + 3 * H_Z_BYTES);
        }
    }
    tc_fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const bool prof_on = a.prof != nullptr && lane == 0;
    if (prof_on) {
        atomicAdd(a.prof + warp * P_COUNT + (issuer ? P_CTAS : P_SETUP), issuer ? 1ull : (unsigned long long)(clock64() - t_start));
        if (!issuer) atomicAdd(a.prof + warp * P_COUNT + P_TILES, (unsigned long long)count);
    }
    long long t_prev = clock64();
    const uint32_t lane_base = (uint32_t)(32 * q) << 16;

    if (issuer) {
        // ================= the issuing warp: owns the tensor-core queue ===========================================
        // The whole warp runs this loop converged; each batch of MMAs sits under elect.sync (see tc_elect_one).
        constexpr uint32_t IDESC_S = tc_idesc(128, 128, 0u);
        constexpr uint32_t IDESC_G32 = tc_idesc(128, 32, 0u);     // A x [Z_hi | Z_lo]
        constexpr uint32_t IDESC_G16 = tc_idesc(128, 16, 0u);     // A x Z_hi
        const uint64_t d_zi_hi = tc_desc(tc_smem_u32(smem + H_OFF_ZI_HI), 128, 256), d_zi_lo = tc_desc(tc_smem_u32(smem + H_OFF_ZI_LO), 128, 256);
        const uint64_t d_zj0 = tc_desc(tc_smem_u32(smem + H_OFF_ZJ), 128, 256);     // + 256 per tile of [buffer][hi | lo]
        const uint64_t d_zit = tc_desc(tc_smem_u32(smem + H_OFF_ZIT), 128, H_ZIT_SBO), d_zjt = tc_desc(tc_smem_u32(smem + H_OFF_ZJT), H_ZJT_LBO, H_ZJT_SBO);
        const uint64_t d_sgt = tc_desc(tc_smem_u32(smem + H_OFF_SGT), H_SGT_LBO, H_SGT_SBO);
        // S(k) = Z_I Z_J^T, split precision hi hi + lo hi + hi lo, one K = 16 step each
        auto issue_s = [&](int k) {
            const int b = k & 1;
            const uint32_t d = tmem + 128u * (uint32_t)b;
            const uint64_t zj_hi = d_zj0 + (uint64_t)(b * (2 * H_Z_BYTES / 16)), zj_lo = zj_hi + H_Z_BYTES / 16;
            if (tc_elect_one()) {
                tc_mma_ss_f16(d, d_zi_hi, zj_hi, IDESC_S, 0);
                tc_mma_ss_f16(d, d_zi_lo, zj_hi, IDESC_S, 1);
                tc_mma_ss_f16(d, d_zi_hi, zj_lo, IDESC_S, 1);
                tc_commit(bar0 + 8u * (uint32_t)b);
            }
            __syncwarp();
        };
        if (count > 0) issue_s(0);
        if (count > 1) issue_s(1);
        for (int k = 0; k < count; ++k) {
            const int b = k & 1;
            tc_wait(bar_sig + 8u * (uint32_t)b, (uint32_t)((k >> 1) & 1), a.err);   // sigma(k), Z_J^T(k), [row][dim] tiles of block k + 2 are in place
            tc_fence_after();
            H_STAMP(P_ISS_WAIT)
            const uint32_t sg = tmem + 128u * (uint32_t)b, acc = tmem + H_COL_ACC + H_ACC_STRIDE * (uint32_t)b;
            const uint64_t zjt = d_zjt + (uint64_t)(b * (H_ZJT_BYTES / 16)), sgt = d_sgt + (uint64_t)(b * (H_SGT_BYTES / 16));
            const bool diag = j_begin + k == I;
            if (tc_elect_one()) {
                // G_I = sigma Z_J: A = sigma in TMEM (8 columns = 16 keys per step; hi at +0 / +8, lo at +16 / +24 of each 32);
                // the sigma_lo Z_hi product lands on the first 16 columns of the same accumulator
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint32_t a_hi = sg + 32u * (uint32_t)(i >> 1) + 8u * (uint32_t)(i & 1);
                    tc_mma_ts_f16(acc + H_GI, a_hi, zjt + i * (2 * H_ZJT_LBO / 16), IDESC_G32, i != 0);
                    tc_mma_ts_f16(acc + H_GI, a_hi + 16u, zjt + i * (2 * H_ZJT_LBO / 16), IDESC_G16, 1);
                }
                // G_J = sigma^T Z_I over K' = 256 (hi / lo of sigma interleaved along K')
                if (!diag) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        tc_mma_ss_f16(acc + H_GJ, sgt + i * (2 * H_SGT_LBO / 16), d_zit + i * 16, IDESC_G32, i != 0);
                }
                tc_commit(bar0 + 16u + 8u * (uint32_t)b);
            }
            __syncwarp();
            if (k + 2 < count) issue_s(k + 2);                        // into the buffer sigma(k) sits in: the pipe runs it after G_I(k)
            H_STAMP(P_ISS_ISSUE)
        }
    } else if (service) {
        // ================= the 4 service warps: operand tiles in, gradient tiles out ===============================
        // Thread r: key row r of the Z_J blocks (its 16 values -> fp16 hi / lo tiles) and query row r / key row r of the
        // gradient accumulators (TMEM lane r).
        float gi[H_D];
#pragma unroll
        for (int k = 0; k < H_D; ++k) gi[k] = 0.f;
        const int64_t row = (int64_t)I * H_TILE + r;
        // gradient tiles of key block Jr leave TMEM: G_I into my registers, G_J into its slot
        auto read_out = [&](int Jr, int set) {
            const uint32_t acc = tmem + lane_base + H_COL_ACC + H_ACC_STRIDE * (uint32_t)set;
            {
                uint32_t g[32];
                tc_ld32(acc + H_GI, g);
                tc_wait_ld();
#pragma unroll
                for (int k = 0; k < H_D; ++k) {
                    const float t = (__uint_as_float(g[k]) + __uint_as_float(g[16 + k])) * gs;
                    gi[k] += t;
                    if (PROBE) a.probe_GI[r * H_D + k] = t;
                }
            }
            if (Jr != I) {
                uint32_t g[32];
                tc_ld32(acc + H_GJ, g);
                tc_wait_ld();
#pragma unroll
                for (int k = 0; k < H_D; ++k) g[k] = __float_as_uint((__uint_as_float(g[k]) + __uint_as_float(g[16 + k])) * gs);
                const int64_t key = (int64_t)Jr * H_TILE + r;
                if (PROBE) {
                    for (int k = 0; k < H_D; ++k) a.probe_GJ[r * H_D + k] = __uint_as_float(g[k]);
                } else if (key < a.n) {
                    uint4 *o = reinterpret_cast<uint4 *>(a.dzT_part + ((int64_t)I * a.n + key) * H_D);
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) o[k4] = make_uint4(g[4 * k4], g[4 * k4 + 1], g[4 * k4 + 2], g[4 * k4 + 3]);
                }
            }
        };
        float y[H_D];                                              // my key row of block k + 2, in flight from L2
        if (2 < count) h_load_row(a, (int64_t)(j_begin + 2) * H_TILE + r, y);
        for (int k = 0; k < count; ++k) {
            const int J = j_begin + k, b = k & 1;
            const uint32_t ph = (uint32_t)((k >> 1) & 1);
            unsigned char *zj_hi = smem + H_OFF_ZJ + b * 2 * H_Z_BYTES, *zj_lo = zj_hi + H_Z_BYTES;
            uint32_t h[8], l[8];
            // Z_J^T(k) from the [row][dim] tiles of block k (same halves, transposed).  Its buffer b was last read by the
            // gradient MMAs of tile k - 2: waited for in iteration k - 1 (below).
            h_load_z_row(r, zj_hi, zj_lo, h, l);
            h_store_zt_row(r, h, l, smem + H_OFF_ZJT + b * H_ZJT_BYTES);
            if (k + 2 < count) {
                tc_wait(bar0 + 8u * (uint32_t)b, ph, a.err);      // the [row][dim] buffer b was read by S(k): it has landed
                h_split_row(y, zs, h, l);
                h_store_z_row(r, h, l, zj_hi, zj_lo);
            }
            tc_fence_before();       // orders the read-out of iteration k - 1 (tcgen05.ld, waited) before the arrival
            tc_fence_async_smem();
            tc_arrive(bar_sig + 8u * (uint32_t)b);
            H_STAMP(P_STORE)
            if (k + 3 < count) h_load_row(a, (int64_t)(J + 3) * H_TILE + r, y);      // next iteration's block, under the wait below
            // the previous tile's gradients (the MMAs of this tile go to the other accumulator set)
            if (k > 0) {
                tc_wait(bar0 + 16u + 8u * (uint32_t)(b ^ 1), (uint32_t)(((k - 1) >> 1) & 1), a.err);
                tc_fence_after();
                H_STAMP(P_WAIT_G)
                read_out(J - 1, b ^ 1);
                H_STAMP(P_READOUT)
            }
        }
        if (count > 0) {
            tc_wait(bar0 + 16u + 8u * (uint32_t)((count - 1) & 1), (uint32_t)(((count - 1) >> 1) & 1), a.err);
            tc_fence_after();
            read_out(j_end - 1, (count - 1) & 1);
        }
        if (!PROBE && row < a.n) {
            float4 *o = reinterpret_cast<float4 *>(a.dz_part + ((int64_t)split * a.n + row) * H_D);
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) o[k4] = make_float4(gi[4 * k4], gi[4 * k4 + 1], gi[4 * k4 + 2], gi[4 * k4 + 3]);
        }
    } else {
        // ================= the 16 compute warps: S -> sigma =========================================================
        const float c1 = -1.4426950408889634f * s2;
        float lacc = 0.f;       // per thread: <= 40 tiles x 32 pairs of O(1) terms; fp64 only from the CTA sum on
        const bool row_ok = (int64_t)I * H_TILE + r < a.n;
        // where my row of sigma^T goes: k' = 2 r (hi), 2 r + 1 (lo): one 32-bit word; key c adds (c / 8) * SBO + (c % 8) * 16
        unsigned char *sgt_row0 = smem + H_OFF_SGT + (4 * cq) * H_SGT_SBO + (r >> 2) * H_SGT_LBO + (r & 3) * 4;
        for (int k = 0; k < count; ++k) {
            const int J = j_begin + k, b = k & 1;
            const uint32_t ph = (uint32_t)((k >> 1) & 1);
            const bool diag = J == I;
            tc_wait(bar0 + 8u * (uint32_t)b, ph, a.err);
            tc_fence_after();
            H_STAMP(P_WAIT_S)
            // ---- compute phase: my row, 32 keys
            const uint32_t s_addr = tmem + lane_base + 128u * (uint32_t)b + 32u * (uint32_t)cq;
            const int64_t key0 = (int64_t)J * H_TILE + 32 * cq;
            const bool ragged = ((int64_t)I * H_TILE + H_TILE > a.n) || ((int64_t)J * H_TILE + H_TILE > a.n);
            uint32_t v[32];
            tc_ld32(s_addr, v);
            tc_wait_ld();
            H_STAMP(P_LD)
            if (PROBE)
                for (int e = 0; e < 32; ++e) a.probe_S[(int64_t)r * H_TILE + 32 * cq + e] = __uint_as_float(v[e]) * s2;
            float msum = 0.f, prod = 1.f;
            if (ragged) h_chain<true>(v, c1, msum, prod, row_ok, a.n - key0);
            else h_chain<false>(v, c1, msum, prod, true, 32);
            {
                // sum softplus = sum max(x, 0) + ln prod (1 + e^-|x|): 32 factors in (1, 2] cannot overflow
                float l2;
                asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(prod));
                const float tile_sum = msum * s2 + l2 * 0.6931471805599453f;
                lacc += diag ? tile_sum : 2.f * tile_sum;
            }
            H_STAMP(P_CHAIN)
            // ---- store phase: sigma hi | lo over my 32 columns of S (TMEM, in place), sigma^T into buffer b -- last read by
            //      the gradient MMAs of tile k - 2, which landed long ago (their barrier is checked, not waited on in practice)
            {
                uint32_t tw[32];
#pragma unroll
                for (int p2 = 0; p2 < 16; ++p2) {
                    tw[p2] = __byte_perm(v[2 * p2], v[2 * p2 + 1], 0x5410);          // sigma_hi of keys 2p, 2p + 1
                    tw[16 + p2] = __byte_perm(v[2 * p2], v[2 * p2 + 1], 0x7632);     // sigma_lo
                }
                tc_st32(s_addr, tw);
            }
            if (k >= 2) tc_wait(bar0 + 16u + 8u * (uint32_t)b, (uint32_t)(((k - 2) >> 1) & 1), a.err);
            H_STAMP(P_WAIT_G)
            {
                unsigned char *sgt_row = sgt_row0 + b * H_SGT_BYTES;
#pragma unroll
                for (int e = 0; e < 32; ++e) *reinterpret_cast<uint32_t *>(sgt_row + (e >> 3) * H_SGT_SBO + (e & 7) * 16) = v[e];
            }
            tc_wait_st();
            H_STAMP(P_STORE)
            tc_fence_before();
            tc_fence_async_smem();
            tc_arrive(bar_sig + 8u * (uint32_t)b);
            H_STAMP(P_ARRIVE)
        }
        if (!PROBE) {
            double s = (double)lacc;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
            if (lane == 0) red[warp] = s;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (!issuer) { H_STAMP(P_TAIL) }
    if (!PROBE && tid == 0) {
        double t = 0.0;
        for (int w = 0; w < H_COMPUTE / 32; ++w) t += red[w];
        if (*reinterpret_cast<volatile uint32_t *>(a.err) != 0) t = __longlong_as_double(0x7ff8000000000000ll);   // a wait expired: NaN loss
        a.loss_part[(int64_t)split * a.T + I] = t;
    }
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(H_TMEM_COLS) : "memory");
}

// splits of the key range of a row block: runs of ~40 tiles at the longest row (a CTA's set-up and drain cost about
// two tiles; measured best at the Pubmed shape, profiles/r02_dec_time_*.log), but at least one CTA per SM
int dec_tc16_splits(int64_t n) {
    const int64_t T = cdiv(n, H_TILE);
    int64_t s = cdiv(T, 40);
    const int64_t fill = cdiv(148, T);
    if (fill > s) s = fill;
    if (s > T) s = T;
    if (s < 1) s = 1;
    return (int)s;
}

static cudaError_t h_attrs() {
    static bool attr_set = false;
    if (attr_set) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(dec_dense_tc16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, H_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(dec_dense_tc16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, H_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
    return cudaSuccess;
}

// err[0] = bounded-wait expiry counter, err[1] = bits of max |Zd|; both zeroed by the caller on this stream
cudaError_t dec_tc16_launch(const float *Zd, int64_t ldz, int64_t n, int d, int splits, float *dz_part, float *dzT_part,
                            double *loss_part, uint32_t *err, cudaStream_t st) {
    cudaError_t e = h_attrs();
    if (e != cudaSuccess) return e;
    const int64_t total = n * d;
    int blocks = (int)((total + 1023) / 1024);
    if (blocks > 148) blocks = 148;
    dec_absmax_kernel<<<blocks, 256, 0, st>>>(Zd, ldz, n, d, err + 1);
    count_launch();
    HArgs a{};
    a.Zd = Zd; a.ldz = ldz; a.n = n; a.d = d; a.T = (int)cdiv(n, H_TILE); a.splits = splits;
    a.dz_part = dz_part; a.dzT_part = dzT_part; a.loss_part = loss_part; a.err = err; a.absmax = err + 1;
    dim3 grid((unsigned)splits, (unsigned)a.T);
    static const bool prof = getenv("GAE_TC_PROF") != nullptr;      // debug: per-phase cycle totals, printed per launch (synchronises)
    constexpr int NW = H_THREADS / 32;
    if (prof) {
        e = cudaMalloc(&a.prof, NW * P_COUNT * sizeof(unsigned long long));
        if (e != cudaSuccess) return e;
        cudaMemsetAsync(a.prof, 0, NW * P_COUNT * sizeof(unsigned long long), st);
    }
    dec_dense_tc16_kernel<false><<<grid, H_THREADS, H_SMEM_BYTES, st>>>(a);
    count_launch();
    if (prof) {
        static unsigned long long h[NW * P_COUNT];
        cudaMemcpyAsync(h, a.prof, sizeof(h), cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        cudaFree(a.prof);
        const int iw = NW - 1;
        const double tiles = h[P_TILES] ? (double)h[P_TILES] : 1.0, ctas = h[iw * P_COUNT + P_CTAS] ? (double)h[iw * P_COUNT + P_CTAS] : 1.0;
        fprintf(stderr, "[gae tc16 prof] ctas %.0f tiles %.0f; cycles per tile (per CTA for setup / tail), lane 0 of each warp; warps 16-19 = service\n", ctas, tiles);
        fprintf(stderr, "[gae tc16 prof] warp  setup   tail | wait_S     ld  chain wait_G  store arrive readout |  sum\n");
        for (int w = 0; w < iw; ++w) {
            const unsigned long long *r = h + w * P_COUNT;
            const double sum = (double)(r[P_WAIT_S] + r[P_LD] + r[P_CHAIN] + r[P_STORE] + r[P_ARRIVE] + r[P_WAIT_G] + r[P_READOUT]) / tiles;
            fprintf(stderr, "[gae tc16 prof] %4d %6.0f %6.0f | %6.0f %6.0f %6.0f %6.0f %6.0f %6.0f %7.0f | %5.0f\n", w, r[P_SETUP] / ctas, r[P_TAIL] / ctas,
                    r[P_WAIT_S] / tiles, r[P_LD] / tiles, r[P_CHAIN] / tiles, r[P_WAIT_G] / tiles, r[P_STORE] / tiles, r[P_ARRIVE] / tiles,
                    r[P_READOUT] / tiles, sum);
        }
        fprintf(stderr, "[gae tc16 prof] issuer per tile: wait %.0f issue %.0f\n", h[iw * P_COUNT + P_ISS_WAIT] / tiles, h[iw * P_COUNT + P_ISS_ISSUE] / tiles);
    }
    return cudaGetLastError();
}

// one tile through the probe instantiation (gae_decoder_tile_probe_f32 with dec_tc = 2)
cudaError_t dec_tc16_probe(const float *Zd, int64_t ldz, int64_t n, int d, int tile_i, int tile_j, float *S, float *G_i,
                           float *G_j, uint32_t *err, cudaStream_t st) {
    cudaError_t e = h_attrs();
    if (e != cudaSuccess) return e;
    dec_absmax_kernel<<<32, 256, 0, st>>>(Zd, ldz, n, d, err + 1);
    HArgs a{};
    a.Zd = Zd; a.ldz = ldz; a.n = n; a.d = d; a.T = (int)cdiv(n, H_TILE); a.splits = 1; a.err = err; a.absmax = err + 1;
    a.probe_S = S; a.probe_GI = G_i; a.probe_GJ = G_j; a.probe_I = tile_i; a.probe_J = tile_j;
    dec_dense_tc16_kernel<true><<<1, H_THREADS, H_SMEM_BYTES, st>>>(a);
    return cudaGetLastError();
}

}  // namespace gae
