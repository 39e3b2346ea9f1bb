// K1 / K2, streaming variant: bulk-async (TMA engine) staged gather + sequential segmented sum.
//
// One warp owns 32 consecutive dst rows (or 32 hub-row segments).  Their edges form one flat
// stream; the warp walks it in batches of 32 edges.  For a batch, lane l resolves "its" edge
// (load-balanced search over the 32 row offsets held in registers), loads the source id and
// issues ONE 1-D bulk copy  cp.async.bulk.shared.global  of the whole source row (d*4 bytes,
// e.g. 256 B) into its slot of a shared-memory stage; completion is signalled on the stage's
// mbarrier (expect_tx / complete_tx).  STAGES batches are in flight per warp, so the gather
// needs no registers and the memory pipeline stays full across row boundaries -- short rows
// no longer cost a dependent latency chain each.  (SASS: UBLKCP + SYNCS.)
// Consumption is warp-uniform: all 32 lanes read slot j (d/32 floats per lane, conflict-free),
// add it to the running row sum and, when slot j is the last edge of its row, store the
// finished row with one coalesced 32-lane store.  The sum over a row's edges therefore runs
// sequentially in CSR order: bit-identical to the sequential fp32 loop of the CPU oracle,
// independent of any tuning knob.
//
// COPY_MODE 1 keeps the same pipeline but moves the rows with cooperative 16-byte
// cp.async (LDGSTS, half-warp per row) and commit/wait groups -- the A/B for the TMA path.
#include "common.cuh"

namespace gae {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}"
        ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

template <int VEC> struct VecT;
template <> struct VecT<1> { using T = float; };
template <> struct VecT<2> { using T = float2; };
template <> struct VecT<4> { using T = float4; };

template <int VEC> __device__ __forceinline__ void vadd(float (&a)[VEC], const typename VecT<VEC>::T &v);
template <> __device__ __forceinline__ void vadd<1>(float (&a)[1], const float &v) { a[0] += v; }
template <> __device__ __forceinline__ void vadd<2>(float (&a)[2], const float2 &v) { a[0] += v.x; a[1] += v.y; }
template <> __device__ __forceinline__ void vadd<4>(float (&a)[4], const float4 &v) {
    a[0] += v.x; a[1] += v.y; a[2] += v.z; a[3] += v.w;
}

// VEC = d/32 floats per lane; one warp per CTA; dynamic smem = STAGES * 32 * d*4 + metadata.
template <int VEC, int STAGES, int COPY_MODE, bool SEG>
__global__ void __launch_bounds__(32) spmm_stream_kernel(const StreamArgs a) {
    constexpr int RB = VEC * 32 * 4;            // bytes per feature row
    constexpr int STAGE_BYTES = 32 * RB;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *stage_base = smem;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES);
    uint32_t *meta_last = reinterpret_cast<uint32_t *>(bars + STAGES);          // [STAGES]
    uint8_t *meta_row = reinterpret_cast<uint8_t *>(meta_last + STAGES);        // [STAGES][32]
    using V = typename VecT<VEC>::T;

    const int lane = threadIdx.x;
    const int64_t item0 = (int64_t)blockIdx.x * 32;
    const int64_t item = item0 + lane;
    const bool valid = item < a.n_items;

    // ---- row descriptors in registers: lane i <-> row (segment) item0 + i --------------------
    int64_t start = 0, end = 0;
    bool zero_row = false;
    if (valid) {
        if (!SEG) {
            start = __ldg(a.rowptr + item);
            end = __ldg(a.rowptr + item + 1);
            if (a.seg_len > 0 && end - start > (int64_t)a.seg_len) end = start;   // hub row: segment pass owns it
            else zero_row = (end == start);
        } else {
            const int32_t k = __ldg(a.seg_row + item);
            const int64_t row = __ldg(a.long_row + k);
            const int64_t s = item - __ldg(a.long_seg_ptr + k);
            const int64_t r0 = __ldg(a.rowptr + row), r1 = __ldg(a.rowptr + row + 1);
            start = r0 + s * (int64_t)a.seg_len;
            end = min(start + (int64_t)a.seg_len, r1);
        }
    }
    const int deg = (int)(end - start);
    int pos = deg;                                // inclusive scan -> exclusive
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, pos, off);
        if (lane >= off) pos += t;
    }
    const int total = __shfl_sync(0xffffffffu, pos, 31);
    pos -= deg;

    // ---- empty rows are written as zeros (warp-uniform loop, coalesced 32-lane stores) -------
    if (!SEG && !a.accumulate) {
        unsigned m = __ballot_sync(0xffffffffu, zero_row);
        while (m) {
            const int i = __ffs(m) - 1;
            m &= m - 1;
            V z;
            memset(&z, 0, sizeof(V));
            __stcs(reinterpret_cast<V *>(a.Y + (item0 + i) * a.ldy) + lane, z);
        }
    }
    if (total == 0) return;
    const int nb = (total + 31) >> 5;

    if (COPY_MODE == 0) {
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < STAGES; ++s) mbar_init(smem_u32(bars + s), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncwarp();
    }

    const char *Xb = reinterpret_cast<const char *>(a.X);
    const int64_t row_bytes_ld = a.ldx * 4;

    auto issue = [&](int b) {
        const int s = b % STAGES;
        const int p = b * 32 + lane;
        const bool has = p < total;
        // load-balanced search: last row i with pos_i <= p (rows are non-decreasing in pos)
        int lo = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            const int cand = lo + step;
            const int pc = __shfl_sync(0xffffffffu, pos, cand & 31);
            if (cand < 32 && pc <= p) lo = cand;
        }
        const int64_t rs = __shfl_sync(0xffffffffu, start, lo);
        const int64_t re = __shfl_sync(0xffffffffu, end, lo);
        const int rp = __shfl_sync(0xffffffffu, pos, lo);
        const int64_t e = rs + (p - rp);
        const int idx = has ? __ldg(a.col + e) : 0;
        const unsigned lm = __ballot_sync(0xffffffffu, has && (e == re - 1));
        if (lane == 0) meta_last[s] = lm;
        meta_row[s * 32 + lane] = (uint8_t)lo;
        const int cnt = min(32, total - b * 32);
        if (COPY_MODE == 0) {
            const uint32_t bar = smem_u32(bars + s);
            // order the previous generic-proxy reads of this stage before the async-proxy writes
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (lane == 0) mbar_expect_tx(bar, (uint32_t)cnt * RB);
            __syncwarp();
            if (has) bulk_g2s(smem_u32(stage_base + s * STAGE_BYTES + lane * RB), Xb + (int64_t)idx * row_bytes_ld, RB, bar);
        } else {
            // cooperative 16-byte cp.async: (32*RB/16) chunks per batch, 32 per instruction
            constexpr int CPR = RB / 16;           // chunks per row
            constexpr int RPI = 32 / CPR;          // rows per instruction (d=64: 2)
            static_assert(CPR <= 32, "row too wide for the cooperative copy");
            const int sub = lane % CPR, which = lane / CPR;
#pragma unroll
            for (int k = 0; k < 32 / RPI; ++k) {
                const int slot = k * RPI + which;
                const int sidx = __shfl_sync(0xffffffffu, idx, slot);
                if (slot < cnt)
                    cp_async16(smem_u32(stage_base + s * STAGE_BYTES + slot * RB + sub * 16),
                               Xb + (int64_t)sidx * row_bytes_ld + sub * 16);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
    };

    const int pre = min(nb, STAGES - 1);
    for (int b = 0; b < pre; ++b) issue(b);
    if (COPY_MODE != 0)
        for (int b = pre; b < STAGES - 1; ++b) asm volatile("cp.async.commit_group;" ::: "memory");  // keep group count uniform

    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.f;

    for (int b = 0; b < nb; ++b) {
        const int s = b % STAGES;
        if (b + STAGES - 1 < nb) issue(b + STAGES - 1);
        else if (COPY_MODE != 0) asm volatile("cp.async.commit_group;" ::: "memory");
        if (COPY_MODE == 0) {
            mbar_wait(smem_u32(bars + s), (uint32_t)((b / STAGES) & 1));
        } else {
            asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 1) : "memory");
            __syncwarp();
        }
        const int cnt = min(32, total - b * 32);
        const unsigned lm = meta_last[s];
        const int myrow = meta_row[s * 32 + lane];
        const V *slots = reinterpret_cast<const V *>(stage_base + s * STAGE_BYTES) + lane;
#pragma unroll 1
        for (int j0 = 0; j0 < cnt; j0 += 4) {
            V v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = slots[(j0 + q) * 32];   // slot stride = RB bytes = 32 V's
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = j0 + q;
                if (j < cnt) {
                    vadd<VEC>(acc, v[q]);
                    if (lm & (1u << j)) {           // warp-uniform: edge j closes its row
                        const int r = __shfl_sync(0xffffffffu, myrow, j);
                        V *o = reinterpret_cast<V *>(a.Y + (item0 + r) * a.ldy) + lane;
                        V out;
                        float *op = reinterpret_cast<float *>(&out);
                        if (!SEG && a.accumulate) {
                            const V old = *o;
                            const float *pp = reinterpret_cast<const float *>(&old);
#pragma unroll
                            for (int t = 0; t < VEC; ++t) op[t] = pp[t] + acc[t];
                        } else {
#pragma unroll
                            for (int t = 0; t < VEC; ++t) op[t] = acc[t];
                        }
                        __stcs(o, out);
#pragma unroll
                        for (int t = 0; t < VEC; ++t) acc[t] = 0.f;
                    }
                }
            }
        }
        __syncwarp();
    }
}

template <int VEC, int STAGES, int COPY_MODE, bool SEG>
static cudaError_t launch_stream_one(const StreamArgs &a, cudaStream_t st) {
    const size_t smem = (size_t)STAGES * 32 * VEC * 32 * 4 + STAGES * 8 + STAGES * 4 + STAGES * 32 + 16;
    auto kern = spmm_stream_kernel<VEC, STAGES, COPY_MODE, SEG>;
    static bool configured = false;   // per instantiation
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const int64_t blocks = cdiv(a.n_items, 32);
    if (blocks == 0) return cudaSuccess;
    kern<<<(unsigned)blocks, 32, smem, st>>>(a);
    count_launch();
    return cudaGetLastError();
}

template <int VEC, bool SEG>
static cudaError_t launch_stream_vec(const StreamArgs &a, int stages, int mode, cudaStream_t st) {
    if (mode == 0) {
        if (stages <= 2) return launch_stream_one<VEC, 2, 0, SEG>(a, st);
        if (stages == 3) return launch_stream_one<VEC, 3, 0, SEG>(a, st);
        return launch_stream_one<VEC, 4, 0, SEG>(a, st);
    }
    if (stages <= 2) return launch_stream_one<VEC, 2, 1, SEG>(a, st);
    if (stages == 3) return launch_stream_one<VEC, 3, 1, SEG>(a, st);
    return launch_stream_one<VEC, 4, 1, SEG>(a, st);
}

// d in {32, 64, 128}; returns cudaErrorInvalidValue for other widths (caller falls back to vec kernel)
cudaError_t spmm_stream_launch(const StreamArgs &a, int d, bool seg, int stages, int mode, cudaStream_t st) {
    if (d == 64) return seg ? launch_stream_vec<2, true>(a, stages, mode, st) : launch_stream_vec<2, false>(a, stages, mode, st);
    if (d == 32) return seg ? launch_stream_vec<1, true>(a, stages, mode, st) : launch_stream_vec<1, false>(a, stages, mode, st);
    if (d == 128) return seg ? launch_stream_vec<4, true>(a, stages, mode, st) : launch_stream_vec<4, false>(a, stages, mode, st);
    return cudaErrorInvalidValue;
}

}  // namespace gae
