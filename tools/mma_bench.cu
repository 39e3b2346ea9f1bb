// Micro-benchmark: cycles per tcgen05.mma (kind::f16, M = 128, cta_group::1) for the shapes the decoder issues.
// One CTA; one elected lane issues REPS x n MMAs back to back, commits, waits; clock64 around issue and around completion.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/mma_bench tools/mma_bench.cu -Igae_dgl_b200/csrc
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "tcgen05.cuh"
using namespace gae;

struct Case { int n; int a_tmem; int lbo_a; int n_acc; const char *name; };

__global__ void __launch_bounds__(128, 1) bench(int N, int a_tmem, int lbo_a, int n_acc, int count, long long *out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ uint64_t bar;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;   // fp16 ones
    const uint32_t b = tc_smem_u32(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (warp == 0) {
        const uint32_t idesc = tc_idesc(128, N, 0u);
        // A: [128 rows][K = 16] K-major, core matrices at LBO (K groups) / SBO (row groups); B: [N rows][K = 16]
        const uint64_t da = tc_desc(tc_smem_u32(smem), (uint32_t)lbo_a, (uint32_t)(32 * lbo_a));
        const uint64_t db = tc_desc(tc_smem_u32(smem + 100 * 1024), 128, 256);
        long long t0 = 0, t1 = 0, t2 = 0;
        t0 = clock64();
        if (tc_elect_one()) {
            for (int i = 0; i < count; ++i) {
                const uint32_t acc = tmem + 256u + 32u * (uint32_t)(i % n_acc);
                if (a_tmem) tc_mma_ts_f16(acc, tmem + 8u * (uint32_t)(i & 15), db, idesc, 1);
                else tc_mma_ss_f16(acc, da + (uint64_t)((i & 15) * (2 * lbo_a / 16)), db, idesc, 1);
            }
            tc_commit(b);
        }
        __syncwarp();
        t1 = clock64();
        tc_wait(b, 0, reinterpret_cast<uint32_t *>(out + 4));
        t2 = clock64();
        if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
    long long *d;
    cudaMalloc(&d, 64);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const Case cases[] = {
        {128, 0, 128, 1, "SS N=128 (S type), one accumulator"}, {128, 0, 128, 2, "SS N=128, two accumulators"},
        {256, 0, 128, 1, "SS N=256, one accumulator"},
        {32, 1, 128, 1, "TS N=32 (G_I hi), one accumulator"},   {32, 1, 128, 4, "TS N=32, four accumulators"},
        {16, 1, 128, 1, "TS N=16 (G_I lo), one accumulator"},   {64, 1, 128, 1, "TS N=64, one accumulator"},
        {32, 0, 144, 1, "SS N=32, A LBO 144 (G_J), one acc"},   {32, 0, 144, 4, "SS N=32, A LBO 144, four acc"},
        {32, 0, 128, 1, "SS N=32, A LBO 128, one acc"},         {64, 0, 128, 1, "SS N=64, A LBO 128, one acc"},
        {8, 1, 128, 1, "TS N=8, one accumulator"},
    };
    for (const Case &c : cases)
        for (int count : {64, 256}) {
            long long h[8] = {0};
            cudaMemset(d, 0, 64);
            bench<<<1, 128, 200 * 1024>>>(c.n, c.a_tmem, c.lbo_a, c.n_acc, count, d);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
            printf("%-40s count %4d: issue %7lld cyc (%.1f / MMA), done %7lld cyc (%.1f / MMA) %s timeouts %lld\n", c.name, count, h[0],
                   (double)h[0] / count, h[1], (double)h[1] / count, e == cudaSuccess ? "" : cudaGetErrorString(e), h[4]);
        }
    return 0;
}
