// Micro-benchmark: how long after a tcgen05.commit (or a plain mbarrier.arrive) does a waiting warp see the phase flip?
// Warp 0 signals, warp 1 waits with one of several wait loops; both stamp clock64 (same SM).  Mean over ROUNDS rounds.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/_build/bar_bench tools/bar_bench.cu -Igae_dgl_b200/csrc
#include <cstdio>
#include <cuda_runtime.h>
#include "tcgen05.cuh"
using namespace gae;

constexpr int ROUNDS = 200;

__device__ __forceinline__ bool test_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ bool try_wait_hint(uint32_t bar, uint32_t parity, uint32_t ns) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.b32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(bar), "r"(parity), "r"(ns) : "memory");
    return done != 0;
}
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
}

// signal: 0 = tcgen05.commit after one small MMA, 1 = mbarrier.arrive by one thread.  method: 0 = lane 0 try_wait + syncwarp (the decoder's
// tc_wait), 1 = all lanes try_wait, 2 = all lanes test_wait spin, 3 = lane 0 test_wait spin + syncwarp, 4 = all lanes try_wait with a 32 ns hint
__global__ void __launch_bounds__(64, 1) bench(int signal, int method, int delay, long long *out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ uint64_t bar, back;
    __shared__ volatile long long t_sig;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 16 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    const uint32_t b = tc_smem_u32(&bar), bb = tc_smem_u32(&back);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 32;" ::"r"(bb) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&tmem_slot)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    long long total = 0, worst = 0;
    for (int it = 0; it < ROUNDS; ++it) {
        const uint32_t ph = (uint32_t)(it & 1);
        if (warp == 0) {
            // let the waiter settle into its wait first
            const long long t0 = clock64();
            while (clock64() - t0 < delay) {}
            if (signal == 0) {
                if (tc_elect_one()) {
                    tc_mma_ss_f16(tmem, tc_desc(tc_smem_u32(smem), 128, 256), tc_desc(tc_smem_u32(smem + 8192), 128, 256), tc_idesc(128, 32, 0u), 0);
                    t_sig = clock64();
                    tc_commit(b);
                }
            } else if (lane == 0) {
                t_sig = clock64();
                tc_arrive(b);
            }
            __syncwarp();
            // wait for the waiter's acknowledgement
            while (!test_wait(bb, ph)) {}
        } else {
            if (method == 0) { uint32_t e = 0; tc_wait(b, ph, &e); }
            else if (method == 1) { while (!try_wait(b, ph)) {} }
            else if (method == 2) { while (!test_wait(b, ph)) {} }
            else if (method == 3) { if (lane == 0) while (!test_wait(b, ph)) {} __syncwarp(); }
            else { while (!try_wait_hint(b, ph, 32)) {} }
            const long long t1 = clock64();
            const long long d = t1 - t_sig;
            total += d;
            worst = d > worst ? d : worst;
            tc_arrive(bb);
        }
    }
    if (threadIdx.x == 32) { out[0] = total / ROUNDS; out[1] = worst; }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
}

int main() {
    long long *d;
    cudaMalloc(&d, 64);
    const char *sig[] = {"tcgen05.commit after 1 MMA", "mbarrier.arrive"};
    const char *meth[] = {"lane 0 try_wait + syncwarp (tc_wait)", "all lanes try_wait", "all lanes test_wait spin", "lane 0 test_wait spin + syncwarp", "all lanes try_wait, 32 ns hint"};
    for (int s = 0; s < 2; ++s)
        for (int m = 0; m < 5; ++m)
            for (int delay : {0, 3000}) {
                long long h[2] = {0, 0};
                bench<<<1, 64, 16 * 1024>>>(s, m, delay, d);
                cudaError_t e = cudaDeviceSynchronize();
                cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                printf("%-28s | %-38s | signal %4d cycles after the wait began: seen after mean %5lld, worst %6lld cycles %s\n", sig[s], meth[m], delay, h[0], h[1],
                       e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
    return 0;
}
