// K7 -- halo exchange of the 1-D vertex partition, fused with the row-block SpMMs that consume it.
//
// rank r owns a contiguous block of destination rows, their in-edge CSR over
// [local | halo] columns and the feature rows of its vertices; per SpMM every needed remote row
// crosses NVLink exactly once.  The reference has no multi-device path (train_inductive.py:26,29
// pick one GPU); this is the B200-native replacement for what DGL's distributed sampler + NCCL
// would do.
//
// Mechanism (all device-side, no NCCL collective and no host synchronisation on the data path):
//   * every rank maps its peers' [local | halo] buffers and a small FLAG block through CUDA IPC;
//   * halo rows are tagged with the first row block ("stage") of the consumer that reads them;
//   * ONE persistent push kernel per SpMM (comm stream) walks the send list stage by stage: it
//     reads the owner's rows from local HBM and stores them straight into the peers' halo regions
//     over NVSwitch (posted 128-bit stores, destinations interleaved over the peers).  When the
//     last CTA has finished stage s it publishes landed[me][s] = epoch in every peer's flag block
//     (fence.sys + st.release.sys);
//   * on the compute stream a one-warp wait kernel acquires landed[q][s] from all peers, then the
//     row-block SpMM of stage s runs (the ordinary gae_spmm_csr_f32 kernels on a sub-CSR): the
//     transfer of stage s+1.. overlaps the aggregation of stage s;
//   * when the last block is done the consumer publishes consumed[me] = epoch to its peers; the
//     next push into this buffer waits for it (write-after-read on the halo region).
// Spin loops carry a wall-clock bound (%globaltimer): on expiry they set an error word in the flag
// block and fall through, so a protocol bug produces a wrong (and reported) result, never a hang.
#include <string.h>

#include <vector>

#include "common.cuh"

namespace gae {

// flag block layout (uint64 words); block size GAE_HALO_FLAG_WORDS
__host__ __device__ __forceinline__ int landed_off(int peer, int stage) { return peer * GAE_HALO_MAX_STAGES + stage; }
__host__ __device__ __forceinline__ int consumed_off(int peer) { return GAE_HALO_MAX_WORLD * GAE_HALO_MAX_STAGES + peer; }
constexpr int HALO_ERR_OFF = GAE_HALO_MAX_WORLD * GAE_HALO_MAX_STAGES + GAE_HALO_MAX_WORLD;
// timeline of the last call (%globaltimer ns of this GPU), for tools/halo_trace.py: push start, per stage
// the time the push published it / the consumer began and ended waiting for it, and the release
constexpr int HALO_TRACE_OFF = HALO_ERR_OFF + 1;
constexpr int HALO_TRACE_WORDS = 2 + 3 * GAE_HALO_MAX_STAGES;
static_assert(HALO_TRACE_OFF + HALO_TRACE_WORDS <= GAE_HALO_FLAG_WORDS, "flag block too small");

__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint64_t *p, uint64_t v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t global_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Wait until *flag >= want.  Bounded: after timeout_ns the error word is set and false returned.
__device__ __forceinline__ bool spin_until(const uint64_t *flag, uint64_t want, uint64_t timeout_ns, uint64_t *err) {
    if (ld_acquire_sys(flag) >= want) return true;
    const uint64_t t0 = global_timer_ns();
    unsigned it = 0;
    while (ld_acquire_sys(flag) < want) {
        __nanosleep(100);
        if ((++it & 63u) == 0 && global_timer_ns() - t0 > timeout_ns) {
            atomicAdd(reinterpret_cast<unsigned long long *>(err), 1ull);
            return false;
        }
    }
    return true;
}

struct PushArgs {
    const float *X;            // my [local | halo] buffer (rows are read from the local part)
    int64_t ldx;
    const int64_t *send_src;   // [m] local row of every entry, sorted by stage, peers interleaved
    const int32_t *send_peer;  // [m]
    const int64_t *send_dst;   // [m] row in the peer's buffer
    float *const *peer_x;
    uint64_t *const *peer_flags;
    uint64_t *my_flags;
    uint32_t *stage_done;      // [n_stages] arrival counters, zero between launches
    int64_t ld_peer;
    int32_t d4, world, rank, n_stages, first_stage;
    uint64_t epoch, timeout_ns;
    int64_t stage_ptr[GAE_HALO_MAX_STAGES + 1];
};

// LPR lanes (a power of two) cover the d4 float4 of a row.  A warp takes 32 consecutive entries of
// the send list: every lane loads one entry's (src, peer, dst) -- one coalesced index load per 32
// rows -- and the rows are then moved 32/LPR at a time, U of those steps in flight, the indices
// handed around by shuffles.
template <int LPR, int U, bool STREAM>
__global__ void __launch_bounds__(512) halo_push_kernel(const PushArgs a) {
    __shared__ float *peer_base[GAE_HALO_MAX_WORLD];
    uint64_t *err = a.my_flags + HALO_ERR_OFF;
    // write-after-read: every consumer must have finished the previous SpMM on its halo region
    if ((int)threadIdx.x < a.world) {
        peer_base[threadIdx.x] = a.peer_x[threadIdx.x];
        if ((int)threadIdx.x != a.rank) spin_until(a.my_flags + consumed_off(threadIdx.x), a.epoch - 1, a.timeout_ns, err);
    }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.first_stage == 0) a.my_flags[HALO_TRACE_OFF] = global_timer_ns();
    constexpr int RPI = 32 / LPR;
    const int lane = threadIdx.x & 31, sub = lane % LPR, rsel = lane / LPR;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int s = a.first_stage; s < a.n_stages; ++s) {
        const int64_t e0 = a.stage_ptr[s], e1 = a.stage_ptr[s + 1];
        // the indices of the NEXT batch are loaded before the rows of the current one are moved
        int64_t base = e0 + warp * 32;
        int64_t nx_src = 0, nx_dst = 0;
        int nx_peer = a.rank;
        if (base + lane < e1) {
            nx_src = __ldg(a.send_src + base + lane);
            nx_dst = __ldg(a.send_dst + base + lane);
            nx_peer = __ldg(a.send_peer + base + lane);
        }
        for (; base < e1; base += n_warps * 32) {
            const int64_t my_src = nx_src, my_dst = nx_dst;
            const int my_peer = nx_peer;
            const int64_t nidx = base + n_warps * 32 + lane;
            if (nidx < e1) {
                nx_src = __ldg(a.send_src + nidx);
                nx_dst = __ldg(a.send_dst + nidx);
                nx_peer = __ldg(a.send_peer + nidx);
            }
            const int cnt = (int)min((int64_t)32, e1 - base);
            for (int k = 0; k < cnt; k += RPI * U) {
                const float4 *sp[U];
                float4 *dp[U];
                bool on[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int row = k + u * RPI + rsel;
                    const int rr = min(row, cnt - 1);
                    const int64_t src = __shfl_sync(0xffffffffu, my_src, rr);
                    const int64_t dst = __shfl_sync(0xffffffffu, my_dst, rr);
                    const int peer = __shfl_sync(0xffffffffu, my_peer, rr);
                    on[u] = row < cnt;
                    sp[u] = reinterpret_cast<const float4 *>(a.X + src * a.ldx);
                    dp[u] = reinterpret_cast<float4 *>(peer_base[peer] + dst * a.ld_peer);
                }
                for (int c = sub; c < a.d4; c += LPR) {
                    float4 v[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) v[u] = STREAM ? __ldcs(sp[u] + c) : __ldg(sp[u] + c);
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        if (on[u]) dp[u][c] = v[u];
                }
            }
        }
        // stage s complete on this CTA; the last CTA to arrive publishes it to every peer
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            const unsigned old = atomicAdd(a.stage_done + s, 1u);
            if (old == gridDim.x - 1) {
                a.stage_done[s] = 0;   // for the next launch (stream-ordered after this one)
                a.my_flags[HALO_TRACE_OFF + 2 + 3 * s] = global_timer_ns();
                __threadfence_system();
                for (int q = 0; q < a.world; ++q)
                    if (q != a.rank) st_release_sys(a.peer_flags[q] + landed_off(a.rank, s), a.epoch);
            }
        }
    }
}

__global__ void halo_wait_kernel(uint64_t *my_flags, int world, int rank, int stage, uint64_t epoch,
                                 uint64_t timeout_ns) {
    const int q = threadIdx.x;
    if (q == 0) my_flags[HALO_TRACE_OFF + 2 + 3 * stage + 1] = global_timer_ns();
    if (q < world && q != rank) spin_until(my_flags + landed_off(q, stage), epoch, timeout_ns, my_flags + HALO_ERR_OFF);
    __syncwarp();
    if (q == 0) my_flags[HALO_TRACE_OFF + 2 + 3 * stage + 2] = global_timer_ns();
}

__global__ void halo_release_kernel(uint64_t *const *peer_flags, int world, int rank, uint64_t epoch) {
    const int q = threadIdx.x;
    if (q == 0) peer_flags[rank][HALO_TRACE_OFF + 1] = global_timer_ns();
    __threadfence_system();
    if (q < world && q != rank) st_release_sys(peer_flags[q] + consumed_off(rank), epoch);
}

static int check_exchange(const gae_halo_exchange_t *ex) {
    GAE_CHECK_ARG(ex, "null exchange descriptor");
    GAE_CHECK_ARG(ex->world >= 1 && ex->world <= GAE_HALO_MAX_WORLD && ex->rank >= 0 && ex->rank < ex->world,
                  "bad world / rank");
    GAE_CHECK_ARG(ex->n_stages >= 1 && ex->n_stages <= GAE_HALO_MAX_STAGES, "1 <= n_stages <= GAE_HALO_MAX_STAGES");
    GAE_CHECK_ARG(ex->d > 0 && ex->d % 4 == 0 && ex->ld >= ex->d && ex->ld % 4 == 0, "rows must be 16-byte aligned (d, ld multiples of 4)");
    GAE_CHECK_ARG(ex->x_local && aligned16(ex->x_local) && ex->peer_x && ex->peer_flags && ex->flags, "null / unaligned buffer");
    GAE_CHECK_ARG(ex->stage_ptr && ex->stage_done, "null stage arrays");
    GAE_CHECK_ARG(ex->stage_ptr[0] == 0, "stage_ptr[0] must be 0");
    for (int s = 0; s < ex->n_stages; ++s) GAE_CHECK_ARG(ex->stage_ptr[s + 1] >= ex->stage_ptr[s], "stage_ptr must be non-decreasing");
    const int64_t m = ex->stage_ptr[ex->n_stages];
    GAE_CHECK_ARG(m == 0 || (ex->send_src && ex->send_peer && ex->send_dst), "null send lists");
    GAE_CHECK_ARG(ex->pre_n_rows >= 0, "pre_n_rows must be >= 0");
    if (ex->pre_n_rows > 0) {
        GAE_CHECK_ARG(ex->pre_rowptr && ex->pre_col && ex->pre_row0 > 0, "folding needs its CSR and staging rows");
        GAE_CHECK_ARG(ex->pre_stage >= 0 && ex->pre_stage < ex->n_stages, "pre_stage out of range");
    }
    return GAE_OK;
}

static uint64_t timeout_of(const gae_halo_exchange_t *ex) {
    return (uint64_t)(ex->timeout_ms > 0 ? ex->timeout_ms : 10000) * 1000000ull;
}

}  // namespace gae

using namespace gae;

extern "C" int gae_halo_push_f32(const gae_halo_exchange_t *ex, uint64_t epoch, void *stream) {
    int rc = check_exchange(ex);
    if (rc) return rc;
    return gae_halo_push_range_f32(ex, epoch, 0, ex->n_stages, stream);
}

extern "C" int gae_halo_push_range_f32(const gae_halo_exchange_t *ex, uint64_t epoch, int32_t stage0, int32_t stage1,
                                       void *stream) {
    int rc = check_exchange(ex);
    if (rc) return rc;
    GAE_CHECK_ARG(epoch >= 1, "epochs count from 1");
    GAE_CHECK_ARG(stage0 >= 0 && stage0 <= stage1 && stage1 <= ex->n_stages, "bad stage range");
    if (ex->world == 1 || stage0 == stage1) return GAE_OK;
    PushArgs a{};
    a.X = ex->x_local; a.ldx = ex->ld; a.send_src = ex->send_src; a.send_peer = ex->send_peer; a.send_dst = ex->send_dst;
    a.peer_x = ex->peer_x; a.peer_flags = ex->peer_flags; a.my_flags = ex->flags; a.stage_done = ex->stage_done;
    a.ld_peer = ex->ld; a.d4 = ex->d / 4; a.world = ex->world; a.rank = ex->rank; a.n_stages = stage1;
    a.first_stage = stage0;
    a.epoch = epoch; a.timeout_ns = timeout_of(ex);
    for (int s = 0; s <= ex->n_stages; ++s) a.stage_ptr[s] = ex->stage_ptr[s];
    // a range with no entries still publishes its stages
    int ctas = ex->push_ctas > 0 ? ex->push_ctas : 64;
    int threads = ex->push_threads > 0 ? ex->push_threads : 256;
    if (threads > 512) threads = 512;
    threads = (threads + 31) / 32 * 32;
    // a few dozen CTAs saturate NVLink; more would only take SMs from the row-block SpMMs running beside it
    int dev = 0, sms = 0;
    GAE_CUDA(cudaGetDevice(&dev));
    GAE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (ctas > 2 * sms) ctas = 2 * sms;
    cudaStream_t st = (cudaStream_t)stream;
    const bool u8 = tuning(T_PUSH_UNROLL) >= 8, cs_ld = tuning(T_PUSH_STREAM_LD) != 0;
#define GAE_PUSH(LPR, U)                                                              \
    do {                                                                              \
        if (cs_ld) halo_push_kernel<LPR, U, true><<<ctas, threads, 0, st>>>(a);       \
        else halo_push_kernel<LPR, U, false><<<ctas, threads, 0, st>>>(a);            \
    } while (0)
    if (a.d4 <= 4) GAE_PUSH(4, 2);
    else if (a.d4 <= 8) GAE_PUSH(8, 4);
    else if (a.d4 <= 16) { if (u8) GAE_PUSH(16, 8); else GAE_PUSH(16, 4); }
    else { if (u8) GAE_PUSH(32, 8); else GAE_PUSH(32, 4); }
#undef GAE_PUSH
    GAE_LAUNCH_CHECK();
    return GAE_OK;
}

extern "C" int gae_halo_wait_f32(const gae_halo_exchange_t *ex, int32_t stage, uint64_t epoch, void *stream) {
    int rc = check_exchange(ex);
    if (rc) return rc;
    GAE_CHECK_ARG(stage >= 0 && stage < ex->n_stages && epoch >= 1, "bad stage / epoch");
    if (ex->world == 1) return GAE_OK;
    halo_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(ex->flags, ex->world, ex->rank, stage, epoch, timeout_of(ex));
    GAE_LAUNCH_CHECK();
    return GAE_OK;
}

extern "C" int gae_halo_release_f32(const gae_halo_exchange_t *ex, uint64_t epoch, void *stream) {
    int rc = check_exchange(ex);
    if (rc) return rc;
    if (ex->world == 1) return GAE_OK;
    halo_release_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(ex->peer_flags, ex->world, ex->rank, epoch);
    GAE_LAUNCH_CHECK();
    return GAE_OK;
}

extern "C" int gae_halo_spmm_f32(const gae_halo_exchange_t *ex, const gae_halo_block_t *blocks, float *Y, int64_t ldy,
                                 uint64_t epoch, void *compute_stream, void *comm_stream, void *aux_stream) {
    int rc = check_exchange(ex);
    if (rc) return rc;
    GAE_CHECK_ARG(blocks && Y && ldy >= ex->d, "null blocks / output");
    GAE_CHECK_ARG(epoch >= 1, "epochs count from 1");
    cudaStream_t cs = (cudaStream_t)compute_stream, ms = (cudaStream_t)comm_stream, as = (cudaStream_t)aux_stream;
    GAE_CHECK_ARG(ex->world == 1 || cs != ms, "the exchange needs its own stream");
    bool two = aux_stream != nullptr && as != cs && as != ms && ex->n_stages > 1;
    for (int s = 0; s < ex->n_stages; ++s) two = two && !blocks[s].accumulate;   // passes that add into Y are ordered
    for (int s = 1; s < ex->n_stages && two; ++s)
        GAE_CHECK_ARG(!blocks[s].partial_ws || blocks[s].partial_ws != blocks[s - 1].partial_ws,
                      "consecutive row blocks run concurrently: they need separate segment workspaces");
    cudaEvent_t ready = nullptr, pushed = nullptr, aux_done = nullptr;
    GAE_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    GAE_CUDA(cudaEventCreateWithFlags(&pushed, cudaEventDisableTiming));
    GAE_CUDA(cudaEventCreateWithFlags(&aux_done, cudaEventDisableTiming));
    // The push may start once everything queued on the compute stream (the producer of the local
    // rows) is done; it then runs ahead of the row blocks.  With folding, the stages before pre_stage
    // go out at once, the folded rows are summed meanwhile (aux stream; compute stream if there is
    // none) and the remaining stages follow.
    const bool fold = ex->pre_n_rows > 0 && ex->world > 1;
    const int split = fold ? ex->pre_stage : ex->n_stages;
    cudaEvent_t folded = nullptr;
    if (fold) GAE_CUDA(cudaEventCreateWithFlags(&folded, cudaEventDisableTiming));
    const bool have_aux = aux_stream != nullptr && as != cs && as != ms;
    rc = (int)cudaEventRecord(ready, cs);
    if (rc == GAE_OK && (two || (fold && have_aux))) rc = (int)cudaStreamWaitEvent(as, ready, 0);
    if (rc == GAE_OK && ex->world > 1) {
        rc = (int)cudaStreamWaitEvent(ms, ready, 0);
        if (rc == GAE_OK) rc = gae_halo_push_range_f32(ex, epoch, 0, split, comm_stream);
        if (rc == GAE_OK && fold) {
            void *fs = have_aux ? aux_stream : compute_stream;
            float *stage_rows = const_cast<float *>(ex->x_local) + ex->pre_row0 * ex->ld;
            rc = gae_spmm_csr_f32(ex->pre_rowptr, ex->pre_col, nullptr, ex->x_local, ex->ld, stage_rows, ex->ld,
                                  ex->pre_n_rows, ex->d, ex->pre_plan, ex->pre_ws, 0, fs);
            if (rc == GAE_OK) rc = (int)cudaEventRecord(folded, (cudaStream_t)fs);
            if (rc == GAE_OK) rc = (int)cudaStreamWaitEvent(ms, folded, 0);
            if (rc == GAE_OK) rc = gae_halo_push_range_f32(ex, epoch, split, ex->n_stages, comm_stream);
        }
        if (rc == GAE_OK) rc = (int)cudaEventRecord(pushed, ms);
    }
    // Row blocks alternate between the compute stream and the auxiliary stream: block s+1 depends on
    // its own flags only (its rows of Y and its segment workspace are private), so its first CTAs
    // fill the SMs that the tail of block s leaves idle instead of waiting behind a kernel boundary.
    for (int s = 0; s < ex->n_stages && rc == GAE_OK; ++s) {
        void *st = (two && (s & 1)) ? aux_stream : compute_stream;
        rc = gae_halo_wait_f32(ex, s, epoch, st);
        const gae_halo_block_t &b = blocks[s];
        if (rc == GAE_OK && b.n_rows > 0)
            rc = gae_spmm_csr_f32(b.rowptr, b.col, nullptr, ex->x_local, ex->ld, Y + b.row0 * ldy, ldy, b.n_rows, ex->d,
                                  b.plan, b.partial_ws, b.accumulate, st);
    }
    if (rc == GAE_OK && two) {
        rc = (int)cudaEventRecord(aux_done, as);
        if (rc == GAE_OK) rc = (int)cudaStreamWaitEvent(cs, aux_done, 0);
    }
    if (rc == GAE_OK) rc = gae_halo_release_f32(ex, epoch, compute_stream);
    // callers may overwrite their local rows after this op: the push must have read them
    if (rc == GAE_OK && ex->world > 1) rc = (int)cudaStreamWaitEvent(cs, pushed, 0);
    cudaEventDestroy(ready);
    cudaEventDestroy(pushed);
    cudaEventDestroy(aux_done);
    if (folded) cudaEventDestroy(folded);
    if (rc > 0) set_error("gae_halo_spmm_f32: CUDA error %d (%s)", rc, cudaGetErrorString((cudaError_t)rc));
    return rc;
}

extern "C" int gae_halo_status(const gae_halo_exchange_t *ex, int64_t *timeouts) {
    GAE_CHECK_ARG(ex && ex->flags && timeouts, "null pointer");
    uint64_t v = 0;
    GAE_CUDA(cudaMemcpy(&v, ex->flags + HALO_ERR_OFF, sizeof(v), cudaMemcpyDeviceToHost));
    *timeouts = (int64_t)v;
    if (v != 0) {
        set_error("halo exchange: %llu flag wait(s) timed out (a peer did not publish in time)", (unsigned long long)v);
        return GAE_ERR_TIMEOUT;
    }
    return GAE_OK;
}

// ---- host planning helpers ------------------------------------------------------------------------
// O(E + V/64): a bitmap over the global vertex range marks the remote sources this rank references;
// word-wise prefix pop-counts turn a global id into its halo slot.  Halo slots are ascending in the
// global id, hence grouped by owner (owners hold contiguous id ranges).

namespace {
struct HaloBitmap {
    std::vector<uint64_t> bits;
    std::vector<int64_t> prefix;   // set bits before word w
    int64_t lo, hi, n_global;
    int build(const int64_t *src, int64_t n_edges, const int64_t *bounds, int world, int rank) {
        n_global = bounds[world];
        lo = bounds[rank];
        hi = bounds[rank + 1];
        bits.assign((size_t)((n_global + 63) / 64), 0ull);
        for (int64_t e = 0; e < n_edges; ++e) {
            const int64_t v = src[e];
            if (v < 0 || v >= n_global) return -1;
            if (v < lo || v >= hi) bits[(size_t)(v >> 6)] |= 1ull << (v & 63);
        }
        prefix.resize(bits.size() + 1);
        int64_t acc = 0;
        for (size_t w = 0; w < bits.size(); ++w) {
            prefix[w] = acc;
            acc += __builtin_popcountll(bits[w]);
        }
        prefix[bits.size()] = acc;
        return 0;
    }
    int64_t n_halo() const { return prefix.back(); }
    int64_t slot(int64_t v) const {   // v is marked
        const uint64_t below = bits[(size_t)(v >> 6)] & ((1ull << (v & 63)) - 1ull);
        return prefix[(size_t)(v >> 6)] + __builtin_popcountll(below);
    }
    int64_t count_below(int64_t v) const {   // marked ids < v, any v in [0, n_global]
        if (v >= n_global) return n_halo();
        const uint64_t below = bits[(size_t)(v >> 6)] & ((1ull << (v & 63)) - 1ull);
        return prefix[(size_t)(v >> 6)] + __builtin_popcountll(below);
    }
};

int check_bounds(const int64_t *bounds, int world, int rank) {
    if (!bounds || world < 1 || world > GAE_HALO_MAX_WORLD || rank < 0 || rank >= world || bounds[0] != 0) return -1;
    for (int q = 0; q < world; ++q)
        if (bounds[q + 1] < bounds[q]) return -1;
    return 0;
}
}  // namespace

extern "C" int gae_halo_plan_count_host(const int64_t *src_global, int64_t n_edges, const int64_t *bounds, int32_t world,
                                        int32_t rank, int64_t *n_halo, int64_t *recv_counts) {
    GAE_CHECK_ARG(check_bounds(bounds, world, rank) == 0, "bad bounds / world / rank");
    GAE_CHECK_ARG(n_edges >= 0 && (n_edges == 0 || src_global) && n_halo && recv_counts, "null pointer");
    HaloBitmap bm;
    GAE_CHECK_ARG(bm.build(src_global, n_edges, bounds, world, rank) == 0, "a source id is outside [0, bounds[world])");
    *n_halo = bm.n_halo();
    for (int q = 0; q < world; ++q) recv_counts[q] = bm.count_below(bounds[q + 1]) - bm.count_below(bounds[q]);
    return GAE_OK;
}

extern "C" int gae_halo_plan_fill_host(const int64_t *src_global, int64_t n_edges, const int64_t *bounds, int32_t world,
                                       int32_t rank, int64_t *halo_ids, int32_t *col_local) {
    GAE_CHECK_ARG(check_bounds(bounds, world, rank) == 0, "bad bounds / world / rank");
    GAE_CHECK_ARG(n_edges >= 0 && (n_edges == 0 || (src_global && col_local)), "null pointer");
    HaloBitmap bm;
    GAE_CHECK_ARG(bm.build(src_global, n_edges, bounds, world, rank) == 0, "a source id is outside [0, bounds[world])");
    const int64_t n_local = bm.hi - bm.lo;
    GAE_CHECK_ARG(n_local + bm.n_halo() < ((int64_t)1 << 31), "[local | halo] column ids must fit int32");
    GAE_CHECK_ARG(bm.n_halo() == 0 || halo_ids, "null halo_ids");
    int64_t k = 0;
    for (size_t w = 0; w < bm.bits.size(); ++w) {
        uint64_t x = bm.bits[w];
        while (x) {
            halo_ids[k++] = (int64_t)w * 64 + __builtin_ctzll(x);
            x &= x - 1;
        }
    }
    for (int64_t e = 0; e < n_edges; ++e) {
        const int64_t v = src_global[e];
        col_local[e] = (int32_t)((v >= bm.lo && v < bm.hi) ? v - bm.lo : n_local + bm.slot(v));
    }
    return GAE_OK;
}

extern "C" int gae_halo_stage_tags_host(const int64_t *rowptr, const int32_t *col_local, int64_t n_local, int64_t n_halo,
                                        const int64_t *row_bounds, int32_t n_stages, int32_t *halo_stage) {
    GAE_CHECK_ARG(rowptr && row_bounds && n_local >= 0 && n_halo >= 0 && n_stages >= 1, "bad arguments");
    GAE_CHECK_ARG(n_halo == 0 || (halo_stage && col_local), "null pointer");
    GAE_CHECK_ARG(row_bounds[0] == 0 && row_bounds[n_stages] == n_local, "row_bounds must span [0, n_local]");
    for (int64_t h = 0; h < n_halo; ++h) halo_stage[h] = -1;
    for (int32_t b = 0; b < n_stages; ++b) {
        GAE_CHECK_ARG(row_bounds[b + 1] >= row_bounds[b], "row_bounds must be non-decreasing");
        for (int64_t e = rowptr[row_bounds[b]]; e < rowptr[row_bounds[b + 1]]; ++e) {
            const int64_t h = (int64_t)col_local[e] - n_local;
            if (h >= 0) {
                GAE_CHECK_ARG(h < n_halo, "column id outside [local | halo]");
                if (halo_stage[h] < 0) halo_stage[h] = b;
            }
        }
    }
    for (int64_t h = 0; h < n_halo; ++h) GAE_CHECK_ARG(halo_stage[h] >= 0, "a halo row is referenced by no edge");
    return GAE_OK;
}

extern "C" int gae_halo_push_lists_host(const int64_t *send_idx, const int32_t *send_stage, const int64_t *send_dst,
                                        const int64_t *send_counts, const int64_t *dst_base, int32_t world,
                                        int32_t n_stages, int64_t *out_src, int32_t *out_peer, int64_t *out_dst,
                                        int64_t *stage_ptr) {
    GAE_CHECK_ARG(send_counts && (dst_base || send_dst) && stage_ptr && world >= 1 && world <= GAE_HALO_MAX_WORLD,
                  "bad arguments");
    GAE_CHECK_ARG(n_stages >= 1 && n_stages <= GAE_HALO_MAX_STAGES, "1 <= n_stages <= GAE_HALO_MAX_STAGES");
    int64_t m = 0;
    std::vector<int64_t> first(world + 1, 0);
    for (int q = 0; q < world; ++q) {
        GAE_CHECK_ARG(send_counts[q] >= 0, "negative send count");
        m += send_counts[q];
        first[q + 1] = m;
    }
    GAE_CHECK_ARG(m == 0 || (send_idx && out_src && out_peer && out_dst), "null list");
    // entries of (stage, peer), in request order; counts first, then positions
    std::vector<int64_t> cnt((size_t)n_stages * world, 0);
    for (int q = 0; q < world; ++q)
        for (int64_t j = first[q]; j < first[q + 1]; ++j) {
            const int s = send_stage ? send_stage[j] : 0;
            GAE_CHECK_ARG(s >= 0 && s < n_stages, "send_stage out of range");
            ++cnt[(size_t)s * world + q];
        }
    std::vector<int64_t> start((size_t)n_stages * world + 1, 0);
    for (size_t i = 0; i < cnt.size(); ++i) start[i + 1] = start[i] + cnt[i];
    std::vector<int64_t> bucket((size_t)m);   // entry ids grouped by (stage, peer)
    {
        std::vector<int64_t> cur(start.begin(), start.end() - 1);
        for (int q = 0; q < world; ++q)
            for (int64_t j = first[q]; j < first[q + 1]; ++j) {
                const int s = send_stage ? send_stage[j] : 0;
                bucket[(size_t)cur[(size_t)s * world + q]++] = j;
            }
    }
    // within a stage: k-th entry of every peer, then the (k+1)-th ... so that all ranks spread
    // their stores over all destinations at any time (one GPU's NVLink ingress would otherwise
    // throttle the box: measured 170 vs 697 GB/s per rank at 8 GPUs)
    int64_t o = 0;
    for (int s = 0; s < n_stages; ++s) {
        stage_ptr[s] = o;
        int64_t longest = 0;
        for (int q = 0; q < world; ++q) longest = cnt[(size_t)s * world + q] > longest ? cnt[(size_t)s * world + q] : longest;
        for (int64_t k = 0; k < longest; ++k)
            for (int q = 0; q < world; ++q)
                if (k < cnt[(size_t)s * world + q]) {
                    const int64_t j = bucket[(size_t)(start[(size_t)s * world + q] + k)];
                    out_src[o] = send_idx[j];
                    out_peer[o] = q;
                    out_dst[o] = send_dst ? send_dst[j] : dst_base[q] + (j - first[q]);
                    ++o;
                }
    }
    stage_ptr[n_stages] = o;
    return GAE_OK;
}
