// K4 dropout, graph-indexing helpers, halo pack / one-sided pull, CUDA IPC plumbing.
#include <math.h>
#include <string.h>

#include "common.cuh"

namespace gae {

// ---- Philox4x32-10 (counter-based; one counter per 4 consecutive elements) -----------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0; key.y += W1;
    }
    return ctr;
}

// Reference gae.py:70: F.dropout(z, p) with training=True always.  Zd = Z * keep / (1-p).
__global__ void dropout_fwd_kernel(const float *__restrict__ Z, int64_t ldz, float *__restrict__ Zd,
                                   int64_t ldzd, uint8_t *__restrict__ mask, int64_t n, int d, float p,
                                   float scale, uint64_t seed, uint64_t offset, int mask_mode,
                                   const uint64_t *__restrict__ rng_state) {
    if (rng_state) {  // device-resident Philox state: CUDA-graph replays draw fresh masks
        seed = rng_state[0];
        offset = rng_state[1];
    }
    const int64_t total = n * d;
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // group of 4 elements
    if (q * 4 >= total) return;
    uint32_t rnd[4] = {0, 0, 0, 0};
    if (mask_mode == 0) {
        const uint64_t c = offset + (uint64_t)q;
        const uint4 r = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0u, 0u),
                                      make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
        rnd[0] = r.x; rnd[1] = r.y; rnd[2] = r.z; rnd[3] = r.w;
    }
    for (int t = 0; t < 4; ++t) {
        const int64_t idx = q * 4 + t;
        if (idx >= total) break;
        const int64_t row = idx / d;
        const int c = (int)(idx % d);
        uint8_t keep;
        if (mask_mode == 0) {
            const float u = (float)(rnd[t] >> 8) * (1.0f / 16777216.0f);  // [0,1)
            keep = u >= p;
            mask[idx] = keep;
        } else {
            keep = mask[idx];
        }
        Zd[row * ldzd + c] = keep ? Z[row * ldz + c] * scale : 0.f;
    }
}

__global__ void rng_advance_kernel(uint64_t *state, uint64_t inc) { state[1] += inc; }

__global__ void dropout_bwd_kernel(const float *__restrict__ dZd, int64_t ld_dzd,
                                   const uint8_t *__restrict__ mask, float *__restrict__ dZ, int64_t ld_dz,
                                   int64_t n, int d, float scale, const float *__restrict__ grad_scale) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * d) return;
    const int64_t row = idx / d;
    const int c = (int)(idx % d);
    const float gs = grad_scale ? *grad_scale : 1.f;
    dZ[row * ld_dz + c] = mask[idx] ? dZd[row * ld_dzd + c] * scale * gs : 0.f;
}

__global__ void in_degrees_kernel(const int64_t *__restrict__ rowptr, int64_t n, int64_t *__restrict__ deg) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v < n) deg[v] = rowptr[v + 1] - rowptr[v];
}

// dgl.batch: add the node offset of the graph an edge belongs to (binary search on edge ptr)
__global__ void batch_offset_cols_kernel(int32_t *__restrict__ col, const int64_t *__restrict__ edge_ptr,
                                         const int64_t *__restrict__ node_off, int64_t n_graphs, int64_t n_edges) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    int64_t lo = 0, hi = n_graphs;  // find g with edge_ptr[g] <= e < edge_ptr[g+1]
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (edge_ptr[mid] <= e) lo = mid; else hi = mid;
    }
    col[e] += (int32_t)node_off[lo];
}

// dgl.batch on device from a PACKED dataset (all member graphs resident in HBM as one CSR with
// graph-local column ids): assemble the block-diagonal union of graphs gid[0..K) -- row
// pointers, columns (+ node offset) and node features -- with no host loop over the members.
__device__ __forceinline__ int64_t find_segment(const int64_t *__restrict__ off, int64_t k_count, int64_t x) {
    int64_t lo = 0, hi = k_count;   // off[lo] <= x < off[lo+1]
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (off[mid] <= x) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void batch_assemble_kernel(const int64_t *__restrict__ rowptr_all, const int32_t *__restrict__ col_all,
                                      const int64_t *__restrict__ node_ptr, const int64_t *__restrict__ gid,
                                      int64_t K, const int64_t *__restrict__ noff, const int64_t *__restrict__ eoff,
                                      int64_t n_out, int64_t e_out, int64_t *__restrict__ out_rowptr,
                                      int32_t *__restrict__ out_col) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_out) {
        const int64_t k = find_segment(noff, K, t);
        const int64_t base = node_ptr[gid[k]];
        const int64_t src = base + (t - noff[k]);
        out_rowptr[t + 1] = eoff[k] + (rowptr_all[src + 1] - rowptr_all[base]);
        if (t == 0) out_rowptr[0] = 0;
    }
    if (t < e_out) {
        const int64_t k = find_segment(eoff, K, t);
        const int64_t e_src = rowptr_all[node_ptr[gid[k]]] + (t - eoff[k]);
        out_col[t] = col_all[e_src] + (int32_t)noff[k];
    }
}

__global__ void batch_features_kernel(const float *__restrict__ feat_all, int64_t ldf, const int64_t *__restrict__ node_ptr,
                                      const int64_t *__restrict__ gid, int64_t K, const int64_t *__restrict__ noff,
                                      int64_t n_out, int d, float *__restrict__ out, int64_t ldo) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t node = t / d;
    const int c = (int)(t % d);
    if (node >= n_out) return;
    const int64_t k = find_segment(noff, K, node);
    out[node * ldo + c] = __ldg(feat_all + (node_ptr[gid[k]] + (node - noff[k])) * ldf + c);
}

// pack rows idx[] of X (128-bit when aligned, one lane group per row)
template <bool VEC>
__global__ void gather_rows_kernel(const float *__restrict__ X, int64_t ldx, const int64_t *__restrict__ idx,
                                   int64_t m, int d, float *__restrict__ out, int64_t ldo) {
    if (VEC) {
        const int d4 = d >> 2;
        const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        const int64_t r = t / d4;
        const int c = (int)(t % d4);
        if (r >= m) return;
        const float4 v = __ldg(reinterpret_cast<const float4 *>(X + idx[r] * ldx) + c);
        __stcs(reinterpret_cast<float4 *>(out + r * ldo) + c, v);
    } else {
        const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        const int64_t r = t / d;
        const int c = (int)(t % d);
        if (r >= m) return;
        out[r * ldo + c] = X[idx[r] * ldx + c];
    }
}

// one-sided halo pull: rows live in PEER memory mapped through CUDA IPC; loads travel over
// NVLink (peer addresses bypass the local L2), stores land in the local halo buffer.
__global__ void pull_rows_p2p_kernel(const float *const *__restrict__ peers, const int32_t *__restrict__ owner,
                                     const int64_t *__restrict__ idx, int64_t m, int64_t ldx, int d4,
                                     float *__restrict__ out, int64_t ldo) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t r = t / d4;
    const int c = (int)(t % d4);
    if (r >= m) return;
    const float *base = peers[owner[r]];
    float4 v;
    const float4 *p = reinterpret_cast<const float4 *>(base + idx[r] * ldx) + c;
    asm volatile("ld.global.relaxed.sys.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    reinterpret_cast<float4 *>(out + r * ldo)[c] = v;
}

// one-sided halo PUSH: the owner reads its own rows (local HBM / L2) and stores them straight
// into each peer's halo region over NVLink.  Stores are posted (no round trip), so this sustains
// far more NVLink bandwidth than remote loads; it also fuses the "pack" step of an all-to-all
// into the transfer -- no staging buffer on either side.
__global__ void push_rows_p2p_kernel(const float *__restrict__ X, int64_t ldx, const int64_t *__restrict__ send_idx,
                                     const int32_t *__restrict__ dst_peer, const int64_t *__restrict__ dst_row,
                                     float *const *__restrict__ peers, int64_t m, int64_t ld_peer, int d4) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t r = t / d4;
    const int c = (int)(t % d4);
    if (r >= m) return;
    const float4 v = __ldg(reinterpret_cast<const float4 *>(X + send_idx[r] * ldx) + c);
    float4 *dst = reinterpret_cast<float4 *>(peers[dst_peer[r]] + dst_row[r] * ld_peer) + c;
    *dst = v;
}

// torch.optim.Adam step (train_inductive.py:40,52; no weight decay, no amsgrad) over up to
// GAE_ADAM_MAX_TENSORS parameter tensors in ONE launch.  Same update as torch's:
//   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= (lr / (1-b1^t)) * m / (sqrt(v) / sqrt(1-b2^t) + eps)
struct AdamArgs {
    float *p[GAE_ADAM_MAX_TENSORS];
    const float *g[GAE_ADAM_MAX_TENSORS];
    float *m[GAE_ADAM_MAX_TENSORS];
    float *v[GAE_ADAM_MAX_TENSORS];
    int64_t end[GAE_ADAM_MAX_TENSORS];   // exclusive prefix ends of the flattened element range
    int32_t n_tensors;
    float b1, b2, eps, step_size, inv_bias2_sqrt;
};

__global__ void adam_step_kernel(const AdamArgs a) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.end[a.n_tensors - 1]) return;
    int t = 0;
    while (i >= a.end[t]) ++t;
    const int64_t k = i - (t ? a.end[t - 1] : 0);
    const float g = a.g[t][k];
    const float m = fmaf(a.b1, a.m[t][k] - g, g);                 // lerp: g + b1 (m - g) == b1 m + (1-b1) g
    const float v = fmaf(a.b2, a.v[t][k], (1.f - a.b2) * g * g);
    a.m[t][k] = m;
    a.v[t][k] = v;
    const float denom = sqrtf(v) * a.inv_bias2_sqrt + a.eps;
    a.p[t][k] -= a.step_size * (m / denom);
}

}  // namespace gae

using namespace gae;

extern "C" int gae_adam_step_f32(int32_t n_tensors, float *const *params, const float *const *grads, float *const *exp_avg,
                                 float *const *exp_avg_sq, const int64_t *numel, float lr, float beta1, float beta2,
                                 float eps, int64_t step, void *stream) {
    GAE_CHECK_ARG(n_tensors >= 1 && n_tensors <= GAE_ADAM_MAX_TENSORS, "1 <= n_tensors <= GAE_ADAM_MAX_TENSORS");
    GAE_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && numel, "null pointer");
    GAE_CHECK_ARG(step >= 1 && lr >= 0.f && beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f,
                  "bad hyper-parameters (step counts from 1)");
    AdamArgs a{};
    int64_t total = 0;
    for (int t = 0; t < n_tensors; ++t) {
        GAE_CHECK_ARG(params[t] && grads[t] && exp_avg[t] && exp_avg_sq[t] && numel[t] > 0, "null / empty tensor");
        a.p[t] = params[t]; a.g[t] = grads[t]; a.m[t] = exp_avg[t]; a.v[t] = exp_avg_sq[t];
        total += numel[t];
        a.end[t] = total;
    }
    a.n_tensors = n_tensors;
    a.b1 = beta1; a.b2 = beta2; a.eps = eps;
    // bias corrections in double on the host, as torch computes them from the python step count
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    a.step_size = (float)((double)lr / bc1);
    a.inv_bias2_sqrt = (float)(1.0 / sqrt(bc2));
    adam_step_kernel<<<(unsigned)cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(a);
    GAE_LAUNCH_CHECK();
    return GAE_OK;
}

extern "C" int gae_dropout_fwd_f32(const float *Z, int64_t ldz, float *Zd, int64_t ldzd, uint8_t *mask,
                                   int64_t n, int32_t d, float p, uint64_t seed, uint64_t offset,
                                   int32_t mask_mode, void *stream) {
    GAE_CHECK_ARG(n >= 0 && d > 0, "bad sizes");
    GAE_CHECK_ARG(p >= 0.f && p < 1.f, "p must be in [0,1)");
    if (n == 0) return GAE_OK;
    GAE_CHECK_ARG(Z && Zd && mask, "null pointer");
    GAE_CHECK_ARG(ldz >= d && ldzd >= d, "leading dimension too small");
    const int64_t groups = cdiv(n * d, 4);
    dropout_fwd_kernel<<<(unsigned)cdiv(groups, 256), 256, 0, (cudaStream_t)stream>>>(
        Z, ldz, Zd, ldzd, mask, n, d, p, 1.0f / (1.0f - p), seed, offset, mask_mode, nullptr);
    GAE_LAUNCH_CHECK();
    return GAE_OK;
}

extern "C" int gae_dropout_fwd_devrng_f32(const float *Z, int64_t ldz, float *Zd, int64_t ldzd, uint8_t *mask,
                                          int64_t n, int32_t d, float p, uint64_t *rng_state, void *stream) {
    GAE_CHECK_ARG(n >= 0 && d > 0, "bad sizes");
    GAE_CHECK_ARG(p >= 0.f && p < 1.f, "p must be in [0,1)");
    if (n == 0) return GAE_OK;
    GAE_CHECK_ARG(Z && Zd && mask && rng_state, "null pointer");
    GAE_CHECK_ARG(ldz >= d && ldzd >= d, "leading dimension too small");
    const int64_t groups = cdiv(n * d, 4);
    cudaStream_t st = (cudaStream_t)stream;
    dropout_fwd_kernel<<<(unsigned)cdiv(groups, 256), 256, 0, st>>>(Z, ldz, Zd, ldzd, mask, n, d, p, 1.0f / (1.0f - p),
                                                                    0, 0, 0, rng_state);
    GAE_LAUNCH_CHECK();
    rng_advance_kernel<<<1, 1, 0, st>>>(rng_state, (uint64_t)groups);
    GAE_LAUNCH_CHECK();
    return GAE_OK;
}

extern "C" int gae_dropout_bwd_f32(const float *dZd, int64_t ld_dzd, const uint8_t *mask, float *dZ,
                                   int64_t ld_dz, int64_t n, int32_t d, float p, const float *grad_scale,
                                   void *stream) {
    GAE_CHECK_ARG(n >= 0 && d > 0, "bad sizes");
    if (n == 0) return GAE_OK;
    GAE_CHECK_ARG(dZd && mask && dZ, "null pointer");
    dropout_bwd_kernel<<<(unsigned)cdiv(n * d, 256), 256, 0, (cudaStream_t)stream>>>(
        dZd, ld_dzd, mask, dZ, ld_dz, n, d, 1.0f / (1.0f - p), grad_scale);
    GAE_LAUNCH_CHECK();
    return GAE_OK;
}

extern "C" int gae_in_degrees_i64(const int64_t *rowptr, int64_t n_rows, int64_t *deg, void *stream) {
    GAE_CHECK_ARG(n_rows >= 0, "bad size");
    if (n_rows == 0) return GAE_OK;
    GAE_CHECK_ARG(rowptr && deg, "null pointer");
    in_degrees_kernel<<<(unsigned)cdiv(n_rows, 256), 256, 0, (cudaStream_t)stream>>>(rowptr, n_rows, deg);
    GAE_LAUNCH_CHECK();
    return GAE_OK;
}

extern "C" int gae_batch_offset_cols_i32(int32_t *col_cat, const int64_t *edge_graph_ptr,
                                         const int64_t *node_off, int64_t n_graphs, int64_t n_edges,
                                         void *stream) {
    GAE_CHECK_ARG(n_graphs >= 0 && n_edges >= 0, "bad sizes");
    if (n_edges == 0 || n_graphs == 0) return GAE_OK;
    GAE_CHECK_ARG(col_cat && edge_graph_ptr && node_off, "null pointer");
    batch_offset_cols_kernel<<<(unsigned)cdiv(n_edges, 256), 256, 0, (cudaStream_t)stream>>>(
        col_cat, edge_graph_ptr, node_off, n_graphs, n_edges);
    GAE_LAUNCH_CHECK();
    return GAE_OK;
}

extern "C" int gae_batch_assemble(const int64_t *rowptr_all, const int32_t *col_all, const int64_t *node_ptr,
                                  const int64_t *gid, int64_t n_graphs, const int64_t *node_off,
                                  const int64_t *edge_off, int64_t n_out, int64_t e_out, int64_t *out_rowptr,
                                  int32_t *out_col, const float *feat_all, int64_t ldf, int32_t d, float *out_feat,
                                  int64_t ld_out, void *stream) {
    GAE_CHECK_ARG(n_graphs > 0 && n_out >= 0 && e_out >= 0, "bad sizes");
    GAE_CHECK_ARG(rowptr_all && node_ptr && gid && node_off && edge_off && out_rowptr, "null pointer");
    GAE_CHECK_ARG(e_out == 0 || (col_all && out_col), "null column arrays");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t work = n_out > e_out ? n_out : e_out;
    if (work > 0) {
        batch_assemble_kernel<<<(unsigned)cdiv(work, 256), 256, 0, st>>>(rowptr_all, col_all, node_ptr, gid, n_graphs,
                                                                          node_off, edge_off, n_out, e_out, out_rowptr,
                                                                          out_col);
        GAE_LAUNCH_CHECK();
    } else {
        GAE_CUDA(cudaMemsetAsync(out_rowptr, 0, sizeof(int64_t), st));
    }
    if (feat_all && out_feat && n_out > 0 && d > 0) {
        GAE_CHECK_ARG(ldf >= d && ld_out >= d, "feature leading dimension too small");
        batch_features_kernel<<<(unsigned)cdiv(n_out * d, 256), 256, 0, st>>>(feat_all, ldf, node_ptr, gid, n_graphs,
                                                                               node_off, n_out, d, out_feat, ld_out);
        GAE_LAUNCH_CHECK();
    }
    return GAE_OK;
}

extern "C" int gae_gather_rows_f32(const float *X, int64_t ldx, const int64_t *idx, int64_t m, int32_t d,
                                   float *out, int64_t ld_out, void *stream) {
    GAE_CHECK_ARG(m >= 0 && d > 0, "bad sizes");
    if (m == 0) return GAE_OK;
    GAE_CHECK_ARG(X && idx && out, "null pointer");
    const bool vec = aligned16(X) && aligned16(out) && ldx % 4 == 0 && ld_out % 4 == 0 && d % 4 == 0;
    if (vec)
        gather_rows_kernel<true><<<(unsigned)cdiv(m * (d / 4), 256), 256, 0, (cudaStream_t)stream>>>(X, ldx, idx, m, d, out, ld_out);
    else
        gather_rows_kernel<false><<<(unsigned)cdiv(m * d, 256), 256, 0, (cudaStream_t)stream>>>(X, ldx, idx, m, d, out, ld_out);
    GAE_LAUNCH_CHECK();
    return GAE_OK;
}

extern "C" int gae_pull_rows_p2p_f32(const float *const *peer_ptrs, const int32_t *owner, const int64_t *idx,
                                     int64_t m, int64_t ldx, int32_t d, float *out, int64_t ld_out,
                                     void *stream) {
    GAE_CHECK_ARG(m >= 0 && d > 0, "bad sizes");
    if (m == 0) return GAE_OK;
    GAE_CHECK_ARG(peer_ptrs && owner && idx && out, "null pointer");
    GAE_CHECK_ARG(d % 4 == 0 && ldx % 4 == 0 && ld_out % 4 == 0 && aligned16(out), "p2p pull needs 16-byte aligned rows");
    pull_rows_p2p_kernel<<<(unsigned)cdiv(m * (d / 4), 256), 256, 0, (cudaStream_t)stream>>>(
        peer_ptrs, owner, idx, m, ldx, d / 4, out, ld_out);
    GAE_LAUNCH_CHECK();
    return GAE_OK;
}

extern "C" int gae_push_rows_p2p_f32(const float *X, int64_t ldx, const int64_t *send_idx, const int32_t *dst_peer,
                                     const int64_t *dst_row, float *const *peer_ptrs, int64_t m, int64_t ld_peer,
                                     int32_t d, void *stream) {
    GAE_CHECK_ARG(m >= 0 && d > 0, "bad sizes");
    if (m == 0) return GAE_OK;
    GAE_CHECK_ARG(X && send_idx && dst_peer && dst_row && peer_ptrs, "null pointer");
    GAE_CHECK_ARG(d % 4 == 0 && ldx % 4 == 0 && ld_peer % 4 == 0 && aligned16(X), "p2p push needs 16-byte aligned rows");
    push_rows_p2p_kernel<<<(unsigned)cdiv(m * (d / 4), 256), 256, 0, (cudaStream_t)stream>>>(
        X, ldx, send_idx, dst_peer, dst_row, peer_ptrs, m, ld_peer, d / 4);
    GAE_LAUNCH_CHECK();
    return GAE_OK;
}

extern "C" int gae_ipc_get_handle(const void *dev_ptr, uint8_t handle_out[64], int64_t *offset_out) {
    GAE_CHECK_ARG(dev_ptr && handle_out && offset_out, "null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaPointerAttributes attr;
    GAE_CUDA(cudaPointerGetAttributes(&attr, dev_ptr));
    GAE_CHECK_ARG(attr.type == cudaMemoryTypeDevice, "not a device pointer");
    // The handle names the whole allocation; report where dev_ptr sits inside it.  The driver
    // entry point is resolved at run time so the library has no link-time libcuda dependency.
    typedef int (*get_range_fn)(unsigned long long *, size_t *, unsigned long long);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    GAE_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) {
        set_error("cuMemGetAddressRange not available");
        return GAE_ERR_UNSUPPORTED;
    }
    unsigned long long base = 0;
    size_t size = 0;
    const int drc = ((get_range_fn)fn)(&base, &size, (unsigned long long)(uintptr_t)dev_ptr);
    if (drc != 0) {
        set_error("cuMemGetAddressRange failed: %d", drc);
        return GAE_ERR_INVALID_ARG;
    }
    cudaIpcMemHandle_t h;
    GAE_CUDA(cudaIpcGetMemHandle(&h, reinterpret_cast<void *>((uintptr_t)base)));
    memcpy(handle_out, &h, 64);
    *offset_out = (int64_t)((unsigned long long)(uintptr_t)dev_ptr - base);
    return GAE_OK;
}

extern "C" int gae_ipc_open_handle(const uint8_t handle[64], void **dev_ptr_out) {
    GAE_CHECK_ARG(handle && dev_ptr_out, "null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    GAE_CUDA(cudaIpcOpenMemHandle(dev_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return GAE_OK;
}

extern "C" int gae_ipc_close_handle(void *dev_ptr) {
    GAE_CHECK_ARG(dev_ptr, "null pointer");
    GAE_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return GAE_OK;
}
