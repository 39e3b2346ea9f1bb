// Whole GAE train step (encoder fwd, fused decoder loss+grad, encoder bwd) behind ONE C call.
//
// On molecule-batch sizes (train_inductive.py: N ~ 3-6 k) a step is ~25 kernels of a few
// microseconds; driving them one by one from Python costs several times the GPU time.  This entry
// point sequences the same kernels (it only calls the public C ABI functions of this library) on
// the caller's stream out of one caller-owned workspace, so the host cost of a step is a single
// FFI call.  Reference flow covered: gae.py:49-55 (forward), train_inductive.py:44-51 (loss,
// backward).  The optimiser step stays with the caller.
#include "common.cuh"

namespace gae {

static inline int64_t up256(int64_t x) { return (x + 255) / 256 * 256; }
static inline int64_t ld4(int64_t d) { return (d + 3) / 4 * 4; }

struct Bump {
    char *base;
    int64_t off;
    explicit Bump(void *p) : base((char *)p), off(0) {}
    template <typename T> T *take(int64_t count) {
        T *p = base ? reinterpret_cast<T *>(base + off) : nullptr;
        off += up256((int64_t)sizeof(T) * count);
        return p;
    }
};

struct StepBuffers {
    float *Y[GAE_MAX_LAYERS];    // aggregated inputs  A H_{l-1}
    float *H[GAE_MAX_LAYERS];    // layer outputs (H[L-1] is the caller's Z_out)
    float *dH[GAE_MAX_LAYERS];   // gradients w.r.t. layer outputs
    float *dY[GAE_MAX_LAYERS];   // gradients w.r.t. aggregated inputs (l > 0)
    float *Zd, *dZd;
    uint8_t *mask;
    float *hub_ws, *hub_ws_t;
    void *lin_ws;
    int64_t lin_ws_bytes;
    void *dec_ws;
    int64_t dec_ws_bytes;
};

static int64_t carve(const gae_step_desc_t *d, int64_t n, const gae_hub_plan_t *plan, const gae_hub_plan_t *plan_t,
                     void *ws, StepBuffers *out) {
    Bump b(ws);
    const int L = d->n_layers;
    int max_d = 0;
    for (int l = 0; l <= L; ++l) max_d = d->dims[l] > max_d ? d->dims[l] : max_d;
    StepBuffers sb{};
    for (int l = 0; l < L; ++l) {
        sb.Y[l] = (l == 0 && d->x_aggregated) ? nullptr : b.take<float>(n * ld4(d->dims[l]));
        sb.H[l] = (l == L - 1) ? nullptr : b.take<float>(n * ld4(d->dims[l + 1]));
        sb.dH[l] = b.take<float>(n * ld4(d->dims[l + 1]));
        sb.dY[l] = (l == 0) ? nullptr : b.take<float>(n * ld4(d->dims[l]));
    }
    const int dz = d->dims[L];
    sb.Zd = b.take<float>(n * ld4(dz));
    sb.dZd = b.take<float>(n * ld4(dz));
    sb.mask = b.take<uint8_t>(n * dz);
    sb.hub_ws = (plan && plan->n_seg > 0) ? b.take<float>(plan->n_seg * ld4(max_d)) : nullptr;
    sb.hub_ws_t = (plan_t && plan_t->n_seg > 0) ? b.take<float>(plan_t->n_seg * ld4(max_d)) : nullptr;
    sb.lin_ws_bytes = 0;
    for (int l = 0; l < L; ++l) {
        const int64_t w = gae_linear_bwd_ws_bytes(n, d->dims[l], d->dims[l + 1]);
        sb.lin_ws_bytes = w > sb.lin_ws_bytes ? w : sb.lin_ws_bytes;
    }
    sb.lin_ws = b.take<char>(sb.lin_ws_bytes);
    sb.dec_ws_bytes = d->per_graph ? gae_decoder_blockdiag_ws_bytes(n, dz) : gae_decoder_ws_bytes(n, dz);
    sb.dec_ws = b.take<char>(sb.dec_ws_bytes);
    if (out) *out = sb;
    return b.off;
}

}  // namespace gae

using namespace gae;

extern "C" int64_t gae_step_ws_bytes(const gae_step_desc_t *desc, int64_t n, const gae_hub_plan_t *plan,
                                     const gae_hub_plan_t *plan_t) {
    if (!desc || n <= 0 || desc->n_layers < 1 || desc->n_layers > GAE_MAX_LAYERS) return 0;
    return carve(desc, n, plan, plan_t, nullptr, nullptr);
}

extern "C" int gae_step_fwd_bwd_f32(const gae_step_desc_t *desc, int64_t n, const int64_t *rowptr, const int32_t *col,
                                    const gae_hub_plan_t *plan, const int64_t *rowptr_t, const int32_t *col_t,
                                    const gae_hub_plan_t *plan_t, const float *X, int64_t ldx,
                                    const float *const *W, const float *const *b, const uint8_t *mask_in,
                                    uint64_t *rng_state, const int64_t *blk_lo, const int64_t *blk_hi, double n_pairs,
                                    int32_t want_grad, float *loss, float *Z_out, int64_t ldz, float *const *dW,
                                    float *const *db, void *ws, int64_t ws_bytes, void *stream) {
    GAE_CHECK_ARG(desc && desc->n_layers >= 1 && desc->n_layers <= GAE_MAX_LAYERS, "bad layer count");
    GAE_CHECK_ARG(n > 0 && rowptr && X && W && b && loss && Z_out, "null pointer / empty graph");
    GAE_CHECK_ARG(mask_in || rng_state, "either a keep-mask or a device RNG state is required");
    GAE_CHECK_ARG(!want_grad || (rowptr_t && dW && db), "gradients need CSR(A^T), dW and db");
    GAE_CHECK_ARG(!desc->per_graph || (blk_lo && blk_hi && n_pairs > 0), "per-graph decoder needs block ranges");
    const int L = desc->n_layers;
    const int dz = desc->dims[L];
    GAE_CHECK_ARG(ldz >= dz && ldz % 4 == 0, "ldz must be a multiple of 4 and >= d_last");
    StepBuffers sb;
    const int64_t need = carve(desc, n, plan, plan_t, ws, &sb);
    if (!ws || ws_bytes < need) {
        set_error("step workspace too small: have %lld need %lld", (long long)ws_bytes, (long long)need);
        return GAE_ERR_WORKSPACE;
    }
    GAE_CHECK_ARG(aligned16(ws), "workspace must be 16-byte aligned");
    int rc;
    // ---- encoder forward: aggregate, then Linear + activation (gae.py:26-31) -----------------------
    const float *h = X;
    int64_t ldh = ldx;
    for (int l = 0; l < L; ++l) {
        const int din = desc->dims[l], dout = desc->dims[l + 1];
        const bool pre = l == 0 && desc->x_aggregated;       // X is A X already
        const float *y = pre ? X : sb.Y[l];
        const int64_t ldy = pre ? ldx : ld4(din);
        float *out = (l == L - 1) ? Z_out : sb.H[l];
        const int64_t ldo = (l == L - 1) ? ldz : ld4(dout);
        // the whole layer in one launch where the aggregated row fits a warp's registers and no row needs the hub-segment
        // plan (gcn_layer.cu); Y is kept only when the backward pass will read it
        const bool fused = !pre && tuning(T_GCN_FUSED) != 0 && din <= 64 && dout <= 64 && (!plan || plan->n_long == 0) &&
                           ldh % 4 == 0 && aligned16(h);
        if (fused) {
            rc = gae_gcn_layer_fwd_f32(rowptr, col, h, ldh, W[l], b[l], out, ldo, want_grad ? sb.Y[l] : nullptr, ld4(din), n, din, dout,
                                       desc->acts[l], stream);
            if (rc) return rc;
            h = out;
            ldh = ldo;
            continue;
        }
        if (!pre) {
            rc = gae_spmm_csr_f32(rowptr, col, nullptr, h, ldh, sb.Y[l], ld4(din), n, din, plan, sb.hub_ws, 0, stream);
            if (rc) return rc;
        }
        rc = gae_linear_fwd_f32(y, ldy, W[l], b[l], out, ldo, n, din, dout, desc->acts[l], stream);
        if (rc) return rc;
        h = out;
        ldh = ldo;
    }
    // ---- decoder: dropout (always on, gae.py:70) + fused BCE loss and unit gradient ------------------
    if (mask_in) {
        GAE_CUDA(cudaMemcpyAsync(sb.mask, mask_in, (size_t)n * dz, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
        rc = gae_dropout_fwd_f32(Z_out, ldz, sb.Zd, ld4(dz), sb.mask, n, dz, desc->dropout_p, 0, 0, 1, stream);
    } else {
        rc = gae_dropout_fwd_devrng_f32(Z_out, ldz, sb.Zd, ld4(dz), sb.mask, n, dz, desc->dropout_p, rng_state, stream);
    }
    if (rc) return rc;
    const int mode = GAE_DEC_LOSS | (want_grad ? GAE_DEC_GRAD : 0);
    if (desc->per_graph)
        rc = gae_decoder_bce_blockdiag_f32(sb.Zd, ld4(dz), n, dz, rowptr, col, rowptr_t, col_t, blk_lo, blk_hi, n_pairs,
                                           desc->pos_weight, mode, loss, sb.dZd, ld4(dz), sb.dec_ws, sb.dec_ws_bytes, stream);
    else
        rc = gae_decoder_bce_f32(sb.Zd, ld4(dz), n, dz, rowptr, col, rowptr_t, col_t, desc->pos_weight, mode, loss,
                                 sb.dZd, ld4(dz), sb.dec_ws, sb.dec_ws_bytes, stream);
    if (rc || !want_grad) return rc;
    // ---- backward (train_inductive.py:51): dropout adjoint, then layer by layer ----------------------
    rc = gae_dropout_bwd_f32(sb.dZd, ld4(dz), sb.mask, sb.dH[L - 1], ld4(dz), n, dz, desc->dropout_p, nullptr, stream);
    if (rc) return rc;
    for (int l = L - 1; l >= 0; --l) {
        const int din = desc->dims[l], dout = desc->dims[l + 1];
        const float *Hout = (l == L - 1) ? Z_out : sb.H[l];
        const int64_t ldo = (l == L - 1) ? ldz : ld4(dout);
        const bool pre = l == 0 && desc->x_aggregated;
        rc = gae_linear_bwd_f32(pre ? X : sb.Y[l], pre ? ldx : ld4(din), W[l], Hout, ldo, sb.dH[l], ld4(dout), sb.dY[l], ld4(din), dW[l], db[l],
                                sb.lin_ws, sb.lin_ws_bytes, n, din, dout, desc->acts[l], stream);
        if (rc) return rc;
        if (l > 0) {   // dH_{l-1} = A^T dY_l ; the input features are a leaf (gae.py:50)
            rc = gae_spmm_csr_f32(rowptr_t, col_t, nullptr, sb.dY[l], ld4(din), sb.dH[l - 1], ld4(din), n, din, plan_t,
                                  sb.hub_ws_t, 0, stream);
            if (rc) return rc;
        }
    }
    return GAE_OK;
}
