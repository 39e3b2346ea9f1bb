/* gae_b200.h -- C ABI of libgae_b200.so: the B200 (sm_100a) GAE encoder/decoder hot path.
 *
 * The reference (shionhonda/gae-dgl) has no FFI layer of its own: its hot path reaches
 * native code through DGL's and PyTorch's operators.  Each entry point below replaces one
 * of those operator calls; the citation is the reference call site (paths relative to
 * /root/reference).  INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions
 *   - Every pointer named *_dev / without suffix is DEVICE memory unless the function name
 *     ends in _host.  The caller owns every buffer, including workspaces; the library never
 *     allocates or frees device memory (exception: the gae_ipc_* helpers map PEER memory).
 *   - All launches are asynchronous on `stream` (a cudaStream_t passed as void*; NULL = the
 *     legacy default stream).  The device is the caller's current device.
 *   - Return value: 0 = ok; negative = invalid argument (GAE_ERR_*); positive = cudaError_t.
 *     gae_last_error_string() gives a per-thread human-readable message.  No C++ exception
 *     crosses this boundary.  The library keeps no global mutable state besides the launch
 *     counter: tuning knobs (gae_set_tuning) and the error string are per host thread; it is
 *     re-entrant.
 *   - Matrices are row-major fp32 with an explicit leading dimension (elements).  The
 *     128-bit vector paths are taken when base pointers are 16-byte aligned and leading
 *     dimensions are multiples of 4; otherwise a scalar path computes the same result.
 *   - Graph indexing: CSR over DESTINATION rows, rowptr int64 [n_rows+1], col int32 [E] =
 *     source ids, duplicates kept (multigraph), i.e. A[v,u] = #edges u->v, which is what
 *     gae.py:18-19 sums and what DGL 0.4 adjacency_matrix() returns (train_inductive.py:44).
 */
#ifndef GAE_B200_H_
#define GAE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GAE_OK 0
#define GAE_ERR_INVALID_ARG (-1)
#define GAE_ERR_UNSUPPORTED (-2)
#define GAE_ERR_WORKSPACE (-3)
#define GAE_ERR_TIMEOUT (-4)

#define GAE_ACT_IDENTITY 0 /* gae.py:43,45  lambda x: x */
#define GAE_ACT_RELU 1     /* gae.py:36-41  F.relu      */

/* ---- library ------------------------------------------------------------------------ */
const char *gae_version(void);
const char *gae_last_error_string(void);
/* Tuning knobs used by bench sweeps.  Keys (default): "spmm_variant" (0: register gather; 1 / 2:
 * streaming cp.async.bulk / LDGSTS), "spmm_unroll" (4), "spmm_block" (64), "spmm_cache" (0),
 * "spmm_rows_per_warp" (1), "spmm_stages" (2), "spmm_bins" (1), "spmm_seg_order" (1), "spmm_fused"
 * (-1 = automatic; 1 / 0: single-launch form of the binned forward on / off),
 * "dec_splits" (0 = auto), "dec_rows" (2), "dec_tc" (-1: by size -- the tcgen05 / TMEM fp16-split pass
 * for d <= 16 from 5632 rows; 2 / 1: that pass / its TF32 predecessor from 512 rows; 0: never),
 * "dec_mma" (1: below that, dense pass as mma.sync TF32 MMAs for d <= 16; 0: SIMT), "push_unroll" (4),
 * "push_stream_ld" (1), "gcn_fused" (1: gae_step_fwd_bwd_f32 runs its layers through gae_gcn_layer_fwd_f32
 * where it applies; 0: SpMM + Linear).  The knobs are PER HOST THREAD.
 * Results are independent of every knob up to fp32 summation order.
 * Unknown keys return GAE_ERR_INVALID_ARG.  gae_get_tuning returns the value or -1. */
int gae_set_tuning(const char *key, int32_t value);
int32_t gae_get_tuning(const char *key);
/* Number of kernels this library has launched from the calling process (for bench.py's
 * "gpu_launches" claim). */
int64_t gae_launch_count(void);

/* ---- hub-row plan (host helper) ------------------------------------------------------- */
/* Rows whose in-degree exceeds seg_len are split into ceil(deg/seg_len) segments that are
 * summed by separate warps into a partial buffer and then reduced in fixed order
 * (deterministic, atomics-free).  Two-call protocol on HOST rowptr:
 *   1) gae_hub_plan_count_host -> n_long, n_seg
 *   2) gae_hub_plan_fill_host  -> long_row[n_long], long_seg_ptr[n_long+1], seg_row[n_seg]
 * The caller uploads the three arrays and passes them in gae_hub_plan_t. */
int gae_hub_plan_count_host(const int64_t *rowptr_host, int64_t n_rows, int32_t seg_len,
                            int64_t *n_long, int64_t *n_seg);
int gae_hub_plan_fill_host(const int64_t *rowptr_host, int64_t n_rows, int32_t seg_len,
                           int32_t *long_row, int64_t *long_seg_ptr, int32_t *seg_row);

/* Degree bins, same two-call protocol: pass NULL arrays to get counts[3] = {n_empty, n_short,
 * n_mid}, then arrays of those sizes (row ids ascending within each bin). */
int gae_row_bins_host(const int64_t *rowptr_host, int64_t n_rows, int32_t seg_len,
                      int32_t short_max, int64_t counts[3], int32_t *empty_rows,
                      int32_t *short_rows, int32_t *mid_rows);

typedef struct gae_hub_plan_t {
    int32_t seg_len;             /* edges per segment (> 0)                       */
    int32_t short_max;           /* degree bins: rows with 1..short_max in-edges are "short" */
    int64_t n_long;              /* rows with deg > seg_len                       */
    int64_t n_seg;               /* total segments over those rows                */
    const int32_t *long_row;     /* [n_long]   row id of k-th long row (ascending) */
    const int64_t *long_seg_ptr; /* [n_long+1] prefix sum of segments per long row */
    const int32_t *seg_row;      /* [n_seg]    index k of the long row a segment belongs to */
    /* Optional degree bins (all NULL / 0 = the row kernel walks every row).  On skewed graphs
     * most rows are empty or tiny (R-MAT C4: 71 % of the rows hold 2.4 % of the edges) and a
     * warp per row wastes the machine on them: empty rows get a streaming zero fill, short rows
     * run four to a warp with all their gathers in flight at once, the rest keep a warp each. */
    int64_t n_empty, n_short, n_mid;
    const int32_t *empty_rows;   /* [n_empty] rows with in-degree 0                         */
    const int32_t *short_rows;   /* [n_short] rows with in-degree in [1, short_max]          */
    const int32_t *mid_rows;     /* [n_mid]   rows with in-degree in (short_max, seg_len]    */
    /* Optional processing order of the segments (a permutation of [0, n_seg), NULL = row-major).
     * Ordering them by their first source id makes concurrently running warps gather from the same
     * narrow band of X, which then stays resident in L2 (results are unchanged: every segment still
     * writes its own partial row, reduced in the same fixed order). */
    const int32_t *seg_order;
} gae_hub_plan_t;

/* ---- K1 / K2: CSR SpMM, Y = A X (sum aggregation) -------------------------------------- */
/* Replaces  g.update_all(fn.copy_src('h','m'), fn.sum('m','h'))   gae.py:18-19,28
 * and, called with CSR(A^T), its autograd adjoint dX = A^T dY      train_inductive.py:51.
 *   vals       : optional per-edge weights [E] (NULL = unweighted, the reference's case)
 *   plan       : optional hub-row plan (NULL = every row is summed by one warp)
 *   partial_ws : [plan->n_seg, d] floats when plan != NULL and plan->n_seg > 0
 *   accumulate : 0 -> Y = A X ; 1 -> Y += A X (used for the remote-source pass, 8e) */
int gae_spmm_csr_f32(const int64_t *rowptr, const int32_t *col, const float *vals,
                     const float *X, int64_t ldx, float *Y, int64_t ldy, int64_t n_rows,
                     int32_t d, const gae_hub_plan_t *plan, float *partial_ws,
                     int32_t accumulate, void *stream);

/* Host-buffer variant (the e2e entry point): X_host -> device staging -> SpMM -> Y_host.
 * rowptr/col/plan are DEVICE-resident graph state (the graph persists across epochs in the
 * reference, train_transductive.py:45; features are re-assigned every epoch).  X_host and
 * Y_host should be pinned.  X_stage [n_src,ldx] and Y_stage [n_rows,ldy] are caller-owned
 * device buffers.  Copies and kernels are enqueued on `stream`; returns without syncing. */
int gae_spmm_csr_f32_host(const int64_t *rowptr, const int32_t *col, const float *X_host,
                          int64_t n_src, int64_t ldx, float *Y_host, int64_t ldy,
                          int64_t n_rows, int32_t d, const gae_hub_plan_t *plan,
                          float *partial_ws, float *X_stage, float *Y_stage, void *stream);

/* ---- K3: NodeApplyModule = Linear + activation ------------------------------------------ */
/* Replaces  h = self.linear(node.data['h']); h = self.activation(h)   gae.py:13-16.
 * H[n,d_out] = act(Yin[n,d_in] W^T + b), W is [d_out,d_in] row-major (nn.Linear layout). */
int gae_linear_fwd_f32(const float *Yin, int64_t ld_in, const float *W, const float *b,
                       float *H, int64_t ld_out, int64_t n, int32_t d_in, int32_t d_out,
                       int32_t act, void *stream);
/* Adjoint.  H is the forward OUTPUT (post-activation; the ReLU mask is H > 0).
 * dYin may be NULL (first layer: input features are a leaf, gae.py:50).
 * ws: gae_linear_bwd_ws_bytes() bytes.  dW [d_out,d_in], db [d_out] are overwritten. */
int64_t gae_linear_bwd_ws_bytes(int64_t n, int32_t d_in, int32_t d_out);
int gae_linear_bwd_f32(const float *Yin, int64_t ld_in, const float *W, const float *H,
                       int64_t ld_out, const float *dH, int64_t ld_dh, float *dYin,
                       int64_t ld_dyin, float *dW, float *db, void *ws, int64_t ws_bytes,
                       int64_t n, int32_t d_in, int32_t d_out, int32_t act, void *stream);

/* ---- K1 + K3 in one launch: a whole GCN layer --------------------------------------------- */
/* Replaces  g.update_all(gcn_msg, gcn_reduce); g.apply_nodes(func=self.apply_mod)   gae.py:26-31
 * (SURVEY.md 8b `gae_gcn_layer_fwd/bwd_f32`):  Hout[n,d_out] = act((A Hin) W^T + b) without the round trip
 * of Y = A Hin through HBM.  d_in, d_out <= 64; Hin (and Y) rows 16-byte aligned, row strides multiples of
 * 4 floats >= d_in rounded up to 4 (else GAE_ERR_UNSUPPORTED: use gae_spmm_csr_f32 + gae_linear_fwd_f32).
 * Y may be NULL (encode / evaluation); training passes it, dW = dPre^T Y needs the aggregated rows. */
int gae_gcn_layer_fwd_f32(const int64_t *rowptr, const int32_t *col, const float *Hin, int64_t ldh,
                          const float *W, const float *b, float *Hout, int64_t ldo, float *Y,
                          int64_t ldy, int64_t n, int32_t d_in, int32_t d_out, int32_t act,
                          void *stream);
/* Adjoint of the layer in one call: dW, db, and -- unless dHin is NULL (first layer, gae.py:50) --
 * dHin = A^T (dPre W) over CSR(A^T) (plan_t / hub_ws_t as for gae_spmm_csr_f32, may be NULL).
 * dY: scratch rows [n, ld_dy]; ws: gae_gcn_layer_bwd_ws_bytes() bytes. */
int64_t gae_gcn_layer_bwd_ws_bytes(int64_t n, int32_t d_in, int32_t d_out);
int gae_gcn_layer_bwd_f32(const int64_t *rowptr_t, const int32_t *col_t, const gae_hub_plan_t *plan_t,
                          float *hub_ws_t, const float *Y, int64_t ldy, const float *W,
                          const float *Hout, int64_t ldo, const float *dHout, int64_t ld_dh, float *dY,
                          int64_t ld_dy, float *dHin, int64_t ld_dhin, float *dW, float *db, void *ws,
                          int64_t ws_bytes, int64_t n, int32_t d_in, int32_t d_out, int32_t act,
                          void *stream);

/* ---- K4: dropout (always on in the reference decoder) ------------------------------------ */
/* Replaces  z = F.dropout(z, self.dropout)   gae.py:70  (training=True regardless of mode).
 * mask_mode 0: draw keep-mask from Philox4x32-10(seed, offset) and WRITE it to mask (u8);
 * mask_mode 1: READ the given keep-mask (parity tests inject the oracle's mask).
 * Zd = Z * mask / (1-p).  Z/Zd are [n,d] with leading dimensions. */
int gae_dropout_fwd_f32(const float *Z, int64_t ldz, float *Zd, int64_t ldzd, uint8_t *mask,
                        int64_t n, int32_t d, float p, uint64_t seed, uint64_t offset,
                        int32_t mask_mode, void *stream);
/* Same with the Philox state {seed, offset} in DEVICE memory (uint64[2]); the offset is advanced
 * on the stream after the draw, so a captured CUDA graph draws a fresh mask on every replay. */
int gae_dropout_fwd_devrng_f32(const float *Z, int64_t ldz, float *Zd, int64_t ldzd,
                               uint8_t *mask, int64_t n, int32_t d, float p,
                               uint64_t *rng_state, void *stream);
/* dZ = dZd * mask / (1-p) * (*grad_scale or 1 if NULL)  */
int gae_dropout_bwd_f32(const float *dZd, int64_t ld_dzd, const uint8_t *mask, float *dZ,
                        int64_t ld_dz, int64_t n, int32_t d, float p, const float *grad_scale,
                        void *stream);

/* ---- K5 / K6: fused InnerProductDecoder + weighted BCE-with-logits ------------------------- */
/* Replaces  torch.mm(z, z.t())                               gae.py:71
 *           adj = g.adjacency_matrix().to_dense()            train_inductive.py:44
 *           BCELoss(adj_logits, adj, pos_weight=pos_weight)  train_inductive.py:48
 * without materialising any N x N array:
 *   L = (1/N^2) [ sum_ij softplus(x_ij) + sum_{e=(i,j)} (pw softplus(-x_ij) - softplus(x_ij)) ],
 *   x_ij = <Zd_i, Zd_j>, the edge sum running over the CSR with multiplicity.
 * mode bit0: compute loss -> *loss (device scalar)
 * mode bit1: compute dZd_unit = dL/dZd for grad_loss = 1 (needs rowptr_t/col_t = CSR(A^T))
 * d <= 128.  ws: gae_decoder_ws_bytes(n, d) bytes. */
#define GAE_DEC_LOSS 1
#define GAE_DEC_GRAD 2
int64_t gae_decoder_ws_bytes(int64_t n, int32_t d);
int gae_decoder_bce_f32(const float *Zd, int64_t ldz, int64_t n, int32_t d,
                        const int64_t *rowptr, const int32_t *col, const int64_t *rowptr_t,
                        const int32_t *col_t, float pos_weight, int32_t mode, float *loss,
                        float *dZd_unit, int64_t ld_dz, void *ws, int64_t ws_bytes,
                        void *stream);
/* Block-diagonal variant for batched graphs (NOT the reference default, which
 * decodes the full (sum n_k)^2 matrix including cross-molecule negatives, train_inductive.py:44-48):
 * the pair sum runs over j in [blk_lo[i], blk_hi[i]) only, normalised by n_pairs = sum_k n_k^2. */
int64_t gae_decoder_blockdiag_ws_bytes(int64_t n, int32_t d);
int gae_decoder_bce_blockdiag_f32(const float *Zd, int64_t ldz, int64_t n, int32_t d,
                                  const int64_t *rowptr, const int32_t *col,
                                  const int64_t *rowptr_t, const int32_t *col_t,
                                  const int64_t *blk_lo, const int64_t *blk_hi, double n_pairs,
                                  float pos_weight, int32_t mode, float *loss, float *dZd_unit,
                                  int64_t ld_dz, void *ws, int64_t ws_bytes, void *stream);
/* Materialised logits X = Zd Zd^T [n,n] (the value GAE.forward returns, gae.py:54-55). */
int gae_decoder_logits_f32(const float *Zd, int64_t ldz, int64_t n, int32_t d, float *X,
                           int64_t ldx, void *stream);

/* ---- whole train step behind one call --------------------------------------------------------- */
/* Encoder forward (gae.py:49-52), dropout + fused decoder loss (gae.py:69-72, train_inductive.py:
 * 44-48) and, with want_grad, the full backward (train_inductive.py:51) sequenced on `stream` out of
 * ONE caller-owned workspace: the host cost of a step is a single call (matters on molecule batches
 * where a step is ~25 kernels of a few microseconds).  The optimiser step stays with the caller.
 *   W, b, dW, db : HOST arrays [n_layers] of DEVICE pointers (W_l is [dims[l+1], dims[l]] row-major)
 *   mask_in      : optional injected keep-mask [n, dims[L]] u8; else rng_state (device u64[2]) is used
 *   blk_lo/hi    : per-node block ranges when desc->per_graph != 0
 *   Z_out        : embeddings [n, ldz] (written; also what gae.py:53 stores back into ndata['h'])
 *   dW, db are the gradients of the MEAN loss (grad_output = 1). */
#define GAE_MAX_LAYERS 8
typedef struct gae_step_desc_t {
    int32_t n_layers;
    int32_t dims[GAE_MAX_LAYERS + 1];   /* in_dim, hidden_dims...                       */
    int32_t acts[GAE_MAX_LAYERS];       /* GAE_ACT_* per layer (gae.py:36-45)            */
    float dropout_p;                    /* decoder dropout, gae.py:64                    */
    float pos_weight;                   /* train_inductive.py:46                         */
    int32_t per_graph;                  /* 0 = full N x N pairs (reference), 1 = block-diagonal */
    int32_t x_aggregated;               /* 1 = X already holds A X (the caller aggregated the input features
                                         * once: they and the graph are fixed across the epochs of
                                         * train_transductive.py:45-46,63); the first SpMM is skipped        */
} gae_step_desc_t;
int64_t gae_step_ws_bytes(const gae_step_desc_t *desc, int64_t n, const gae_hub_plan_t *plan,
                          const gae_hub_plan_t *plan_t);
int gae_step_fwd_bwd_f32(const gae_step_desc_t *desc, int64_t n, const int64_t *rowptr,
                         const int32_t *col, const gae_hub_plan_t *plan, const int64_t *rowptr_t,
                         const int32_t *col_t, const gae_hub_plan_t *plan_t, const float *X,
                         int64_t ldx, const float *const *W, const float *const *b,
                         const uint8_t *mask_in, uint64_t *rng_state, const int64_t *blk_lo,
                         const int64_t *blk_hi, double n_pairs, int32_t want_grad, float *loss,
                         float *Z_out, int64_t ldz, float *const *dW, float *const *db, void *ws,
                         int64_t ws_bytes, void *stream);

/* ---- graph indexing on device ("bit-exact adjacency/degree indexing") ---------------------- */
/* in_degrees (train_transductive.py:55): deg[v] = rowptr[v+1]-rowptr[v] as int64. */
int gae_in_degrees_i64(const int64_t *rowptr, int64_t n_rows, int64_t *deg, void *stream);
/* dgl.batch (train_inductive.py:34): block-diagonal union of K CSR graphs already
 * concatenated back to back: col_cat [E_total] holds each graph's LOCAL source ids,
 * edge_graph_ptr [K+1] the edge offsets, node_off [K+1] the node prefix sums.  Adds the node
 * offset of its graph to every col entry (in place). */
int gae_batch_offset_cols_i32(int32_t *col_cat, const int64_t *edge_graph_ptr,
                              const int64_t *node_off, int64_t n_graphs, int64_t n_edges,
                              void *stream);

/* dgl.batch from a PACKED dataset (8f rank 1): every member graph is resident in HBM inside one
 * CSR (rowptr_all int64 over all nodes with global edge offsets, col_all int32 with graph-LOCAL
 * source ids, node_ptr [G+1] node prefix sums, feat_all [sum n, ldf]).  Builds the union of graphs
 * gid[0..n_graphs): out_rowptr [n_out+1], out_col [e_out] (+ node offsets), out_feat [n_out, ld_out]
 * (feat_all / out_feat may be NULL).  node_off / edge_off [n_graphs+1] are the prefix sums of the
 * selected graphs' node / edge counts (device). */
int gae_batch_assemble(const int64_t *rowptr_all, const int32_t *col_all, const int64_t *node_ptr,
                       const int64_t *gid, int64_t n_graphs, const int64_t *node_off,
                       const int64_t *edge_off, int64_t n_out, int64_t e_out,
                       int64_t *out_rowptr, int32_t *out_col, const float *feat_all, int64_t ldf,
                       int32_t d, float *out_feat, int64_t ld_out, void *stream);

/* Diagnostic of the tcgen05 / TMEM dense pass (csrc/decoder_tc.cu, the default for d <= 16 and n >= 512): runs
 * ONE 128 x 128 tile (tile_i <= tile_j) through the kernel's own code path and returns what the tensor cores
 * produced -- S = Z_I Z_J^T [128,128], G_i = sigmoid(S) Z_J [128,16], G_j = sigmoid(S)^T Z_I [128,16] (left
 * untouched on a diagonal tile) -- so that each shared-memory / instruction descriptor can be checked against
 * a host product on its own.  Synchronises the stream; *timeouts = bounded waits that expired (0 when sound). */
int gae_decoder_tile_probe_f32(const float *Zd, int64_t ldz, int64_t n, int32_t d, int32_t tile_i,
                               int32_t tile_j, float *S, float *G_i, float *G_j, int32_t *timeouts,
                               void *stream);

/* ---- K7: halo exchange helpers (8e) -------------------------------------------------------- */
/* Pack rows idx[0..m) of X into out (send buffer of the all-to-all-v). */
int gae_gather_rows_f32(const float *X, int64_t ldx, const int64_t *idx, int64_t m, int32_t d,
                        float *out, int64_t ld_out, void *stream);
/* One-sided halo pull over NVLink peer memory: for i in [0,m): out[i,:] = peer_base[owner
 * slot][idx[i],:], where peer_ptrs is a device array of P mapped base pointers (own rank's
 * slot may be its local pointer) and owner[i] selects the slot. */
int gae_pull_rows_p2p_f32(const float *const *peer_ptrs, const int32_t *owner,
                          const int64_t *idx, int64_t m, int64_t ldx, int32_t d, float *out,
                          int64_t ld_out, void *stream);
/* One-sided halo PUSH (the default exchange): for i in [0,m): peer_ptrs[dst_peer[i]][dst_row[i],:]
 * = X[send_idx[i],:].  The owner reads its own rows and stores them into the peers' halo regions
 * over NVLink (posted writes); fuses the pack step of an all-to-all into the transfer. */
int gae_push_rows_p2p_f32(const float *X, int64_t ldx, const int64_t *send_idx,
                          const int32_t *dst_peer, const int64_t *dst_row,
                          float *const *peer_ptrs, int64_t m, int64_t ld_peer, int32_t d,
                          void *stream);
/* CUDA IPC plumbing for the pull / push paths: 64-byte handles. */
int gae_ipc_get_handle(const void *dev_ptr, uint8_t handle_out[64], int64_t *offset_out);
int gae_ipc_open_handle(const uint8_t handle[64], void **dev_ptr_out);
int gae_ipc_close_handle(void *dev_ptr);

/* ---- K7b: staged halo exchange fused with the row-block SpMMs (8e; the default multi-GPU path) ---- */
/* Device-side protocol, no collective and no host sync on the data path.  Every rank maps its peers'
 * [local | halo] feature buffers and one FLAG block (GAE_HALO_FLAG_WORDS uint64, zero-initialised once)
 * per partitioned operator through CUDA IPC (gae_ipc_*).  Halo rows are tagged with a "stage" by their
 * consumer: either the first ROW BLOCK that reads them, or a POPULARITY CLASS (stage 0 = the few remote
 * sources most of the edges point at; the consumer then sums class by class, blocks[s].accumulate = 1
 * for s > 0).  Per SpMM and rank:
 *   gae_halo_push_f32     one persistent kernel: waits until every peer has consumed epoch-1, then pushes
 *                         the rows of stage 0, 1, ... into the peers' halo regions (posted NVLink stores)
 *                         and publishes landed[rank][stage] = epoch to every peer after each stage;
 *   gae_halo_wait_f32     one-warp kernel: acquires landed[q][stage] >= epoch from all peers q;
 *   gae_halo_release_f32  publishes consumed[rank] = epoch to every peer (my halo may be overwritten);
 *   gae_halo_spmm_f32     the whole operator Y = (A X)[my rows]: push on comm_stream, and for
 *                         s = 0..n_stages-1 { wait(s); SpMM of row block s } + release, so the transfer
 *                         of later stages overlaps the aggregation of earlier ones.  Row blocks alternate
 *                         between compute_stream and aux_stream (NULL = all on compute_stream; consecutive
 *                         blocks then need no separate partial_ws) so that the tail of one block overlaps
 *                         the head of the next; everything is joined back into compute_stream.
 * `epoch` counts the calls on one operator from 1.  Flag waits are bounded by timeout_ms (default 10 s):
 * on expiry an error word is set and the kernel falls through -- gae_halo_status() reports it -- so a
 * protocol fault gives a wrong, reported result and never a hung GPU.  Replaces nothing in the reference
 * (single device, train_inductive.py:26,29); it is how update_all (gae.py:28) spans the GPUs of a box. */
#define GAE_HALO_MAX_WORLD 16
#define GAE_HALO_MAX_STAGES 32
#define GAE_HALO_FLAG_WORDS 1024
typedef struct gae_halo_exchange_t {
    int32_t world, rank, n_stages, d;
    int64_t ld;                   /* row stride (floats) of EVERY rank's [local | halo] buffer          */
    const float *x_local;         /* my [local | halo] buffer (device)                                  */
    float *const *peer_x;         /* DEVICE array [world]: that buffer of every rank, IPC-mapped        */
    uint64_t *const *peer_flags;  /* DEVICE array [world]: this operator's flag block on every rank     */
    uint64_t *flags;              /* my flag block (device)                                             */
    const int64_t *send_src;      /* DEVICE [m]: local row of each entry, sorted by stage               */
    const int32_t *send_peer;     /* DEVICE [m]: destination rank                                       */
    const int64_t *send_dst;      /* DEVICE [m]: row in the destination's buffer                        */
    const int64_t *stage_ptr;     /* HOST [n_stages+1]: entry range of each stage                       */
    uint32_t *stage_done;         /* DEVICE [n_stages]: zero-initialised arrival counters (scratch)     */
    int32_t push_ctas;            /* 0 = default (64)                                                   */
    int32_t push_threads;         /* 0 = default (256)                                                  */
    int32_t timeout_ms;           /* 0 = default (10000)                                                */
    /* Optional sender-side FOLDING (pre_n_rows > 0).  A consumer may ask an owner not for source rows but
     * for their SUM over the edges into one of its destination rows: on skewed graphs the rarely used remote
     * sources mostly feed hub destinations, so one folded row replaces many copied ones (R-MAT, 8 ranks: the
     * exchange shrinks to ~0.56 of the deduplicated halo).  The owner computes the folded rows with the
     * ordinary SpMM -- CSR (pre_rowptr, pre_col) over its LOCAL rows -- into staging rows
     * [pre_row0, pre_row0 + pre_n_rows) of its own buffer, behind the halo region; send entries of stages
     * >= pre_stage may name staging rows as their source.  gae_halo_spmm_f32 runs that SpMM on aux_stream
     * while stages < pre_stage are already being pushed. */
    const int64_t *pre_rowptr;    /* DEVICE [pre_n_rows+1]                                              */
    const int32_t *pre_col;       /* DEVICE: local source rows                                          */
    const gae_hub_plan_t *pre_plan;
    float *pre_ws;                /* segment workspace of pre_plan                                      */
    int64_t pre_n_rows;           /* 0 = no folding                                                     */
    int64_t pre_row0;
    int32_t pre_stage;
} gae_halo_exchange_t;
typedef struct gae_halo_block_t {
    int64_t row0, n_rows;         /* local row range of the block                                       */
    const int64_t *rowptr;        /* DEVICE [n_rows+1], rebased to 0                                    */
    const int32_t *col;           /* DEVICE: the block's slice of the [local | halo] column array       */
    const gae_hub_plan_t *plan;   /* hub plan of the block (may be NULL)                                */
    float *partial_ws;            /* its segment workspace                                              */
    int32_t accumulate;           /* 0: Y[rows] = A_s X ; 1: Y[rows] += A_s X (a later PASS over the same rows:  */
                                  /* the stages are then column classes of the halo instead of row blocks)      */
} gae_halo_block_t;
int gae_halo_push_f32(const gae_halo_exchange_t *ex, uint64_t epoch, void *stream);
/* Stages [stage0, stage1) only (the folded rows of later stages may not exist yet). */
int gae_halo_push_range_f32(const gae_halo_exchange_t *ex, uint64_t epoch, int32_t stage0, int32_t stage1,
                            void *stream);
int gae_halo_wait_f32(const gae_halo_exchange_t *ex, int32_t stage, uint64_t epoch, void *stream);
int gae_halo_release_f32(const gae_halo_exchange_t *ex, uint64_t epoch, void *stream);
int gae_halo_spmm_f32(const gae_halo_exchange_t *ex, const gae_halo_block_t *blocks /* HOST [n_stages] */,
                      float *Y, int64_t ldy, uint64_t epoch, void *compute_stream, void *comm_stream,
                      void *aux_stream);
/* Synchronous: copies the error word back; GAE_ERR_TIMEOUT if any flag wait expired since start. */
int gae_halo_status(const gae_halo_exchange_t *ex, int64_t *timeouts);

/* Halo planning on HOST arrays, O(E + V/64).  src_global[E]: global source id of every edge whose row
 * this rank owns; bounds[world+1]: contiguous vertex blocks.  Two-call protocol:
 *   count -> n_halo, recv_counts[world] (halo rows owned by each peer)
 *   fill  -> halo_ids[n_halo] (ascending = grouped by owner), col_local[E] = column in [local | halo] */
int gae_halo_plan_count_host(const int64_t *src_global, int64_t n_edges, const int64_t *bounds,
                             int32_t world, int32_t rank, int64_t *n_halo, int64_t *recv_counts);
int gae_halo_plan_fill_host(const int64_t *src_global, int64_t n_edges, const int64_t *bounds,
                            int32_t world, int32_t rank, int64_t *halo_ids, int32_t *col_local);
/* halo_stage[h] = first row block (row_bounds[n_stages+1] over the local rows) that reads halo row h. */
int gae_halo_stage_tags_host(const int64_t *rowptr, const int32_t *col_local, int64_t n_local,
                             int64_t n_halo, const int64_t *row_bounds, int32_t n_stages,
                             int32_t *halo_stage);
/* Send lists of the push kernel.  send_idx: my local rows requested by the peers, grouped by peer in
 * request order (send_counts[world]); send_stage: the stage of each entry at its consumer (NULL = 0);
 * destination of entry j of peer q: send_dst[j] when given (the consumer chose its halo layout), else
 * dst_base[q] + (position in q's request list).  Output: entries sorted by stage and, within a stage,
 * interleaved over the peers; stage_ptr[n_stages+1]. */
int gae_halo_push_lists_host(const int64_t *send_idx, const int32_t *send_stage, const int64_t *send_dst,
                             const int64_t *send_counts, const int64_t *dst_base, int32_t world,
                             int32_t n_stages, int64_t *out_src, int32_t *out_peer, int64_t *out_dst,
                             int64_t *stage_ptr);

/* ---- optimiser step (train_inductive.py:40,52: torch.optim.Adam, no weight decay / amsgrad) ---- */
/* One launch over up to GAE_ADAM_MAX_TENSORS parameter tensors (device pointers passed in HOST arrays):
 *   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= lr/(1-b1^step) * m / (sqrt(v)/sqrt(1-b2^step) + eps)
 * `step` counts from 1.  exp_avg / exp_avg_sq are the state tensors torch.optim.Adam keeps under the same
 * names, so optimiser checkpoints stay interchangeable. */
#define GAE_ADAM_MAX_TENSORS 16
int gae_adam_step_f32(int32_t n_tensors, float *const *params, const float *const *grads,
                      float *const *exp_avg, float *const *exp_avg_sq, const int64_t *numel, float lr,
                      float beta1, float beta2, float eps, int64_t step, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GAE_B200_H_ */
