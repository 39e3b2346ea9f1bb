/* CPU restatement of the reference's message-passing SpMM.  TEST INFRASTRUCTURE ONLY
 * (checker for tests/, smoke() and bench.py's cpu_baseline / --impl reference legs).
 *
 * Follows /root/reference/gae_dgl/gae.py:18-19,28:
 *     gcn_msg = fn.copy_src('h','m'); gcn_reduce = fn.sum('m','h'); g.update_all(...)
 * i.e. Y[v,:] = sum over in-edges (u->v) of X[u,:], multigraph, zero rows for
 * in-degree 0.  DGL's CPU kernel for this is a row-parallel OpenMP loop over the
 * in-edge CSR; this file restates that shape (one thread per block of dst rows, edges in
 * CSR order, fp32 accumulate).  PARITY UNPINNED: DGL itself is not installed here.
 */
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int oracle_spmm_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* fp32 accumulate, CSR order */
void oracle_spmm_csr_f32(const int64_t *rowptr, const int32_t *col, const float *vals,
                         const float *X, int64_t ldx, float *Y, int64_t ldy,
                         int64_t n_rows, int32_t d) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t v = 0; v < n_rows; ++v) {
        float *y = Y + v * ldy;
        memset(y, 0, sizeof(float) * (size_t)d);
        for (int64_t e = rowptr[v]; e < rowptr[v + 1]; ++e) {
            const float *x = X + (int64_t)col[e] * ldx;
            if (vals) {
                const float w = vals[e];
                for (int32_t k = 0; k < d; ++k) y[k] += w * x[k];
            } else {
                for (int32_t k = 0; k < d; ++k) y[k] += x[k];
            }
        }
    }
}

/* fp64 accumulate of fp32 inputs: ground truth for tolerance checks */
void oracle_spmm_csr_f64acc(const int64_t *rowptr, const int32_t *col, const float *X,
                            int64_t ldx, double *Y, int64_t ldy, int64_t n_rows, int32_t d) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t v = 0; v < n_rows; ++v) {
        double *y = Y + v * ldy;
        for (int32_t k = 0; k < d; ++k) y[k] = 0.0;
        for (int64_t e = rowptr[v]; e < rowptr[v + 1]; ++e) {
            const float *x = X + (int64_t)col[e] * ldx;
            for (int32_t k = 0; k < d; ++k) y[k] += (double)x[k];
        }
    }
}
