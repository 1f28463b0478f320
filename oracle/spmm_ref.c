/*
 * spmm_ref.c — CPU ORACLE / BASELINE (test infrastructure, never shipped in the product path).
 *
 * C restatement of what TensorFlow 2.7's CPU kernel does for
 *   tf.sparse.sparse_dense_matmul(adj_i, embeddings)           (prediction/IDDGCN.py:70)
 * and of its gradient w.r.t. the dense operand (adjoint SpMM, tf sparse_grad).  TensorFlow itself is a
 * third-party dependency of the reference that is absent from /root/reference and not installable here
 * (README.md:17-19 pins only `tensorflow>=2.7.0`); its published algorithm for SparseTensorDenseMatMul
 * on CPU is a sequential loop over the stored non-zeros in row-major order,
 *       out[row, :] += value * b[col, :]
 * which is what spmm_coo_seq() below does (threads = 1, exactly TF's order).  The *_omp variants give
 * the same per-row summation order but spread destination rows over all host threads — the
 * "with all the host threads it can use" arm of bench.py --impl reference.
 *
 * Only tests/ and bench.py's cpu_baseline / --impl reference legs may load this library.
 */
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int spmm_ref_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* TF-CPU order: one thread, nnz in stored order. out must be zeroed by the caller. */
void spmm_coo_seq(const int32_t* row, const int32_t* col, const float* val /* may be NULL = 1.0 */, int64_t nnz,
                  const float* b, int64_t D, float* out) {
  for (int64_t e = 0; e < nnz; ++e) {
    const float v = val ? val[e] : 1.0f;
    const float* src = b + (int64_t)col[e] * D;
    float* dst = out + (int64_t)row[e] * D;
    for (int64_t d = 0; d < D; ++d) dst[d] += v * src[d];
  }
}

/* out[s, :] = sum_{e in ptr[s]..ptr[s+1]} val[e] * b[idx[e], :]   for s in [0, num_segs).
 * Forward: ptr = csr_rowptr over R*N segments, idx = csr_col, b = X [N,D], out = S [R*N, D]. */
void spmm_csr_omp(const int32_t* ptr, const int32_t* idx, const float* val, int64_t num_segs, const float* b,
                  int64_t D, float* out, int threads) {
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 1024)
  for (int64_t s = 0; s < num_segs; ++s) {
    float* dst = out + s * D;
    memset(dst, 0, (size_t)D * sizeof(float));
    for (int32_t e = ptr[s]; e < ptr[s + 1]; ++e) {
      const float v = val ? val[e] : 1.0f;
      const float* src = b + (int64_t)idx[e] * D;
      for (int64_t d = 0; d < D; ++d) dst[d] += v * src[d];
    }
  }
}

/* Adjoint with relations fused: dX[k, :] = sum_r sum_{e in cscptr[r*N+k]..} val[perm[e]] * dS[r*n_rows + row[e], :] */
void spmm_csc_bwd_omp(const int32_t* cscptr, const int32_t* cscrow, const int32_t* perm, const float* val, int R,
                      int64_t N, int64_t n_rows, const float* dS, int64_t D, float* dX, int threads) {
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 1024)
  for (int64_t k = 0; k < N; ++k) {
    float* dst = dX + k * D;
    memset(dst, 0, (size_t)D * sizeof(float));
    for (int r = 0; r < R; ++r) {
      const int64_t s = (int64_t)r * N + k;
      const float* base = dS + (int64_t)r * n_rows * D;
      for (int32_t e = cscptr[s]; e < cscptr[s + 1]; ++e) {
        const float v = val ? val[perm[e]] : 1.0f;
        const float* src = base + (int64_t)cscrow[e] * D;
        for (int64_t d = 0; d < D; ++d) dst[d] += v * src[d];
      }
    }
  }
}

/* Same product with a double accumulator per output element, rounded to float once at the end: the accuracy reference
 * for rows of very high degree (a sequential fp32 sum of 10^6 terms is itself ~1e-5 off), used next to spmm_csr_omp by
 * bench.py's parity check.  D <= 1024. */
void spmm_csr_omp_acc64(const int32_t* ptr, const int32_t* idx, const float* val, int64_t num_segs, const float* b,
                        int64_t D, float* out, int threads) {
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 1024)
  for (int64_t s = 0; s < num_segs; ++s) {
    double acc[1024];
    for (int64_t d = 0; d < D; ++d) acc[d] = 0.0;
    for (int32_t e = ptr[s]; e < ptr[s + 1]; ++e) {
      const double v = val ? (double)val[e] : 1.0;
      const float* src = b + (int64_t)idx[e] * D;
      for (int64_t d = 0; d < D; ++d) acc[d] += v * (double)src[d];
    }
    float* dst = out + s * D;
    for (int64_t d = 0; d < D; ++d) dst[d] = (float)acc[d];
  }
}
