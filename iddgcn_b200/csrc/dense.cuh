// dense.cuh — small fp32 SIMT GEMMs used by the triple stage (feature width D is 64..128 in every
// configuration of the reference, so these are not tensor-core shaped: K or N is tiny and 1e-5 parity
// rules out TF32/bf16 MMA; SURVEY §7 "fp32 tolerance vs tensor cores").
#pragma once
#include "common.cuh"

namespace iddgcn {

// C[b] (+)= op(A[b]) · op(B[b]),  b < batch.  Row-major operands with leading dimensions.
//   transA == false: A is [M,K] (lda >= K);  true: A is [K,M] (lda >= M)
//   transB == false: B is [K,N] (ldb >= N);  true: B is [N,K] (ldb >= K)
// Deterministic: one thread block owns each C tile and sums K in ascending order.
int gemm_f32(bool transA, bool transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, int64_t strideA,
             const float* B, int64_t ldb, int64_t strideB, float* C, int64_t ldc, int64_t strideC, int batch,
             bool accumulate, cudaStream_t st);

// the register-tiled fp32 SIMT kernels, never the tensor-core path
int gemm_f32_simt(bool transA, bool transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda,
                  int64_t strideA, const float* B, int64_t ldb, int64_t strideB, float* C, int64_t ldc, int64_t strideC,
                  int batch, bool accumulate, cudaStream_t st);

// C[b] (+)= A[b]ᵀ · B[b] with a LONG reduction dimension K (A: [K,M], B: [K,N]; M,N small): split-K into
// a workspace followed by a fixed-order reduction — deterministic, no atomics.
// workspace must hold gemm_tn_workspace_elems(...) floats.
size_t gemm_tn_workspace_elems(int64_t M, int64_t N, int64_t K, int batch);
int gemm_tn_splitk_f32(int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, int64_t strideA, const float* B,
                       int64_t ldb, int64_t strideB, float* C, int64_t ldc, int64_t strideC, int batch,
                       bool accumulate, float* workspace, cudaStream_t st);

// ---- tensor-core path (tc_gemm.cu): 3xTF32 on tcgen05 for long-M, weight-sized N = K in {64, 128} ----------------------
// gemm_f32 takes it by itself when the call is eligible and IDDGCN_GEMM_TF32X3 is not "0".
bool tc_gemm_enabled();
bool tc_gemm_eligible(bool transA, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
                      int64_t ldb, const float* C, int64_t ldc, int64_t min_rows = 32768);
int tc_gemm_prepare(cudaStream_t st);   // allocation + attributes; call before capturing a stream that will use it
int gemm_tf32x3(bool transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, int64_t strideA,
                const float* B, int64_t strideB, float* C, int64_t ldc, int64_t strideC, int batch, bool accumulate,
                cudaStream_t st);

// weight-gradient product on the tensor cores: partial sums over tc_gemm_tn_splits(K) row ranges into ws[s][M][N]
int tc_gemm_tn_splits(int64_t rows);
bool tc_gemm_tn_eligible(int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B, int64_t ldb,
                         int64_t min_rows = 32768);
int gemm_tn_tf32x3_partials(int64_t M, int64_t K, const float* A, int64_t lda, const float* B, int64_t ldb, float* ws,
                            cudaStream_t st);

// the two implementations behind gemm_tn_splitk_f32 (same workspace contract)
int gemm_tn_splitk_simt(int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, int64_t strideA, const float* B,
                        int64_t ldb, int64_t strideB, float* C, int64_t ldc, int64_t strideC, int batch,
                        bool accumulate, float* workspace, cudaStream_t st);
int gemm_tn_tc(int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, int64_t strideA, const float* B,
               int64_t ldb, int64_t strideB, float* C, int64_t ldc, int64_t strideC, int batch, bool accumulate,
               float* workspace, cudaStream_t st);

// out[b] = in[b]ᵀ for small (weight-sized) row-major matrices [rows, cols]
int transpose_small(const float* in, float* out, int rows, int cols, int batch, cudaStream_t st);

}  // namespace iddgcn
