// triple.cuh — the per-triple stage of IDDGCN_Layer.call and the decoder/loss/optimizer kernels.
#pragma once
#include "graph.cuh"

namespace iddgcn {

constexpr int RMAX = 8;  // relations handled in registers by the per-triple kernels

// HT[half][k][:] = emb[idx_half[k]][:]   (Embedding lookups, prediction/IDDGCN.py:226-227)
int gather_embeddings(const iddgcn_plan* p, const float* emb, int64_t D, float* HT, cudaStream_t st);

// Epilogue of the layer forward (IDDGCN.py:66-79).  P holds [Hin;Tin]·W_self on entry ((2,T,D)) and the
// layer outputs sigmoid(P + Σ_i gate_i Z_i[idx]) on exit; alpha [T,R] is saved when non-NULL.
int layer_combine(const iddgcn_plan* p, const float* Hin, const float* Z, const float* W_alpha, const float* b_alpha,
                  int64_t D, float* P_inout, float* alpha_out, cudaStream_t st);

// dHT (2,T,D) holds dL/d(outputs) on entry and dL/d(pre-activations) on exit; dz [T,R] = softmax-logit grads.
int layer_bwd_pre(const iddgcn_plan* p, const float* HT_out, const float* alpha, const float* Z, int64_t D,
                  float* dHT_inout, float* dz, cudaStream_t st);

// dHin[k][d] += Σ_i dz[k][i] · W_alpha[d][i]     (head half only)
int layer_bwd_post(const float* dz, const float* W_alpha, int64_t T, int64_t D, int R, float* dHin, cudaStream_t st);

// out[i][n][:] (+)= Σ_{k: head_k=n} w_i(k) dP[0][k][:] + Σ_{k: tail_k=n} w_i(k) dP[1][k][:]
//   alpha != NULL: w_i(k) = sigmoid(alpha[k][i]), i < R   (dZ_i, SURVEY App. A eq. 10)
//   alpha == NULL: single output plane, w = 1            (embedding-lookup gradient, eq. 12)
int segment_scatter(const iddgcn_plan* p, const float* dP, const float* alpha, int64_t D, float* out, int accumulate,
                    cudaStream_t st);

// out[c] (+)= Σ_rows in[row][c]; cols <= RMAX. scratch >= colsum_scratch_elems(rows, cols)
size_t colsum_scratch_elems(int64_t rows, int cols);
int colsum(const float* in, int64_t rows, int cols, float* out, int accumulate, float* scratch, cudaStream_t st);

// DistMult (IDDGCN.py:103-109): y[k] = sigmoid(Σ_d H[k,d]·Rel[r_k,d]·T[k,d]);  HT is (2,T,D)
int distmult_fwd(const iddgcn_plan* p, const float* HT, const float* rel_emb, int64_t D, float* y, cudaStream_t st);
// dHT (2,T,D) = dq·Rel[r]⊙(T|H), d_rel_emb[r] += Σ_{k: r_k=r} dq_k H_k⊙T_k with dq = dy·y(1-y)
size_t distmult_scratch_elems(const iddgcn_plan* p, int64_t D);
int distmult_bwd(const iddgcn_plan* p, const float* HT, const float* rel_emb, const float* y, const float* dy,
                 int64_t D, float* dHT, float* d_rel_emb, float* scratch, cudaStream_t st);

// Keras BinaryCrossentropy on y = [pos | neg] with labels [1 | 0] (IDDGCN.py:161-168):
// loss_out[0] = mean BCE (before the 1/N scaling), dy = d(mean BCE)/dy · grad_scale.
size_t bce_scratch_elems(int64_t T_total);
int bce_loss_grad(const float* y, int64_t T_pos, int64_t T_total, float grad_scale, float* dy, float* loss_out,
                  float* scratch, cudaStream_t st);

// Keras Adam; step read from a device counter when step_dev != NULL (CUDA-graph friendly)
int adam_apply(float* theta, const float* grad, float* m, float* v, int64_t n, float lr, float b1, float b2,
               float eps, int64_t step, const int64_t* step_dev, cudaStream_t st);
int increment_counter(int64_t* counter, cudaStream_t st);

}  // namespace iddgcn
