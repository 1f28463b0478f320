/*
 * iddgcn.h — C ABI of libiddgcn_b200.so (sm_100a).
 *
 * Drop-in boundary for the relational directed graph-convolution hot path of
 * AhauBioinformatics/IDDGCN.  The reference has no FFI of its own: the seam is
 * its Keras-layer Python API, which executes stock TensorFlow ops.  Each entry
 * point below names the reference call site (file:line under the reference
 * tree) whose arithmetic it replaces.  The Python host layer
 * (iddgcn_b200 package) binds these through ctypes and re-exposes the reference's
 * own names (get_adj_mats, IDDGCN_Layer, DistMult, get_IDDGCN_Model).
 *
 * Conventions
 *   - every function returns 0 on success, a negative IDDGCN_E* code on error;
 *     iddgcn_last_error() returns a thread-local message for the last failure.
 *   - all tensor arguments are plain pointers + sizes; unless a name ends in
 *     `_host`, pointers are DEVICE pointers on the handle's device, row-major,
 *     contiguous, caller-owned.  The library never returns memory the caller
 *     must free except through the opaque handles (create/destroy pairs).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *     Calls are asynchronous on that stream unless stated otherwise.
 *   - handles are bound to the device that was current at creation and are not
 *     thread-safe; the library is re-entrant across handles.
 *   - there is NO CPU fallback: with no CUDA device every compute entry point
 *     returns IDDGCN_ECUDA.
 *   - integer outputs (CSR/CSC) are bit-exact w.r.t. the reference semantics;
 *     fp32 kernels accumulate in fp32 with a fixed, atomics-free order, so
 *     results are run-to-run reproducible.
 */
#ifndef IDDGCN_H_
#define IDDGCN_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IDDGCN_OK        0
#define IDDGCN_EINVAL   -1   /* bad argument (shape, range, NULL)            */
#define IDDGCN_ECUDA    -2   /* CUDA runtime error / no device               */
#define IDDGCN_ENOMEM   -3   /* device or host allocation failed             */
#define IDDGCN_ERANGE   -4   /* entity index outside [0, num_entities)       */
#define IDDGCN_ELIMIT   -5   /* size exceeds an implementation limit (int32) */

/* edge-value modes shared by the SpMM family */
#define IDDGCN_VALS_ONES     0   /* a_e = 1.0 (utils1.py:439)                              */
#define IDDGCN_VALS_EXPLICIT 1   /* a_e = vals[e]              (explaiNE.py:86 watches adj.values) */
#define IDDGCN_VALS_SIGMOID  2   /* a_e = sigmoid(vals[e])     (GnnExplainer.py:33, IDDGCN_explain.py:52) */

typedef struct iddgcn_graph iddgcn_graph_t;
typedef struct iddgcn_model iddgcn_model_t;
typedef struct iddgcn_plan  iddgcn_plan_t;

const char* iddgcn_last_error(void);
/* library / build identification: "iddgcn_b200 <ver> sm_100a" */
const char* iddgcn_version(void);
/* number of kernel launches issued by this library on the calling thread's
 * handles since process start (bench.py's gpu_launches). */
int64_t iddgcn_launch_count(void);
/* Graph / plan handles and build temporaries come from a small caching pool (cudaFree costs tens of ms in a
 * context that maps several GB); this returns every cached block to the driver. */
int iddgcn_pool_trim(void);

/* ------------------------------------------------------------------------
 * Graph construction — replaces utils1.get_adj_mats + distinct
 * (prediction/utils1.py:415-451): per relation i keep triples with rel==i,
 * take (head,tail), np.unique(axis=0) (lexicographic sort + dedup), values 1.
 * Built entirely on device: pack (rel,head,tail) keys -> LSD radix sort ->
 * unique -> relation-segmented CSR; then (rel,tail,head) sort -> CSC plus the
 * CSC->CSR edge permutation.
 *
 *   triples      int64 [M,3] rows (head, rel, tail)  (IDDGCN.py:368 passes X_train.values)
 *   triples_on_device  0: host pointer, 1: device pointer
 *   num_entities N  = number of source nodes (columns)
 *   row_lo,row_hi   destination-row block kept by this handle: triples whose
 *                   head is outside [row_lo,row_hi) are dropped (row-block
 *                   partition for multi-GPU; pass 0,N for the whole graph)
 *   Triples whose rel is outside [0,R) are ignored (the reference's boolean
 *   filter utils1.py:425 never selects them); head/tail outside [0,N) ->
 *   IDDGCN_ERANGE.  Synchronises the stream before returning (nnz is needed
 *   on the host).
 * Layout produced (all int32, device):
 *   csr_rowptr [R*n_rows+1]  segment s = r*n_rows + (head-row_lo)
 *   csr_col    [nnz]         tail of each edge, ascending within a segment
 *   csr_row    [nnz]         local head (head-row_lo) of each edge
 *   csc_rowptr [R*N+1]       segment s = r*N + tail
 *   csc_row    [nnz]         local head of each in-edge, ascending within a segment
 *   csc_perm   [nnz]         CSR edge id of each CSC entry
 * ------------------------------------------------------------------------ */
int iddgcn_graph_build(const int64_t* triples, int64_t M, int triples_on_device,
                       int64_t num_entities, int num_relations,
                       int64_t row_lo, int64_t row_hi,
                       void* stream, iddgcn_graph_t** out);
int iddgcn_graph_destroy(iddgcn_graph_t* g);
/* nnz_per_rel may be NULL, else int64[R] */
int iddgcn_graph_info(const iddgcn_graph_t* g, int64_t* num_entities, int* num_relations,
                      int64_t* n_rows, int64_t* row_lo, int64_t* nnz, int64_t* nnz_per_rel);
/* copy the integer structure to HOST buffers (any pointer may be NULL). Blocking. */
int iddgcn_graph_export_host(const iddgcn_graph_t* g,
                             int32_t* csr_rowptr, int32_t* csr_col, int32_t* csr_row,
                             int32_t* csc_rowptr, int32_t* csc_row, int32_t* csc_perm);
/* borrow the device arrays (valid until destroy) — for zero-copy wrapping as tensors */
int iddgcn_graph_device_ptrs(const iddgcn_graph_t* g,
                             const int32_t** csr_rowptr, const int32_t** csr_col, const int32_t** csr_row,
                             const int32_t** csc_rowptr, const int32_t** csc_row, const int32_t** csc_perm);

/* ------------------------------------------------------------------------
 * Relational SpMM family — replaces tf.sparse.sparse_dense_matmul(adj_i, embeddings)
 * (prediction/IDDGCN.py:70) for all relations in ONE launch, and its TF
 * gradient (sparse_grad._SparseTensorDenseMatMulGrad; used through
 * IDDGCN.py:172, explaiNE.py:17, GnnExplainer.py:51, IDDGCN_explain.py:84).
 *
 *   X   [N, D]           source features           (fp32, D % 4 == 0, 16-byte aligned)
 *   S   [R, n_rows, D]   S_r = A_r · X
 *   dS  [R, n_rows, D]
 *   dX  [N, D]           dX = Σ_r A_rᵀ · dS_r      (accumulate!=0: dX += ...)
 *   vals [nnz]           per-edge values / mask logits in CSR edge order, or NULL
 *   dvals [nnz]          ∂L/∂vals[e]: mode EXPLICIT/ONES: <dS[row_e], X[col_e]>;
 *                        mode SIGMOID: that · σ(v)(1-σ(v))   (SparseDenseCwiseMul grad)
 * Per-edge weights are applied in registers and never materialised in HBM.
 * Sum order inside a row is ascending column (TF-CPU order)
 * per lane; no atomics anywhere.
 * ------------------------------------------------------------------------ */
int iddgcn_spmm_fwd(const iddgcn_graph_t* g, const float* X, int64_t D,
                    const float* vals, int vals_mode, float* S, void* stream);
int iddgcn_spmm_bwd_x(const iddgcn_graph_t* g, const float* dS, int64_t D,
                      const float* vals, int vals_mode, float* dX, int accumulate, void* stream);
int iddgcn_spmm_bwd_vals(const iddgcn_graph_t* g, const float* dS, const float* X, int64_t D,
                         const float* vals, int vals_mode, float* dvals, void* stream);
/* Σ_e a_e per relation -> out[R] (structure_loss counts, IDDGCN_explain.py:31-34) */
int iddgcn_edge_value_sums(const iddgcn_graph_t* g, const float* vals, int vals_mode,
                           float* out_per_rel, void* stream);

/* Row-range flavours (same kernels, same reference lines: IDDGCN.py:70 and its TF gradient).
 * A caller that pipelines a step (chunked host copies, or the grid-partitioned multi-GPU step of dist.py) computes
 * only part of the output rows per call:
 *   iddgcn_spmm_fwd_rows   : S rows [row_begin,row_end) of the flattened [R*n_rows, D] output; S_rows points at the
 *                            FIRST row of the range.
 *   iddgcn_spmm_bwd_x_rows : dX rows of source nodes [node_begin,node_end); dX_rows points at the first of them and
 *                            may be memory of a PEER GPU mapped into this process (NVLink peer stores: the partial
 *                            rows land in the owner's stage buffer as they are produced — a fused reduce-scatter).
 *                            addends[0..n_addends) (<= 8, host array of device pointers, each at the first row of the
 *                            range) are summed in the order given and added to the gathered sum in the epilogue:
 *                            dX_rows[k] = Σ_r A_rᵀ dS_r [k] + (Σ_j addends[j][k])   (how the owner folds the partials its
 *                            peers pushed; n_addends=1 with addends[0]==dX_rows is the classic "accumulate").
 *   iddgcn_spmm_prepare    : size the hub-row workspace for feature width D now (needed before a stream capture). */
int iddgcn_spmm_fwd_rows(const iddgcn_graph_t* g, const float* X, int64_t D, const float* vals, int vals_mode,
                         float* S_rows, int64_t row_begin, int64_t row_end, void* stream);
int iddgcn_spmm_bwd_x_rows(const iddgcn_graph_t* g, const float* dS, int64_t D, const float* vals, int vals_mode,
                           float* dX_rows, int64_t node_begin, int64_t node_end,
                           const float* const* addends, int n_addends, void* stream);
int iddgcn_spmm_prepare(const iddgcn_graph_t* g, int64_t D, void* stream);

/* End-to-end variant with HOST buffers (pinned or pageable): uploads X and dS,
 * runs fwd + bwd_x, downloads S and dX; copies are chunked and overlapped with
 * the kernels on internal streams.  Blocking.  Any of S_host/dX_host may be
 * NULL to skip that download. */
int iddgcn_spmm_fwd_bwd_host(const iddgcn_graph_t* g, const float* X_host, const float* dS_host,
                             int64_t D, float* S_host, float* dX_host);

/* DLPack flavour of the two hot entry points: arguments are DLManagedTensor*
 * (dlpack.h v0.x ABI, as produced by tensor.__dlpack__()); dtype float32,
 * device kDLCUDA matching the handle, compact row-major.  The capsules stay
 * owned by the caller (the library never calls the deleter). */
int iddgcn_spmm_fwd_dl(const iddgcn_graph_t* g, void* X_dl, void* vals_dl, int vals_mode,
                       void* S_dl, void* stream);
int iddgcn_spmm_bwd_x_dl(const iddgcn_graph_t* g, void* dS_dl, void* vals_dl, int vals_mode,
                         void* dX_dl, int accumulate, void* stream);
/* DLPack flavours of the layer / model entry points declared further down (same arithmetic, same reference lines:
 * IDDGCN.py:60-79 and IDDGCN.py:147-153); tensor shapes: HT_* / dHT_* [2,T,D] (head plane then tail plane),
 * Z / dZ [R,N,D], W_self / dW_self [D,D], W_alpha / dW_alpha [D,R], b_alpha / db_alpha [R], alpha [T,R], vals [nnz]
 * or NULL, y [T].  D is read from W_self; every shape is validated against the plan / graph / model handle. */
int iddgcn_layer_fwd_dl(const iddgcn_plan_t* p, void* HT_in_dl, void* Z_dl, void* W_self_dl, void* W_alpha_dl,
                        void* b_alpha_dl, void* HT_out_dl, void* alpha_out_dl, void* stream);
int iddgcn_layer_bwd_dl(const iddgcn_plan_t* p, void* HT_in_dl, void* HT_out_dl, void* alpha_dl, void* Z_dl,
                        void* W_self_dl, void* W_alpha_dl, void* dHT_out_dl, void* dHT_in_dl, void* dZ_dl,
                        int accumulate_dZ, void* dW_self_dl, void* dW_alpha_dl, void* db_alpha_dl, void* stream);
int iddgcn_model_forward_dl(iddgcn_model_t* m, const iddgcn_graph_t* g, const iddgcn_plan_t* p, void* vals_dl,
                            int vals_mode, void* y_dl, void* stream);

/* ------------------------------------------------------------------------
 * Triple plan — the per-call index structure the triple stage needs:
 * head/rel/tail as int32 plus the permutations that sort triples by head and
 * by tail, so that every gather's backward (tf.nn.embedding_lookup grad =
 * UnsortedSegmentSum, IDDGCN.py:71-72,226-227) is an atomics-free segment sum.
 *   head, rel, tail   int64 [T] (device if on_device else host)
 * ------------------------------------------------------------------------ */
int iddgcn_plan_create(const int64_t* head, const int64_t* rel, const int64_t* tail, int64_t T,
                       int on_device, int64_t num_entities, int num_relations,
                       void* stream, iddgcn_plan_t** out);
int iddgcn_plan_destroy(iddgcn_plan_t* p);

/* ------------------------------------------------------------------------
 * Layer-level operators (IDDGCN_Layer.call, prediction/IDDGCN.py:60-79) and
 * their backward (what tape.gradient emits for them, IDDGCN.py:172; maths in
 * SURVEY.md Appendix A).  Per-triple activations are (2,T,D) buffers: plane 0
 * = head side, plane 1 = tail side.
 *
 * iddgcn_relation_transform: Z[r] = S[r] · W[r]  ([R,N,D]·[R,D,D]) — the
 *   per-relation transform hoisted from the T gathered rows (IDDGCN.py:76-77)
 *   to the N aggregated rows; same dot products, same values.
 *   _bwd: dW[r] += S[r]ᵀ·dZ[r];  dS[r] (+)= dZ[r]·W[r]ᵀ.
 * iddgcn_dense_transform: C[M,D] (+)= A[M,D]·W (trans_w = 0; W is [D,D]
 *   row-major — tf.matmul(head_e, self_kernel) / tf.matmul(update,
 *   relation_kernel[i]), IDDGCN.py:62-63,76-77) or A·Wᵀ (trans_w = 1, the data
 *   gradient of that product).  path 0 = automatic: fp32 SIMT, or — when
 *   M >= 32768 and D is 64 or 128 — the tensor cores (tcgen05, 3xTF32 split,
 *   ~2^-21 relative per product; IDDGCN_GEMM_TF32X3=0 turns that off);
 *   1 = always fp32 SIMT; 2 = always tensor cores (D in {64,128}).  The layer
 *   and model entry points use the automatic choice.
 * iddgcn_dense_transform_grad: dW[D,D] (+)= Aᵀ[D,M]·dC[M,D], the weight gradient
 *   of the product above (what tape.gradient emits for tf.matmul, IDDGCN.py:172);
 *   a reduction over M rows summed in a fixed order (deterministic).  Same
 *   `path` values; the tensor-core path keeps the accumulation inside the
 *   tensor core to 256 rows at a time and sums those pieces in fp32 registers.
 *   Synchronises the stream before returning (internal scratch).
 * iddgcn_layer_fwd: for every triple k
 *     alpha = softmax(Hin[k]·W_alpha + b_alpha); gate_i = sigmoid(alpha_i)   (:66,:75)
 *     P_H = Hin[k]·W_self + Σ_i gate_i · Z_i[head_k]                          (:62,:76)
 *     P_T = Tin[k]·W_self + Σ_i gate_i · Z_i[tail_k]                          (:63,:77)
 *     Hout = sigmoid(P_H), Tout = sigmoid(P_T)                                (:79)
 *   Hin/Tin may be NULL: then they are gathered from `emb` rows head_k / tail_k
 *   (the Embedding lookups, IDDGCN.py:226-227).  Hout/Tout must be the two
 *   planes of one (2,T,D) buffer.  alpha_out [T,R] is saved for the backward
 *   pass (may be NULL for inference).
 * iddgcn_layer_bwd: dHT_out holds dL/d(Hout,Tout) on entry and is overwritten
 *   with the pre-activation gradients; writes dHT_in; ACCUMULATES dW_self,
 *   dW_alpha, db_alpha; dZ [R,N,D] accumulated or overwritten.
 * iddgcn_segment_scatter: gradient of a row gather without atomics —
 *   out[i][n] (+)= Σ_{k:head_k=n} w_i(k)·dHT[0][k] + Σ_{k:tail_k=n} w_i(k)·dHT[1][k],
 *   w_i(k) = sigmoid(alpha[k][i]) (R planes) or 1 when alpha is NULL (1 plane;
 *   UnsortedSegmentSum of tf.nn.embedding_lookup, IDDGCN.py:71-72,226-227).
 * iddgcn_distmult_fwd/_bwd: DistMult.call (IDDGCN.py:103-109):
 *   y = sigmoid(Σ_d H·Rel[r]·T); bwd writes dHT and ACCUMULATES d_rel_emb.
 * iddgcn_bce_loss_grad: Keras BinaryCrossentropy (clip 1e-7, log(p+1e-7)) on
 *   y = [T_pos positives | negatives] (IDDGCN.py:161-166); loss_out[0] = mean,
 *   dy = d mean/dy · grad_scale (the reference's loss *= 1/N, :168).
 * ------------------------------------------------------------------------ */
int iddgcn_relation_transform(const float* S, const float* W, int R, int64_t N, int64_t D,
                              float* Z, void* stream);
int iddgcn_relation_transform_bwd(const float* S, const float* W, const float* dZ, int R, int64_t N, int64_t D,
                                  float* dW, float* dS, int accumulate_dS, void* stream);
int iddgcn_dense_transform(const float* A, const float* W, int64_t M, int64_t D, int trans_w,
                           float* C, int accumulate, int path, void* stream);
int iddgcn_dense_transform_grad(const float* A, const float* dC, int64_t M, int64_t D, float* dW,
                                int accumulate, int path, void* stream);
int iddgcn_layer_fwd(const iddgcn_plan_t* p, const float* emb,
                     const float* Hin, const float* Tin, const float* Z,
                     const float* W_self, const float* W_alpha, const float* b_alpha,
                     int64_t D, float* Hout, float* Tout, float* alpha_out, void* stream);
int iddgcn_layer_bwd(const iddgcn_plan_t* p, const float* HT_in, const float* HT_out, const float* alpha,
                     const float* Z, const float* W_self, const float* W_alpha, int64_t D,
                     float* dHT_out, float* dHT_in, float* dZ, int accumulate_dZ,
                     float* dW_self, float* dW_alpha, float* db_alpha, void* stream);
int iddgcn_segment_scatter(const iddgcn_plan_t* p, const float* dHT, const float* alpha, int64_t D,
                           float* out, int accumulate, void* stream);
int iddgcn_distmult_fwd(const iddgcn_plan_t* p, const float* HT, const float* rel_emb, int64_t D,
                        float* y, void* stream);
int iddgcn_distmult_bwd(const iddgcn_plan_t* p, const float* HT, const float* rel_emb, const float* y,
                        const float* dy, int64_t D, float* dHT, float* d_rel_emb, void* stream);
int iddgcn_bce_loss_grad(const float* y, int64_t T_pos, int64_t T_total, float grad_scale,
                         float* dy, float* loss_out, void* stream);

/* ------------------------------------------------------------------------
 * Whole-model handle — get_IDDGCN_Model (IDDGCN.py:201-285): Embedding(N,D),
 * L IDDGCN_Layer all fed the raw embedding table, DistMult decoder
 * (IDDGCN.py:82-109); IDDGCN_Model.train_step (IDDGCN.py:123-178): positive +
 * negative forward, BinaryCrossentropy (clip 1e-7) · 1/N, backward, Keras Adam.
 *
 * Parameters live in ONE flat fp32 device buffer in Keras trainable_weights
 * order, each tensor padded to a multiple of 4 floats (so Adam is a single
 * fused launch and .h5 I/O is a slice table, iddgcn_model_param_layout):
 *   entity_embeddings [N,D];
 *   per layer l=0..L-1: relation_kernels [R,D,D], self_kernel [D,D],
 *                       relation_weights [R] (unused by call, IDDGCN.py:39-44),
 *                       W_alpha [D,R], b_alpha [R];
 *   DistMult rel_embedding [R,D].
 * ------------------------------------------------------------------------ */
int iddgcn_model_create(int64_t num_entities, int num_relations, int64_t dim, int num_layers,
                        iddgcn_model_t** out);
int iddgcn_model_destroy(iddgcn_model_t* m);
/* length of the flat buffer (floats, including padding) */
int64_t iddgcn_model_num_params(const iddgcn_model_t* m);
/* offsets[2+5L], sizes[2+5L]: start and element count of each weight in Keras order */
int iddgcn_model_param_layout(const iddgcn_model_t* m, int64_t* offsets, int64_t* sizes);
/* device pointer to the flat parameter buffer / its gradient / Adam moments */
int iddgcn_model_buffers(iddgcn_model_t* m, float** params, float** grads, float** adam_m, float** adam_v);
/* drop activations cached for a plan (call before destroying the plan) */
int iddgcn_model_release_plan(iddgcn_model_t* m, const iddgcn_plan_t* p);
/* drop the train step captured on graph g (call before destroying g).  The captured step is keyed on unique handle
 * ids and on a workspace generation counter, so a recycled heap address or a regrown workspace never replays it. */
int iddgcn_model_release_graph(iddgcn_model_t* m, const iddgcn_graph_t* g);
/* Per-kernel-group timing of ONE iddgcn_model_train_step (measurement aid for the layer-level roofline; the groups are
 * the phases of IDDGCN.py:123-178 in execution order): enable, run a step, read IDDGCN_PROFILE_GROUPS floats (ms). */
#define IDDGCN_PROFILE_GROUPS 21
int iddgcn_model_profile(iddgcn_model_t* m, int enable);
int iddgcn_model_profile_read(iddgcn_model_t* m, float* ms_by_group, int n_groups);
/* y [T] = model(head, rel, tail | graph, optional edge vals)   (IDDGCN.py:147-153, IDDGCN_eval.py:97-105) */
int iddgcn_model_forward(iddgcn_model_t* m, const iddgcn_graph_t* g, const iddgcn_plan_t* p,
                         const float* vals, int vals_mode, float* y, void* stream);
/* One training step (IDDGCN.py:123-178).  loss_out: DEVICE float[1] receiving the
 * BCE mean BEFORE the 1/N scaling (what tf.print(loss) shows, :167); y_out
 * [T_pos+T_neg] optional (device).  step_index t>=1 drives Adam bias correction.
 * apply_update=0 leaves parameters untouched (gradients only).  neg may be NULL. */
int iddgcn_model_train_step(iddgcn_model_t* m, const iddgcn_graph_t* g,
                            const iddgcn_plan_t* pos, const iddgcn_plan_t* neg,
                            float lr, float beta1, float beta2, float eps, int64_t step_index,
                            int apply_update, float* loss_out, float* y_out, void* stream);
/* num_steps training steps; the step is captured once into a CUDA graph and
 * replayed (model.fit(epochs=5000) of IDDGCN.py:399-412 is launch-bound on a B200).
 * loss_history: DEVICE float[num_steps] or NULL. */
int iddgcn_model_train_loop(iddgcn_model_t* m, const iddgcn_graph_t* g,
                            const iddgcn_plan_t* pos, const iddgcn_plan_t* neg,
                            float lr, float beta1, float beta2, float eps,
                            int64_t first_step_index, int64_t num_steps, float* loss_history, void* stream);
/* Backward of a scalar objective through model_forward for one plan:
 * dy [T] is ∂L/∂y.  Accumulates parameter grads into the grads buffer
 * (zero_grads!=0 clears it first) and, if dvals!=NULL, writes ∂L/∂vals [nnz]
 * (explaiNE.py:17; GnnExplainer.py:51). Must follow a model_forward call with
 * the same (g,p,vals) on the same stream. */
int iddgcn_model_backward(iddgcn_model_t* m, const iddgcn_graph_t* g, const iddgcn_plan_t* p,
                          const float* vals, int vals_mode, const float* dy,
                          int zero_grads, float* dvals, void* stream);
/* Keras Adam on an arbitrary flat buffer (OptimizerV2 Adam: lr_t = lr·sqrt(1-b2^t)/(1-b1^t),
 * theta -= lr_t·m/(sqrt(v)+eps); IDDGCN.py:174, GnnExplainer.py:52, IDDGCN_explain.py:85) */
int iddgcn_adam_step(float* theta, const float* grad, float* m, float* v, int64_t n,
                     float lr, float beta1, float beta2, float eps, int64_t step_index, void* stream);

/* ------------------------------------------------------------------------
 * Explainer support — explanation/GnnExplainer.py:13-74, IDDGCN_explain.py:16-120.
 * The reference explains one test triple at a time on its 1-hop computation
 * graph.  Here B triples are batched block-diagonally: triple b's sub-graph
 * uses node ids [b·N, (b+1)·N) of one graph with B·N nodes (embeddings tiled),
 * so one forward/backward of the kernels above serves the whole batch.
 *
 * iddgcn_subgraph_triples: emits, for every triple b, all stored edges of the
 *   full graph g that touch head[b] or tail[b] (get_computation_graph,
 *   GnnExplainer.py:18-24) as int64 (head+b·N, rel, tail+b·N) rows, duplicates
 *   included (graph_build removes them).  head/tail: DEVICE int64[B].  Call with
 *   triples_out = NULL to obtain *M_out, then with a DEVICE buffer of M rows.
 * iddgcn_block_relation_sums: c[r][b] = Σ a_e over block b of relation r
 *   (structure_loss counts, IDDGCN_explain.py:31-34), c is [R,B].
 * iddgcn_explain_loss: loss_b = -before_b·log(pred_b+1e-5) (GnnExplainer.py:48);
 *   with targets != NULL: (that + mean_i (target_i - c_ib/Σ_i c_ib)²)/2
 *   (IDDGCN_explain.py:68-73).  Writes loss[B], dy[B] = ∂loss/∂pred and
 *   coef[R,B] = ∂loss/∂c.
 * iddgcn_structure_grad: dvals[e] += coef[rel_e][block_e]·∂a_e/∂vals[e].
 * ------------------------------------------------------------------------ */
int iddgcn_subgraph_triples(const iddgcn_graph_t* g, const int64_t* head, const int64_t* tail, int64_t B,
                            int64_t* triples_out, int64_t capacity_rows, int64_t* M_out, void* stream);
int iddgcn_block_relation_sums(const iddgcn_graph_t* g, const float* vals, int vals_mode, int64_t N, int64_t B,
                               float* out, void* stream);
int iddgcn_explain_loss(const float* before, const float* pred, const float* c, const float* targets, int R,
                        int64_t B, float* loss, float* dy, float* coef, void* stream);
int iddgcn_structure_grad(const iddgcn_graph_t* g, const float* vals, int vals_mode, const float* coef,
                          int64_t N, int64_t B, float* dvals, void* stream);

/* ------------------------------------------------------------------------
 * Exchange kernels over NVSwitch multicast memory (NVLS) for the row-partitioned
 * SpMM (no counterpart in the reference: its only multi-device construct,
 * tf.distribute.MirroredStrategy in GnnExplainer.py:105, is pinned to one GPU).
 *   iddgcn_mc_store:        all-gather leg — store n local floats once through the
 *                           MULTICAST address mc_dst; the switch replicates them
 *                           into every rank's copy of the symmetric buffer.
 *   iddgcn_mc_reduce_load:  reduce-scatter leg — dst[i] = Σ_ranks buf_rank[i] for the
 *                           range at MULTICAST address mc_src (in-switch sum).
 * n % 4 == 0, 16-byte aligned; `ctas` = thread blocks to use (<=0: 32).  The
 * multicast mapping and cross-rank barriers are the host layer's job.
 * ------------------------------------------------------------------------ */
int iddgcn_mc_store(void* mc_dst, const float* src, int64_t n, int ctas, void* stream);
int iddgcn_mc_reduce_load(const void* mc_src, float* dst, int64_t n, int ctas, void* stream);

/* ------------------------------------------------------------------------
 * Plumbing of the fused peer-memory exchange of the grid-partitioned SpMM (dist.py; no reference counterpart, see
 * above).  Ranks order their steps with 32-bit epoch flags in symmetric memory instead of host barriers:
 *   iddgcn_peer_signal : once everything enqueued before it on `stream` is done, store `value` (release, system scope)
 *                        into remote_slots[0..n) — one slot on each peer (peer-mapped addresses).
 *   iddgcn_peer_wait   : hold `stream` until every local_slots[i] has reached `value` (wrap-safe compare); after
 *                        timeout_ms (<=0: 20000) it gives up and sets *timed_out_dev (device int, may be NULL) to 1.
 *   iddgcn_copy_async / iddgcn_copy2d_async : cudaMemcpy(2D)Async with cudaMemcpyDefault — copy-engine transfers
 *                        between any of host / device / peer memory (the all-gather push and the host pipelines).
 *   iddgcn_peer_push   : the same push done by `ctas` thread blocks with 16-byte stores (A/B alternative to the DMA).
 * n <= 8 peers.
 * ------------------------------------------------------------------------ */
int iddgcn_peer_signal(void* const* remote_slots, int n, uint32_t value, void* stream);
int iddgcn_peer_wait(void* const* local_slots, int n, uint32_t value, int timeout_ms, int* timed_out_dev,
                     void* stream);
int iddgcn_copy_async(void* dst, const void* src, int64_t bytes, void* stream);
int iddgcn_copy2d_async(void* dst, int64_t dpitch, const void* src, int64_t spitch, int64_t width_bytes,
                        int64_t rows, void* stream);
int iddgcn_peer_push(float* dst_peer, const float* src, int64_t n, int ctas, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IDDGCN_H_ */
