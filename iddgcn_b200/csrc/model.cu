// model.cu — whole-model handle: get_IDDGCN_Model wiring (prediction/IDDGCN.py:201-285) and
// IDDGCN_Model.train_step (IDDGCN.py:123-178), orchestrating the kernels of spmm.cu / dense.cu / triple.cu.
//
// What is hoisted relative to the reference's op order (values unchanged):
//   * S_i = A_i·E is computed ONCE per step (the reference recomputes it 3 layers x 2 calls x R times);
//   * the per-relation transform is applied to the N aggregated rows (Z_i = S_i·W_i) instead of the T gathered
//     rows — the same row·matrix dot products, 44x fewer of them at the shipped sizes;
//   * head and tail sides share one (2,T,D) buffer so every self-kernel matmul is one GEMM.
#include <unordered_map>
#include <vector>
#include "dense.cuh"
#include "dlpack.cuh"
#include "graph.cuh"
#include "triple.cuh"

namespace iddgcn {
int spmm_fwd_range(const iddgcn_graph* g, const float* X, int64_t D, const float* vals, int mode, float* S,
                   int64_t row_begin, int64_t row_end, cudaStream_t st);
int spmm_bwd_x_range(const iddgcn_graph* g, const float* dS, int64_t D, const float* vals, int mode, float* dX,
                     int accumulate, int64_t row_begin, int64_t row_end, cudaStream_t st);
int spmm_prepare(const iddgcn_graph* g, int64_t D, cudaStream_t st);
}  // namespace iddgcn

using namespace iddgcn;

struct LayerOffsets { int64_t rk, sk, rw, wa, ba; };
struct ParamLayout {
  int64_t emb = 0;
  std::vector<LayerOffsets> layer;
  int64_t rel_emb = 0;
  int64_t total = 0;
};

struct TripleWS {
  int64_t T = 0;
  const iddgcn_plan* owner = nullptr;  // the plan whose activations are cached here (forward -> backward pairing)
  std::vector<float*> HT;     // L+1 buffers (2,T,D)
  std::vector<float*> alpha;  // L buffers (T,R)
  float* dHT[2] = {nullptr, nullptr};
  float* dz = nullptr;
  float* y = nullptr;
  void release() {
    for (float* p : HT) cudaFree(p);
    for (float* p : alpha) cudaFree(p);
    cudaFree(dHT[0]); cudaFree(dHT[1]); cudaFree(dz); cudaFree(y);
    HT.clear(); alpha.clear(); dHT[0] = dHT[1] = dz = y = nullptr; T = 0; owner = nullptr;
  }
};

struct iddgcn_model {
  int device = 0;
  int64_t N = 0, D = 0;
  int R = 0, L = 0;
  ParamLayout lay;
  float *params = nullptr, *grads = nullptr, *adam_m = nullptr, *adam_v = nullptr;
  int64_t* step_dev = nullptr;
  float* loss_dev = nullptr;
  // node-level workspace
  float *S = nullptr, *Z = nullptr, *dZ = nullptr, *dS = nullptr;
  DevBuf<float> gemm_ws, small_ws, y_all, dy_all, wt_ws;   // wt_ws: transposed weights [R,D,D]
  // activation workspaces are keyed by the NUMBER of triples, not by plan: a stream of same-sized plans (explainer
  // batches) reuses one set of buffers instead of allocating and freeing ~13 of them per plan
  std::unordered_map<int64_t, TripleWS> ws;
  // positives and negatives of a train step run as ONE batch of triples: (pos, neg) -> concatenated plan
  struct Merged { const iddgcn_plan* pos; const iddgcn_plan* neg; iddgcn_plan* both; };
  std::vector<Merged> merged;
  // captured training step.  The CUDA graph bakes in raw device pointers (workspaces of this handle, the graph's
  // CSR arrays and hub workspace, the plans' index arrays), so it is keyed on the handles' UNIQUE IDS (never on heap
  // addresses, which are recycled) and on ws_gen, which every (re)allocation of a captured workspace bumps.
  cudaGraphExec_t graph_exec = nullptr;
  uint64_t graph_uid[3] = {0, 0, 0};
  uint64_t graph_gen = 0;
  const void* graph_heavy_ws = nullptr;
  float graph_hyper[4] = {0, 0, 0, 0};
  uint64_t ws_gen = 1;
  cudaStream_t own_stream = nullptr;

  // optional per-group timing of one train step (iddgcn_model_profile): an event after every kernel group
  bool prof_on = false;
  std::vector<cudaEvent_t> prof_ev;
  std::vector<int> prof_tag;
  size_t prof_n = 0;

  float* P(int64_t off) const { return params + off; }
  float* G(int64_t off) const { return grads + off; }
  size_t plane() const { return (size_t)R * N * D; }
};

// kernel groups of a train step (tools/bench_layer.py holds the byte / flop model of each)
enum ProfTag {
  PT_START = 0, PT_ZERO, PT_SPMM_FWD, PT_REL_TRANSFORM, PT_GATHER_EMB, PT_SELF_TRANSFORM, PT_COMBINE, PT_DISTMULT_FWD,
  PT_BCE, PT_DISTMULT_BWD, PT_BWD_PRE, PT_DW_SELF, PT_DW_ALPHA, PT_SCATTER_DZ, PT_DIN, PT_BWD_POST, PT_SCATTER_EMB,
  PT_DW_REL, PT_DS, PT_SPMM_BWD, PT_ADAM, PT_COUNT
};

static void prof_mark(iddgcn_model* m, int tag, cudaStream_t st) {
  if (!m->prof_on) return;
  if (m->prof_n == m->prof_ev.size()) {
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    m->prof_ev.push_back(e);
    m->prof_tag.push_back(0);
  }
  m->prof_tag[m->prof_n] = tag;
  cudaEventRecord(m->prof_ev[m->prof_n++], st);
}

static int64_t pad4(int64_t n) { return (n + 3) / 4 * 4; }

static ParamLayout make_layout(int64_t N, int R, int64_t D, int L) {
  ParamLayout p;
  int64_t o = 0;
  p.emb = o; o += pad4(N * D);
  for (int l = 0; l < L; ++l) {
    LayerOffsets lo;
    lo.rk = o; o += pad4((int64_t)R * D * D);
    lo.sk = o; o += pad4(D * D);
    lo.rw = o; o += pad4(R);
    lo.wa = o; o += pad4(D * R);
    lo.ba = o; o += pad4(R);
    p.layer.push_back(lo);
  }
  p.rel_emb = o; o += pad4((int64_t)R * D);
  p.total = o;
  return p;
}

static void model_free(iddgcn_model* m) {
  if (!m) return;
  if (m->graph_exec) cudaGraphExecDestroy(m->graph_exec);
  if (m->own_stream) cudaStreamDestroy(m->own_stream);
  for (cudaEvent_t e : m->prof_ev) cudaEventDestroy(e);
  for (auto& kv : m->ws) kv.second.release();
  for (auto& e : m->merged) plan_destroy_internal(e.both);
  cudaFree(m->params); cudaFree(m->grads); cudaFree(m->adam_m); cudaFree(m->adam_v);
  cudaFree(m->step_dev); cudaFree(m->loss_dev);
  cudaFree(m->S); cudaFree(m->Z); cudaFree(m->dZ); cudaFree(m->dS);
  delete m;
}

constexpr size_t MAX_TRIPLE_WS = 8;   // distinct triple counts kept alive per model

static int ensure_triple_ws(iddgcn_model* m, const iddgcn_plan* p, TripleWS** out) {
  if (m->ws.find(p->T) == m->ws.end() && m->ws.size() >= MAX_TRIPLE_WS) {
    // a stream of differently sized plans would otherwise grow the map without bound: drop every other size
    IDD_CUDA(cudaDeviceSynchronize());
    for (auto& kv : m->ws) kv.second.release();
    m->ws.clear();
    ++m->ws_gen;
  }
  TripleWS& w = m->ws[p->T];
  if (w.T != p->T || w.HT.empty()) {
    w.release();
    ++m->ws_gen;
    const size_t td = (size_t)(p->T ? p->T : 1) * m->D;
    const size_t tr = (size_t)(p->T ? p->T : 1) * m->R;
    for (int l = 0; l <= m->L; ++l) {
      float* q = nullptr;
      IDD_CUDA(cudaMalloc((void**)&q, 2 * td * sizeof(float)));
      w.HT.push_back(q);
    }
    for (int l = 0; l < m->L; ++l) {
      float* q = nullptr;
      IDD_CUDA(cudaMalloc((void**)&q, tr * sizeof(float)));
      w.alpha.push_back(q);
    }
    IDD_CUDA(cudaMalloc((void**)&w.dHT[0], 2 * td * sizeof(float)));
    IDD_CUDA(cudaMalloc((void**)&w.dHT[1], 2 * td * sizeof(float)));
    IDD_CUDA(cudaMalloc((void**)&w.dz, tr * sizeof(float)));
    IDD_CUDA(cudaMalloc((void**)&w.y, (size_t)(p->T ? p->T : 1) * sizeof(float)));
    w.T = p->T;
  }
  // shared scratch sized for this plan
  size_t gw = gemm_tn_workspace_elems(m->D, m->D, 2 * p->T, 1);
  size_t g2 = gemm_tn_workspace_elems(m->D, m->R, p->T, 1);
  size_t g3 = gemm_tn_workspace_elems(m->D, m->D, m->N, m->R);
  if (g2 > gw) gw = g2;
  if (g3 > gw) gw = g3;
  if (gw > m->gemm_ws.n) { IDD_CUDA(m->gemm_ws.ensure(gw)); ++m->ws_gen; }
  size_t sw = colsum_scratch_elems(p->T, m->R);
  size_t s2 = distmult_scratch_elems(p, m->D);
  if (s2 > sw) sw = s2;
  if (sw > m->small_ws.n) { IDD_CUDA(m->small_ws.ensure(sw)); ++m->ws_gen; }
  *out = &w;
  return IDDGCN_OK;
}

static int check_compat(const iddgcn_model* m, const iddgcn_graph* g, const iddgcn_plan* p) {
  IDD_CHECK(m && g && p, IDDGCN_EINVAL, "model, graph or plan is NULL");
  IDD_CHECK(g->N == m->N && g->R == m->R && g->n_rows == m->N && g->row_lo == 0, IDDGCN_EINVAL,
            "graph (N=%lld,R=%d,rows=%lld) does not match model (N=%lld,R=%d); the model path needs the whole graph",
            (long long)g->N, g->R, (long long)g->n_rows, (long long)m->N, m->R);
  IDD_CHECK(p->N == m->N && p->R == m->R, IDDGCN_EINVAL, "plan (N=%lld,R=%d) does not match model",
            (long long)p->N, p->R);
  return IDDGCN_OK;
}

// S = A·E and Z_l = S·W_l for every layer
static int node_forward(iddgcn_model* m, const iddgcn_graph* g, const float* vals, int mode, cudaStream_t st) {
  IDD_TRY(spmm_fwd_range(g, m->P(m->lay.emb), m->D, vals, mode, m->S, 0, (int64_t)m->R * m->N, st));
  prof_mark(m, PT_SPMM_FWD, st);
  for (int l = 0; l < m->L; ++l)
    IDD_TRY(gemm_f32(false, false, m->N, m->D, m->D, m->S, m->D, m->N * m->D, m->P(m->lay.layer[l].rk), m->D,
                     m->D * m->D, m->Z + (size_t)l * m->plane(), m->D, m->N * m->D, m->R, false, st));
  prof_mark(m, PT_REL_TRANSFORM, st);
  return IDDGCN_OK;
}

static int triple_forward(iddgcn_model* m, const iddgcn_plan* p, TripleWS* w, float* y_out, cudaStream_t st) {
  w->owner = p;
  if (p->T == 0) return IDDGCN_OK;
  IDD_TRY(gather_embeddings(p, m->P(m->lay.emb), m->D, w->HT[0], st));
  prof_mark(m, PT_GATHER_EMB, st);
  for (int l = 0; l < m->L; ++l) {
    const LayerOffsets& lo = m->lay.layer[l];
    IDD_TRY(gemm_f32(false, false, 2 * p->T, m->D, m->D, w->HT[l], m->D, 0, m->P(lo.sk), m->D, 0, w->HT[l + 1], m->D,
                     0, 1, false, st));
    prof_mark(m, PT_SELF_TRANSFORM, st);
    IDD_TRY(layer_combine(p, w->HT[l], m->Z + (size_t)l * m->plane(), m->P(lo.wa), m->P(lo.ba), m->D, w->HT[l + 1],
                          w->alpha[l], st));
    prof_mark(m, PT_COMBINE, st);
  }
  IDD_TRY(distmult_fwd(p, w->HT[m->L], m->P(m->lay.rel_emb), m->D, y_out, st));
  prof_mark(m, PT_DISTMULT_FWD, st);
  return IDDGCN_OK;
}

// per-triple backward; accumulates into grads (sk, wa, ba, rel_emb, emb-lookup part) and into dZ
static int triple_backward(iddgcn_model* m, const iddgcn_plan* p, TripleWS* w, const float* y, const float* dy,
                           cudaStream_t st) {
  if (p->T == 0) return IDDGCN_OK;
  const int64_t T = p->T, D = m->D;
  int cur = 0;
  IDD_TRY(distmult_bwd(p, w->HT[m->L], m->P(m->lay.rel_emb), y, dy, D, w->dHT[cur], m->G(m->lay.rel_emb),
                       m->small_ws.p, st));
  prof_mark(m, PT_DISTMULT_BWD, st);
  for (int l = m->L - 1; l >= 0; --l) {
    const LayerOffsets& lo = m->lay.layer[l];
    const float* Zl = m->Z + (size_t)l * m->plane();
    IDD_TRY(layer_bwd_pre(p, w->HT[l + 1], w->alpha[l], Zl, D, w->dHT[cur], w->dz, st));
    prof_mark(m, PT_BWD_PRE, st);
    float* dP = w->dHT[cur];
    // dW_self = [Hin;Tin]ᵀ·dP  (eq. 8), dW_alpha = Hinᵀ·dz, db_alpha = colsum(dz)  (eq. 7)
    IDD_TRY(gemm_tn_splitk_f32(D, D, 2 * T, w->HT[l], D, 0, dP, D, 0, m->G(lo.sk), D, 0, 1, true, m->gemm_ws.p, st));
    prof_mark(m, PT_DW_SELF, st);
    IDD_TRY(gemm_tn_splitk_f32(D, m->R, T, w->HT[l], D, 0, w->dz, m->R, 0, m->G(lo.wa), m->R, 0, 1, true,
                               m->gemm_ws.p, st));
    IDD_TRY(colsum(w->dz, T, m->R, m->G(lo.ba), 1, m->small_ws.p, st));
    prof_mark(m, PT_DW_ALPHA, st);
    // dZ_i += scatter(gate_i · dP)  (eq. 10)
    IDD_TRY(segment_scatter(p, dP, w->alpha[l], D, m->dZ + (size_t)l * m->plane(), 1, st));
    prof_mark(m, PT_SCATTER_DZ, st);
    // d[Hin;Tin] = dP·W_selfᵀ (+ dz·W_alphaᵀ on the head side)  (eq. 9)
    IDD_TRY(transpose_small(m->P(lo.sk), m->wt_ws.p, (int)D, (int)D, 1, st));
    IDD_TRY(gemm_f32(false, false, 2 * T, D, D, dP, D, 0, m->wt_ws.p, D, 0, w->dHT[cur ^ 1], D, 0, 1, false, st));
    prof_mark(m, PT_DIN, st);
    IDD_TRY(layer_bwd_post(w->dz, m->P(lo.wa), T, D, m->R, w->dHT[cur ^ 1], st));
    prof_mark(m, PT_BWD_POST, st);
    cur ^= 1;
  }
  // gradient of the two Embedding lookups (eq. 12, second and third term)
  IDD_TRY(segment_scatter(p, w->dHT[cur], nullptr, D, m->G(m->lay.emb), 1, st));
  prof_mark(m, PT_SCATTER_EMB, st);
  return IDDGCN_OK;
}

// node-level backward: dW_i = S_iᵀ·dZ_i, dS_i = Σ_l dZ_i·W_iᵀ (eq. 11), dE += Σ_i A_iᵀ·dS_i (eq. 12), dvals (eq. 13)
static int node_backward(iddgcn_model* m, const iddgcn_graph* g, const float* vals, int mode, float* dvals,
                         cudaStream_t st);

static int zero_step_buffers(iddgcn_model* m, bool grads_too, cudaStream_t st) {
  if (grads_too) IDD_CUDA(cudaMemsetAsync(m->grads, 0, (size_t)m->lay.total * sizeof(float), st));
  IDD_CUDA(cudaMemsetAsync(m->dZ, 0, (size_t)m->L * m->plane() * sizeof(float), st));
  return IDDGCN_OK;
}

namespace iddgcn {
int sddmm_dispatch(const iddgcn_graph* g, const float* dS, const float* X, int64_t D, const float* vals, int mode,
                   float* dvals, cudaStream_t st);
}

static int node_backward(iddgcn_model* m, const iddgcn_graph* g, const float* vals, int mode, float* dvals,
                         cudaStream_t st) {
  const int64_t N = m->N, D = m->D;
  for (int l = 0; l < m->L; ++l) {
    const LayerOffsets& lo = m->lay.layer[l];
    float* dZl = m->dZ + (size_t)l * m->plane();
    IDD_TRY(gemm_tn_splitk_f32(D, D, N, m->S, D, N * D, dZl, D, N * D, m->G(lo.rk), D, D * D, m->R, true,
                               m->gemm_ws.p, st));
    prof_mark(m, PT_DW_REL, st);
    IDD_TRY(transpose_small(m->P(lo.rk), m->wt_ws.p, (int)D, (int)D, m->R, st));
    IDD_TRY(gemm_f32(false, false, N, D, D, dZl, D, N * D, m->wt_ws.p, D, D * D, m->dS, D, N * D, m->R, l > 0, st));
    prof_mark(m, PT_DS, st);
  }
  IDD_TRY(spmm_bwd_x_range(g, m->dS, D, vals, mode, m->G(m->lay.emb), 1, 0, N, st));
  prof_mark(m, PT_SPMM_BWD, st);
  if (dvals) IDD_TRY(sddmm_dispatch(g, m->dS, m->P(m->lay.emb), D, vals, mode, dvals, st));
  return IDDGCN_OK;
}

// One batch holds the positives followed by the negatives (same parameters, same graph): every per-triple kernel
// runs once per step instead of twice; only the loss distinguishes the two halves (labels by position).
static int train_step_body(iddgcn_model* m, const iddgcn_graph* g, const iddgcn_plan* both, int64_t Tp, TripleWS* w,
                           float lr, float b1, float b2, float eps, int64_t step, bool step_on_device,
                           int apply_update, float* loss_out, float* y_out, cudaStream_t st) {
  const int64_t Tt = both->T;
  m->prof_n = 0;
  prof_mark(m, PT_START, st);
  IDD_TRY(zero_step_buffers(m, true, st));
  prof_mark(m, PT_ZERO, st);
  IDD_TRY(node_forward(m, g, nullptr, IDDGCN_VALS_ONES, st));
  IDD_TRY(triple_forward(m, both, w, m->y_all.p, st));                       // IDDGCN.py:147-160
  // y_pred = concat, y_true = [1..1 | 0..0], loss = BCE; loss *= 1/N  (IDDGCN.py:161-168)
  IDD_TRY(bce_loss_grad(m->y_all.p, Tp, Tt, 1.0f / (float)m->N, m->dy_all.p, m->loss_dev, m->small_ws.p, st));
  prof_mark(m, PT_BCE, st);
  IDD_TRY(triple_backward(m, both, w, m->y_all.p, m->dy_all.p, st));
  IDD_TRY(node_backward(m, g, nullptr, IDDGCN_VALS_ONES, nullptr, st));
  if (apply_update) {
    if (step_on_device) IDD_TRY(increment_counter(m->step_dev, st));
    IDD_TRY(adam_apply(m->params, m->grads, m->adam_m, m->adam_v, m->lay.total, lr, b1, b2, eps, step,
                       step_on_device ? m->step_dev : nullptr, st));          // IDDGCN.py:174
    prof_mark(m, PT_ADAM, st);
  }
  if (loss_out) IDD_CUDA(cudaMemcpyAsync(loss_out, m->loss_dev, sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (y_out) IDD_CUDA(cudaMemcpyAsync(y_out, m->y_all.p, (size_t)Tt * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return IDDGCN_OK;
}

static int prepare_train(iddgcn_model* m, const iddgcn_graph* g, const iddgcn_plan* pos, const iddgcn_plan* neg,
                         cudaStream_t st, const iddgcn_plan** both, TripleWS** w) {
  IDD_TRY(check_compat(m, g, pos));
  if (neg) IDD_TRY(check_compat(m, g, neg));
  IDD_CHECK(pos->T + (neg ? neg->T : 0) > 0, IDDGCN_EINVAL, "train step needs at least one triple");
  const iddgcn_plan* b = pos;
  if (neg && neg->T > 0) {
    b = nullptr;
    for (auto& e : m->merged)
      if (e.pos == pos && e.neg == neg && e.both->T == pos->T + neg->T) b = e.both;
    if (!b) {
      iddgcn_plan* np = nullptr;
      IDD_TRY(plan_concat(pos, neg, st, &np));
      m->merged.push_back({pos, neg, np});
      b = np;
    }
  }
  IDD_TRY(ensure_triple_ws(m, b, w));
  const size_t Tt = (size_t)b->T;
  if (Tt > m->y_all.n) { IDD_CUDA(m->y_all.ensure(Tt)); ++m->ws_gen; }
  if (Tt > m->dy_all.n) { IDD_CUDA(m->dy_all.ensure(Tt)); ++m->ws_gen; }
  const size_t bs = bce_scratch_elems((int64_t)Tt);
  if (bs > m->small_ws.n) { IDD_CUDA(m->small_ws.ensure(bs)); ++m->ws_gen; }
  // hub rows of a power-law graph need a split-pass workspace; size it now, a capture could not allocate it
  IDD_TRY(spmm_prepare(g, m->D, st));
  // the tensor-core GEMM allocates its weight-split scratch lazily, which a stream capture would forbid
  if ((m->D == 64 || m->D == 128) && (m->N >= 32768 || 2 * b->T >= 32768) && tc_gemm_enabled())
    IDD_TRY(tc_gemm_prepare(st));
  *both = b;
  return IDDGCN_OK;
}

extern "C" {

int iddgcn_model_create(int64_t num_entities, int num_relations, int64_t dim, int num_layers, iddgcn_model_t** out) {
  IDD_CHECK(out != nullptr, IDDGCN_EINVAL, "out is NULL");
  *out = nullptr;
  IDD_CHECK(num_entities > 0 && num_relations > 0 && dim > 0 && num_layers > 0, IDDGCN_EINVAL, "sizes must be positive");
  IDD_CHECK(num_relations <= RMAX, IDDGCN_ELIMIT, "the model path supports at most %d relations", RMAX);
  IDD_CHECK(dim % 4 == 0 && dim <= 512, IDDGCN_ELIMIT, "embedding dim must be a multiple of 4 and <= 512");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("no CUDA device: libiddgcn_b200 has no CPU fallback");
    return IDDGCN_ECUDA;
  }
  iddgcn_model* m = new iddgcn_model();
  m->N = num_entities; m->R = num_relations; m->D = dim; m->L = num_layers;
  m->lay = make_layout(m->N, m->R, m->D, m->L);
  auto fail = [&](int rc) { model_free(m); return rc; };
  auto alloc0 = [&](float** p, size_t n) -> int {
    IDD_CUDA(cudaMalloc((void**)p, n * sizeof(float)));
    IDD_CUDA(cudaMemset(*p, 0, n * sizeof(float)));
    return IDDGCN_OK;
  };
  int rc;
  if (cudaGetDevice(&m->device) != cudaSuccess) return fail(IDDGCN_ECUDA);
  const size_t tot = (size_t)m->lay.total;
  if ((rc = alloc0(&m->params, tot)) || (rc = alloc0(&m->grads, tot)) || (rc = alloc0(&m->adam_m, tot)) ||
      (rc = alloc0(&m->adam_v, tot)) || (rc = alloc0(&m->S, m->plane())) || (rc = alloc0(&m->dS, m->plane())) ||
      (rc = alloc0(&m->Z, m->L * m->plane())) || (rc = alloc0(&m->dZ, m->L * m->plane())) ||
      (rc = alloc0(&m->loss_dev, 4)))
    return fail(rc);
  if (m->wt_ws.ensure((size_t)m->R * m->D * m->D) != cudaSuccess) {
    set_error("cudaMalloc(weight transpose scratch) failed");
    return fail(IDDGCN_ENOMEM);
  }
  if (cudaMalloc((void**)&m->step_dev, sizeof(int64_t)) != cudaSuccess ||
      cudaMemset(m->step_dev, 0, sizeof(int64_t)) != cudaSuccess) {
    set_error("cudaMalloc(step counter) failed");
    return fail(IDDGCN_ENOMEM);
  }
  *out = m;
  return IDDGCN_OK;
}

int iddgcn_model_destroy(iddgcn_model_t* m) {
  model_free(m);
  return IDDGCN_OK;
}

int64_t iddgcn_model_num_params(const iddgcn_model_t* m) { return m ? m->lay.total : 0; }

// offsets[2+5L] (floats, into the flat buffer) in Keras weight order; sizes likewise (unpadded element counts)
int iddgcn_model_param_layout(const iddgcn_model_t* m, int64_t* offsets, int64_t* sizes) {
  IDD_CHECK(m && offsets && sizes, IDDGCN_EINVAL, "NULL argument");
  int k = 0;
  offsets[k] = m->lay.emb; sizes[k++] = m->N * m->D;
  for (int l = 0; l < m->L; ++l) {
    const LayerOffsets& lo = m->lay.layer[l];
    offsets[k] = lo.rk; sizes[k++] = (int64_t)m->R * m->D * m->D;
    offsets[k] = lo.sk; sizes[k++] = m->D * m->D;
    offsets[k] = lo.rw; sizes[k++] = m->R;
    offsets[k] = lo.wa; sizes[k++] = m->D * m->R;
    offsets[k] = lo.ba; sizes[k++] = m->R;
  }
  offsets[k] = m->lay.rel_emb; sizes[k++] = (int64_t)m->R * m->D;
  return IDDGCN_OK;
}

int iddgcn_model_buffers(iddgcn_model_t* m, float** params, float** grads, float** adam_m, float** adam_v) {
  IDD_CHECK(m != nullptr, IDDGCN_EINVAL, "model is NULL");
  if (params) *params = m->params;
  if (grads) *grads = m->grads;
  if (adam_m) *adam_m = m->adam_m;
  if (adam_v) *adam_v = m->adam_v;
  return IDDGCN_OK;
}

// drop the cached activations of a plan (call before destroying the plan)
int iddgcn_model_release_plan(iddgcn_model_t* m, const iddgcn_plan_t* p) {
  IDD_CHECK(m != nullptr, IDDGCN_EINVAL, "model is NULL");
  for (auto& kv : m->ws)
    if (kv.second.owner == p) kv.second.owner = nullptr;   // buffers stay for the next plan of this size
  for (size_t i = 0; i < m->merged.size();) {
    if (m->merged[i].pos == p || m->merged[i].neg == p) {
      cudaDeviceSynchronize();
      for (auto& kv : m->ws)
        if (kv.second.owner == m->merged[i].both) kv.second.owner = nullptr;
      plan_destroy_internal(m->merged[i].both);
      m->merged.erase(m->merged.begin() + i);
    } else {
      ++i;
    }
  }
  if (m->graph_exec && (m->graph_uid[1] == p->uid || m->graph_uid[2] == p->uid)) {
    cudaGraphExecDestroy(m->graph_exec);
    m->graph_exec = nullptr;
  }
  return IDDGCN_OK;
}

// drop the captured train step if it was captured on this graph (Graph.close calls this before destroying it)
int iddgcn_model_release_graph(iddgcn_model_t* m, const iddgcn_graph_t* g) {
  IDD_CHECK(m != nullptr, IDDGCN_EINVAL, "model is NULL");
  if (m->graph_exec && g && m->graph_uid[0] == g->uid) {
    cudaDeviceSynchronize();
    cudaGraphExecDestroy(m->graph_exec);
    m->graph_exec = nullptr;
  }
  return IDDGCN_OK;
}

// Per-group timing of iddgcn_model_train_step (not of the captured loop): enable, run a step, read.  ms_by_group has
// IDDGCN_PROFILE_GROUPS floats (order: ProfTag, see tools/bench_layer.py); reading synchronises the device.
int iddgcn_model_profile(iddgcn_model_t* m, int enable) {
  IDD_CHECK(m != nullptr, IDDGCN_EINVAL, "model is NULL");
  m->prof_on = enable != 0;
  m->prof_n = 0;
  return IDDGCN_OK;
}

int iddgcn_model_profile_read(iddgcn_model_t* m, float* ms_by_group, int n_groups) {
  IDD_CHECK(m && ms_by_group && n_groups >= (int)PT_COUNT, IDDGCN_EINVAL, "need room for %d groups", (int)PT_COUNT);
  IDD_CUDA(cudaDeviceSynchronize());
  for (int i = 0; i < n_groups; ++i) ms_by_group[i] = 0.f;
  for (size_t i = 1; i < m->prof_n; ++i) {
    float ms = 0.f;
    IDD_CUDA(cudaEventElapsedTime(&ms, m->prof_ev[i - 1], m->prof_ev[i]));
    ms_by_group[m->prof_tag[i]] += ms;
  }
  return IDDGCN_OK;
}

int iddgcn_model_forward(iddgcn_model_t* m, const iddgcn_graph_t* g, const iddgcn_plan_t* p, const float* vals,
                         int vals_mode, float* y, void* stream) {
  IDD_TRY(check_compat(m, g, p));
  IDD_CHECK(vals_mode >= 0 && vals_mode <= 2 && (vals_mode == 0 || vals || g->nnz == 0), IDDGCN_EINVAL,
            "bad vals / vals_mode");
  cudaStream_t st = (cudaStream_t)stream;
  TripleWS* w;
  IDD_TRY(ensure_triple_ws(m, p, &w));
  IDD_TRY(node_forward(m, g, vals, vals_mode, st));
  IDD_TRY(triple_forward(m, p, w, w->y, st));
  if (y && p->T) IDD_CUDA(cudaMemcpyAsync(y, w->y, (size_t)p->T * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return IDDGCN_OK;
}

// DLPack flavour of iddgcn_model_forward: vals_dl (NULL for the unweighted adjacency) [nnz], y_dl [T]
int iddgcn_model_forward_dl(iddgcn_model_t* m, const iddgcn_graph_t* g, const iddgcn_plan_t* p, void* vals_dl,
                            int vals_mode, void* y_dl, void* stream) {
  IDD_TRY(check_compat(m, g, p));
  float *vals = nullptr, *y = nullptr;
  if (vals_dl) IDD_TRY(dl_f32_shaped(vals_dl, "vals", m->device, g->nnz, "[nnz]", &vals));
  IDD_TRY(dl_f32_shaped(y_dl, "y", m->device, p->T, "[T]", &y));
  return iddgcn_model_forward(m, g, p, vals, vals_mode, y, stream);
}

int iddgcn_model_backward(iddgcn_model_t* m, const iddgcn_graph_t* g, const iddgcn_plan_t* p, const float* vals,
                          int vals_mode, const float* dy, int zero_grads, float* dvals, void* stream) {
  IDD_TRY(check_compat(m, g, p));
  IDD_CHECK(dy != nullptr || p->T == 0, IDDGCN_EINVAL, "dy is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  auto it = m->ws.find(p->T);
  IDD_CHECK(it != m->ws.end() && it->second.T == p->T && it->second.owner == p, IDDGCN_EINVAL,
            "model_backward must follow model_forward on the same plan");
  IDD_TRY(zero_step_buffers(m, zero_grads != 0, st));
  IDD_TRY(triple_backward(m, p, &it->second, it->second.y, dy, st));
  IDD_TRY(node_backward(m, g, vals, vals_mode, dvals, st));
  return IDDGCN_OK;
}

int iddgcn_model_train_step(iddgcn_model_t* m, const iddgcn_graph_t* g, const iddgcn_plan_t* pos,
                            const iddgcn_plan_t* neg, float lr, float beta1, float beta2, float eps,
                            int64_t step_index, int apply_update, float* loss_out, float* y_out, void* stream) {
  TripleWS* w;
  const iddgcn_plan* both;
  IDD_TRY(prepare_train(m, g, pos, neg, (cudaStream_t)stream, &both, &w));
  IDD_CHECK(step_index >= 1 || !apply_update, IDDGCN_EINVAL, "Adam step_index starts at 1");
  return train_step_body(m, g, both, pos->T, w, lr, beta1, beta2, eps, step_index, false, apply_update, loss_out,
                         y_out, (cudaStream_t)stream);
}

// num_steps training steps with the step captured ONCE into a CUDA graph and replayed (the shipped config is
// launch-bound: ~110 small kernels per step).  first_step_index = Adam t of the first step (>=1).
// loss_history: DEVICE float[num_steps] or NULL.  Synchronous w.r.t. `stream` ordering: work is enqueued on
// `stream` (or an internal stream when stream is the legacy default stream, which cannot be captured).
int iddgcn_model_train_loop(iddgcn_model_t* m, const iddgcn_graph_t* g, const iddgcn_plan_t* pos,
                            const iddgcn_plan_t* neg, float lr, float beta1, float beta2, float eps,
                            int64_t first_step_index, int64_t num_steps, float* loss_history, void* stream) {
  IDD_CHECK(first_step_index >= 1 && num_steps >= 0, IDDGCN_EINVAL, "bad step range");
  cudaStream_t st = (cudaStream_t)stream;
  bool own = false;
  if (st == nullptr || st == cudaStreamLegacy || st == cudaStreamPerThread) {
    if (!m->own_stream) IDD_CUDA(cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking));
    IDD_CUDA(cudaDeviceSynchronize());
    st = m->own_stream;
    own = true;
  }
  TripleWS* w;
  const iddgcn_plan* both;
  IDD_TRY(prepare_train(m, g, pos, neg, st, &both, &w));
  const int64_t t0 = first_step_index - 1;  // the graph increments the counter before Adam reads it
  IDD_CUDA(cudaMemcpyAsync(m->step_dev, &t0, sizeof(int64_t), cudaMemcpyHostToDevice, st));
  IDD_CUDA(cudaStreamSynchronize(st));
  const uint64_t neg_uid = neg ? neg->uid : 0;
  const bool same = m->graph_exec && m->graph_uid[0] == g->uid && m->graph_uid[1] == pos->uid &&
                    m->graph_uid[2] == neg_uid && m->graph_gen == m->ws_gen && m->graph_heavy_ws == g->heavy_ws &&
                    m->graph_hyper[0] == lr && m->graph_hyper[1] == beta1 && m->graph_hyper[2] == beta2 &&
                    m->graph_hyper[3] == eps;
  if (!same) {
    if (m->graph_exec) { cudaGraphExecDestroy(m->graph_exec); m->graph_exec = nullptr; }
    cudaGraph_t graph = nullptr;
    const bool prof = m->prof_on;
    m->prof_on = false;                       // the per-group events belong to the eager step only
    IDD_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    int rc = train_step_body(m, g, both, pos->T, w, lr, beta1, beta2, eps, 0, true, 1, nullptr, nullptr, st);
    cudaError_t ce = cudaStreamEndCapture(st, &graph);
    m->prof_on = prof;
    if (rc != IDDGCN_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    IDD_CUDA(ce);
    ce = cudaGraphInstantiate(&m->graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    IDD_CUDA(ce);
    m->graph_uid[0] = g->uid; m->graph_uid[1] = pos->uid; m->graph_uid[2] = neg_uid;
    m->graph_gen = m->ws_gen;
    m->graph_heavy_ws = g->heavy_ws;
    m->graph_hyper[0] = lr; m->graph_hyper[1] = beta1; m->graph_hyper[2] = beta2; m->graph_hyper[3] = eps;
  }
  for (int64_t i = 0; i < num_steps; ++i) {
    IDD_CUDA(cudaGraphLaunch(m->graph_exec, st));
    if (loss_history)
      IDD_CUDA(cudaMemcpyAsync(loss_history + i, m->loss_dev, sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  if (own) IDD_CUDA(cudaStreamSynchronize(st));
  return IDDGCN_OK;
}

int iddgcn_adam_step(float* theta, const float* grad, float* m, float* v, int64_t n, float lr, float beta1,
                     float beta2, float eps, int64_t step_index, void* stream) {
  IDD_CHECK(theta && grad && m && v, IDDGCN_EINVAL, "NULL buffer");
  IDD_CHECK(step_index >= 1, IDDGCN_EINVAL, "Adam step_index starts at 1");
  return adam_apply(theta, grad, m, v, n, lr, beta1, beta2, eps, step_index, nullptr, (cudaStream_t)stream);
}

// ---- layer-level entry points (used by the Python IDDGCN_Layer / DistMult autograd functions) ------------
int iddgcn_relation_transform(const float* S, const float* W, int R, int64_t N, int64_t D, float* Z, void* stream) {
  IDD_CHECK(S && W && Z, IDDGCN_EINVAL, "NULL buffer");
  return gemm_f32(false, false, N, D, D, S, D, N * D, W, D, D * D, Z, D, N * D, R, false, (cudaStream_t)stream);
}

int iddgcn_dense_transform(const float* A, const float* W, int64_t M, int64_t D, int trans_w, float* C,
                           int accumulate, int path, void* stream) {
  IDD_CHECK(A && W && C, IDDGCN_EINVAL, "NULL buffer");
  IDD_CHECK(M >= 0 && D > 0, IDDGCN_EINVAL, "bad shape M=%lld D=%lld", (long long)M, (long long)D);
  IDD_CHECK(path >= 0 && path <= 2, IDDGCN_EINVAL, "path must be 0 (automatic), 1 (fp32 SIMT) or 2 (tensor cores)");
  cudaStream_t st = (cudaStream_t)stream;
  if (M == 0) return IDDGCN_OK;
  if (path == 1) return gemm_f32_simt(false, trans_w != 0, M, D, D, A, D, 0, W, D, 0, C, D, 0, 1, accumulate != 0, st);
  if (path == 2) {
    IDD_CHECK(tc_gemm_eligible(false, M, D, D, A, D, W, D, C, D, 1), IDDGCN_EINVAL,
              "the tensor-core path needs D in {64, 128}, A and W 16-byte aligned and C 32-byte aligned (D=%lld)",
              (long long)D);
    return gemm_tf32x3(trans_w != 0, M, D, D, A, D, 0, W, 0, C, D, 0, 1, accumulate != 0, st);
  }
  return gemm_f32(false, trans_w != 0, M, D, D, A, D, 0, W, D, 0, C, D, 0, 1, accumulate != 0, st);
}

int iddgcn_dense_transform_grad(const float* A, const float* dC, int64_t M, int64_t D, float* dW, int accumulate,
                                int path, void* stream) {
  IDD_CHECK(A && dC && dW, IDDGCN_EINVAL, "NULL buffer");
  IDD_CHECK(M >= 0 && D > 0, IDDGCN_EINVAL, "bad shape M=%lld D=%lld", (long long)M, (long long)D);
  IDD_CHECK(path >= 0 && path <= 2, IDDGCN_EINVAL, "path must be 0 (automatic), 1 (fp32 SIMT) or 2 (tensor cores)");
  cudaStream_t st = (cudaStream_t)stream;
  DevBuf<float> ws;
  IDD_CUDA(ws.ensure(gemm_tn_workspace_elems(D, D, M, 1)));
  int rc;
  if (path == 1) {
    rc = gemm_tn_splitk_simt(D, D, M, A, D, 0, dC, D, 0, dW, D, 0, 1, accumulate != 0, ws.p, st);
  } else if (path == 2) {
    IDD_CHECK(tc_gemm_tn_eligible(D, D, M, A, D, dC, D, 1), IDDGCN_EINVAL,
              "the tensor-core path needs D in {64, 128}, M >= 1 and 16-byte aligned buffers (D=%lld)", (long long)D);
    rc = gemm_tn_tc(D, D, M, A, D, 0, dC, D, 0, dW, D, 0, 1, accumulate != 0, ws.p, st);
  } else {
    rc = gemm_tn_splitk_f32(D, D, M, A, D, 0, dC, D, 0, dW, D, 0, 1, accumulate != 0, ws.p, st);
  }
  IDD_CUDA(cudaStreamSynchronize(st));   // the workspace is released on return
  return rc;
}

int iddgcn_layer_fwd(const iddgcn_plan_t* p, const float* emb, const float* Hin, const float* Tin, const float* Z,
                     const float* W_self, const float* W_alpha, const float* b_alpha, int64_t D, float* Hout,
                     float* Tout, float* alpha_out, void* stream) {
  IDD_CHECK(p && Z && W_self && W_alpha && b_alpha && Hout && Tout, IDDGCN_EINVAL, "NULL buffer");
  IDD_CHECK((Hin && Tin) || emb, IDDGCN_EINVAL, "give Hin/Tin or the embedding table");
  IDD_CHECK(Tout == Hout + (size_t)p->T * D, IDDGCN_EINVAL,
            "Hout/Tout must be the two planes of one (2,T,D) buffer");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t T = p->T;
  if (T == 0) return IDDGCN_OK;
  const float* hin = Hin;
  DevBuf<float> tmp;
  if (!(Hin && Tin)) {
    IDD_CUDA(tmp.ensure((size_t)2 * T * D));
    IDD_TRY(gather_embeddings(p, emb, D, tmp.p, st));
    hin = tmp.p;
    IDD_TRY(gemm_f32(false, false, 2 * T, D, D, tmp.p, D, 0, W_self, D, 0, Hout, D, 0, 1, false, st));
  } else if (Tin == Hin + (size_t)T * D) {
    IDD_TRY(gemm_f32(false, false, 2 * T, D, D, Hin, D, 0, W_self, D, 0, Hout, D, 0, 1, false, st));
  } else {
    IDD_TRY(gemm_f32(false, false, T, D, D, Hin, D, 0, W_self, D, 0, Hout, D, 0, 1, false, st));
    IDD_TRY(gemm_f32(false, false, T, D, D, Tin, D, 0, W_self, D, 0, Tout, D, 0, 1, false, st));
  }
  IDD_TRY(layer_combine(p, hin, Z, W_alpha, b_alpha, D, Hout, alpha_out, st));
  if (tmp.p) IDD_CUDA(cudaStreamSynchronize(st));  // tmp is freed on return
  return IDDGCN_OK;
}

// Backward of iddgcn_layer_fwd (SURVEY App. A eq. 5-10).  HT_in / HT_out / dHT_out / dHT_in are (2,T,D) buffers
// (head plane then tail plane); dHT_out is overwritten with the pre-activation gradients.  Weight gradients are
// ACCUMULATED into dW_self [D,D], dW_alpha [D,R], db_alpha [R]; dZ [R,N,D] is accumulated when accumulate_dZ != 0,
// else overwritten.
int iddgcn_layer_bwd(const iddgcn_plan_t* p, const float* HT_in, const float* HT_out, const float* alpha,
                     const float* Z, const float* W_self, const float* W_alpha, int64_t D, float* dHT_out,
                     float* dHT_in, float* dZ, int accumulate_dZ, float* dW_self, float* dW_alpha, float* db_alpha,
                     void* stream) {
  IDD_CHECK(p && HT_in && HT_out && alpha && Z && W_self && W_alpha && dHT_out && dHT_in && dZ && dW_self &&
                dW_alpha && db_alpha, IDDGCN_EINVAL, "NULL buffer");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t T = p->T;
  const int R = p->R;
  DevBuf<float> dz, gws, sws;
  IDD_CUDA(dz.ensure((size_t)(T ? T : 1) * R));
  size_t gw = gemm_tn_workspace_elems(D, D, 2 * T, 1), g2 = gemm_tn_workspace_elems(D, R, T, 1);
  IDD_CUDA(gws.ensure(gw > g2 ? gw : g2));
  IDD_CUDA(sws.ensure(colsum_scratch_elems(T, R)));
  IDD_TRY(layer_bwd_pre(p, HT_out, alpha, Z, D, dHT_out, dz.p, st));
  IDD_TRY(gemm_tn_splitk_f32(D, D, 2 * T, HT_in, D, 0, dHT_out, D, 0, dW_self, D, 0, 1, true, gws.p, st));
  IDD_TRY(gemm_tn_splitk_f32(D, R, T, HT_in, D, 0, dz.p, R, 0, dW_alpha, R, 0, 1, true, gws.p, st));
  IDD_TRY(colsum(dz.p, T, R, db_alpha, 1, sws.p, st));
  IDD_TRY(segment_scatter(p, dHT_out, alpha, D, dZ, accumulate_dZ, st));
  IDD_TRY(gemm_f32(false, true, 2 * T, D, D, dHT_out, D, 0, W_self, D, 0, dHT_in, D, 0, 1, false, st));
  IDD_TRY(layer_bwd_post(dz.p, W_alpha, T, D, R, dHT_in, st));
  IDD_CUDA(cudaStreamSynchronize(st));  // temporaries are freed on return
  return IDDGCN_OK;
}

// DLPack flavour of the two layer entry points.  Every tensor is a DLManagedTensor* (float32, CUDA, compact):
//   HT_in / HT_out / dHT_out / dHT_in [2,T,D] (head plane, tail plane), Z / dZ [R,N,D], W_self / dW_self [D,D],
//   W_alpha / dW_alpha [D,R], b_alpha / db_alpha [R], alpha [T,R].  D is read from W_self.
int iddgcn_layer_fwd_dl(const iddgcn_plan_t* p, void* HT_in_dl, void* Z_dl, void* W_self_dl, void* W_alpha_dl,
                        void* b_alpha_dl, void* HT_out_dl, void* alpha_out_dl, void* stream) {
  IDD_CHECK(p != nullptr, IDDGCN_EINVAL, "plan is NULL");
  float *ws = nullptr, *hin, *z, *wa, *ba, *hout, *al;
  int64_t n = 0, D = 0;
  IDD_TRY(dl_f32_cuda(W_self_dl, "W_self", p->device, &ws, &n, &D));
  IDD_CHECK(n == D * D, IDDGCN_EINVAL, "W_self must be [D,D]");
  const int64_t T = p->T, R = p->R;
  IDD_TRY(dl_f32_shaped(HT_in_dl, "HT_in", p->device, 2 * T * D, "[2,T,D]", &hin));
  IDD_TRY(dl_f32_shaped(Z_dl, "Z", p->device, R * p->N * D, "[R,N,D]", &z));
  IDD_TRY(dl_f32_shaped(W_alpha_dl, "W_alpha", p->device, D * R, "[D,R]", &wa));
  IDD_TRY(dl_f32_shaped(b_alpha_dl, "b_alpha", p->device, R, "[R]", &ba));
  IDD_TRY(dl_f32_shaped(HT_out_dl, "HT_out", p->device, 2 * T * D, "[2,T,D]", &hout));
  IDD_TRY(dl_f32_shaped(alpha_out_dl, "alpha", p->device, T * R, "[T,R]", &al));
  return iddgcn_layer_fwd(p, nullptr, hin, hin + (size_t)T * D, z, ws, wa, ba, D, hout, hout + (size_t)T * D, al,
                          stream);
}

int iddgcn_layer_bwd_dl(const iddgcn_plan_t* p, void* HT_in_dl, void* HT_out_dl, void* alpha_dl, void* Z_dl,
                        void* W_self_dl, void* W_alpha_dl, void* dHT_out_dl, void* dHT_in_dl, void* dZ_dl,
                        int accumulate_dZ, void* dW_self_dl, void* dW_alpha_dl, void* db_alpha_dl, void* stream) {
  IDD_CHECK(p != nullptr, IDDGCN_EINVAL, "plan is NULL");
  float *ws = nullptr, *hin, *hout, *al, *z, *wa, *dho, *dhi, *dz, *dws, *dwa, *dba;
  int64_t n = 0, D = 0;
  IDD_TRY(dl_f32_cuda(W_self_dl, "W_self", p->device, &ws, &n, &D));
  IDD_CHECK(n == D * D, IDDGCN_EINVAL, "W_self must be [D,D]");
  const int64_t T = p->T, R = p->R;
  IDD_TRY(dl_f32_shaped(HT_in_dl, "HT_in", p->device, 2 * T * D, "[2,T,D]", &hin));
  IDD_TRY(dl_f32_shaped(HT_out_dl, "HT_out", p->device, 2 * T * D, "[2,T,D]", &hout));
  IDD_TRY(dl_f32_shaped(alpha_dl, "alpha", p->device, T * R, "[T,R]", &al));
  IDD_TRY(dl_f32_shaped(Z_dl, "Z", p->device, R * p->N * D, "[R,N,D]", &z));
  IDD_TRY(dl_f32_shaped(W_alpha_dl, "W_alpha", p->device, D * R, "[D,R]", &wa));
  IDD_TRY(dl_f32_shaped(dHT_out_dl, "dHT_out", p->device, 2 * T * D, "[2,T,D]", &dho));
  IDD_TRY(dl_f32_shaped(dHT_in_dl, "dHT_in", p->device, 2 * T * D, "[2,T,D]", &dhi));
  IDD_TRY(dl_f32_shaped(dZ_dl, "dZ", p->device, R * p->N * D, "[R,N,D]", &dz));
  IDD_TRY(dl_f32_shaped(dW_self_dl, "dW_self", p->device, D * D, "[D,D]", &dws));
  IDD_TRY(dl_f32_shaped(dW_alpha_dl, "dW_alpha", p->device, D * R, "[D,R]", &dwa));
  IDD_TRY(dl_f32_shaped(db_alpha_dl, "db_alpha", p->device, R, "[R]", &dba));
  return iddgcn_layer_bwd(p, hin, hout, al, z, ws, wa, D, dho, dhi, dz, accumulate_dZ, dws, dwa, dba, stream);
}

// out[plane][n][:] (+)= segment sums of dHT (2,T,D) by head and by tail; alpha NULL -> one plane, weight 1
int iddgcn_segment_scatter(const iddgcn_plan_t* p, const float* dHT, const float* alpha, int64_t D, float* out,
                           int accumulate, void* stream) {
  IDD_CHECK(p && dHT && out, IDDGCN_EINVAL, "NULL buffer");
  return segment_scatter(p, dHT, alpha, D, out, accumulate, (cudaStream_t)stream);
}

int iddgcn_distmult_fwd(const iddgcn_plan_t* p, const float* HT, const float* rel_emb, int64_t D, float* y,
                        void* stream) {
  IDD_CHECK(p && HT && rel_emb && y, IDDGCN_EINVAL, "NULL buffer");
  return distmult_fwd(p, HT, rel_emb, D, y, (cudaStream_t)stream);
}

// dHT (2,T,D) written; d_rel_emb [R,D] ACCUMULATED
int iddgcn_distmult_bwd(const iddgcn_plan_t* p, const float* HT, const float* rel_emb, const float* y,
                        const float* dy, int64_t D, float* dHT, float* d_rel_emb, void* stream) {
  IDD_CHECK(p && HT && rel_emb && y && dy && dHT && d_rel_emb, IDDGCN_EINVAL, "NULL buffer");
  cudaStream_t st = (cudaStream_t)stream;
  DevBuf<float> scratch;
  IDD_CUDA(scratch.ensure(distmult_scratch_elems(p, D)));
  IDD_TRY(distmult_bwd(p, HT, rel_emb, y, dy, D, dHT, d_rel_emb, scratch.p, st));
  IDD_CUDA(cudaStreamSynchronize(st));
  return IDDGCN_OK;
}

// dW[r] += S[r]ᵀ·dZ[r]  and  dS[r] (+)= dZ[r]·W[r]ᵀ  — backward of iddgcn_relation_transform
int iddgcn_relation_transform_bwd(const float* S, const float* W, const float* dZ, int R, int64_t N, int64_t D,
                                  float* dW, float* dS, int accumulate_dS, void* stream) {
  IDD_CHECK(S && W && dZ && dW && dS, IDDGCN_EINVAL, "NULL buffer");
  cudaStream_t st = (cudaStream_t)stream;
  DevBuf<float> gws;
  IDD_CUDA(gws.ensure(gemm_tn_workspace_elems(D, D, N, R)));
  IDD_TRY(gemm_tn_splitk_f32(D, D, N, S, D, N * D, dZ, D, N * D, dW, D, D * D, R, true, gws.p, st));
  IDD_TRY(gemm_f32(false, true, N, D, D, dZ, D, N * D, W, D, D * D, dS, D, N * D, R, accumulate_dS != 0, st));
  IDD_CUDA(cudaStreamSynchronize(st));
  return IDDGCN_OK;
}

// Keras BinaryCrossentropy over y = [T_pos positives | rest negatives]; loss_out, dy on device
int iddgcn_bce_loss_grad(const float* y, int64_t T_pos, int64_t T_total, float grad_scale, float* dy,
                         float* loss_out, void* stream) {
  IDD_CHECK(y && dy && loss_out, IDDGCN_EINVAL, "NULL buffer");
  cudaStream_t st = (cudaStream_t)stream;
  DevBuf<float> scratch;
  IDD_CUDA(scratch.ensure(bce_scratch_elems(T_total)));
  IDD_TRY(bce_loss_grad(y, T_pos, T_total, grad_scale, dy, loss_out, scratch.p, st));
  IDD_CUDA(cudaStreamSynchronize(st));
  return IDDGCN_OK;
}

}  // extern "C"
