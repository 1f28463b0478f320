// explain.cu — explainer-side kernels: ragged 1-hop computation-graph extraction for a batch of triples
// (explanation/GnnExplainer.py:13-28, IDDGCN_explain.py:16-45) and the mask-optimisation losses
// (GnnExplainer.py:48; IDDGCN_explain.py:29-41,68-73).
//
// Batching: the reference explains test triples one by one.  Their computation graphs are independent, so B of
// them are laid out block-diagonally: triple b's sub-graph lives on node ids [b·N, (b+1)·N) of one big graph with
// B·N nodes, the embedding table is tiled, and ONE forward/backward of the ordinary kernels serves the whole batch.
// The sub-graph of (h, t) = every stored edge touching h or t = CSR rows (r, h), (r, t) and CSC rows (r, h),
// (r, t) of the full graph; duplicates are removed by the graph build that consumes the emitted triples.
#include "graph.cuh"
#include "primitives.cuh"

namespace iddgcn {

// one thread per (triple b, which in {head, tail}, relation r, direction in {out, in})
__global__ void subgraph_count_kernel(const int32_t* __restrict__ csr_rowptr, const int32_t* __restrict__ csc_rowptr,
                                      const int64_t* __restrict__ head, const int64_t* __restrict__ tail, int64_t B,
                                      int64_t N, int R, uint32_t* __restrict__ counts) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = B * 4 * R;
  if (i > total) return;
  if (i == total) { counts[i] = 0; return; }
  const int dir = (int)(i % 2);
  const int r = (int)((i / 2) % R);
  const int which = (int)((i / (2 * R)) % 2);
  const int64_t b = i / (4 * R);
  const int64_t node = which == 0 ? head[b] : tail[b];
  const int32_t* ptr = dir == 0 ? csr_rowptr : csc_rowptr;
  const int64_t s = (int64_t)r * N + node;
  counts[i] = (uint32_t)(ptr[s + 1] - ptr[s]);
}

__global__ void subgraph_fill_kernel(const int32_t* __restrict__ csr_rowptr, const int32_t* __restrict__ csr_col,
                                     const int32_t* __restrict__ csc_rowptr, const int32_t* __restrict__ csc_row,
                                     const int64_t* __restrict__ head, const int64_t* __restrict__ tail, int64_t B,
                                     int64_t N, int R, const uint32_t* __restrict__ offsets,
                                     int64_t* __restrict__ triples_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 4 * R) return;
  const int dir = (int)(i % 2);
  const int r = (int)((i / 2) % R);
  const int which = (int)((i / (2 * R)) % 2);
  const int64_t b = i / (4 * R);
  const int64_t node = which == 0 ? head[b] : tail[b];
  const int64_t s = (int64_t)r * N + node;
  const int64_t shift = b * N;
  int64_t o = offsets[i];
  if (dir == 0) {  // out-edges node -> col
    for (int e = csr_rowptr[s]; e < csr_rowptr[s + 1]; ++e, ++o) {
      triples_out[3 * o + 0] = node + shift;
      triples_out[3 * o + 1] = r;
      triples_out[3 * o + 2] = (int64_t)csr_col[e] + shift;
    }
  } else {         // in-edges row -> node
    for (int e = csc_rowptr[s]; e < csc_rowptr[s + 1]; ++e, ++o) {
      triples_out[3 * o + 0] = (int64_t)csc_row[e] + shift;
      triples_out[3 * o + 1] = r;
      triples_out[3 * o + 2] = node + shift;
    }
  }
}

// c[r][b] = Σ_{e in block b of relation r} a_e   (a_e = sigmoid(vals[e]) or vals[e] or 1)
// In the block-diagonal graph the edges of (relation r, block b) are the contiguous CSR range
// rowptr[r·NB + b·N] .. rowptr[r·NB + (b+1)·N].
__global__ void __launch_bounds__(128)
    block_relation_sums_kernel(const int32_t* __restrict__ rowptr, const float* __restrict__ vals, int mode,
                               int64_t N, int64_t B, float* __restrict__ out) {
  __shared__ float s[128];
  const int64_t b = blockIdx.x;
  const int r = blockIdx.y;
  const int64_t NB = N * B;
  const int beg = rowptr[(int64_t)r * NB + b * N], end = rowptr[(int64_t)r * NB + (b + 1) * N];
  float acc = 0.f;
  for (int e = beg + threadIdx.x; e < end; e += 128) {
    float a = 1.f;
    if (mode == IDDGCN_VALS_EXPLICIT) a = vals[e];
    else if (mode == IDDGCN_VALS_SIGMOID) a = sigmoidf_(vals[e]);
    acc += a;
  }
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[(int64_t)r * B + b] = s[0];
}

// Per-triple explainer objective.
//   pred_loss = -before · log(pred + 1e-5)                                   (GnnExplainer.py:48)
//   with structure term (targets != NULL):  loss = (pred_loss + mean_i (target_i - c_i/Σc)²) / 2   (IDDGCN_explain.py:29-41,68-73)
// Outputs: loss[b], dy[b] = ∂loss/∂pred, coef[r][b] = ∂loss/∂c_{r,b} (0 without the structure term).
__global__ void explain_loss_kernel(const float* __restrict__ before, const float* __restrict__ pred,
                                    const float* __restrict__ c, const float* __restrict__ targets, int R, int64_t B,
                                    float* __restrict__ loss, float* __restrict__ dy, float* __restrict__ coef) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float bp = before[b], p = pred[b];
  float l = -bp * logf(p + 1e-5f);
  float g = -bp / (p + 1e-5f);
  if (targets) {
    float tot = 0.f;
    for (int i = 0; i < R; ++i) tot += c[(int64_t)i * B + b];
    float sl = 0.f, dot = 0.f;
    for (int i = 0; i < R; ++i) {
      const float ratio = c[(int64_t)i * B + b] / tot;
      const float d = targets[i] - ratio;
      sl += d * d;
      dot += (-2.f * d / (float)R) * ratio;       // Σ_j (∂sl/∂ratio_j)·ratio_j
    }
    sl /= (float)R;
    for (int i = 0; i < R; ++i) {
      const float ratio = c[(int64_t)i * B + b] / tot;
      const float dsl_dratio = -2.f * (targets[i] - ratio) / (float)R;
      coef[(int64_t)i * B + b] = 0.5f * (dsl_dratio - dot) / tot;   // ∂ratio_j/∂c_i = (δ_ij - ratio_j)/tot
    }
    l = 0.5f * (l + sl);
    g *= 0.5f;
  } else if (coef) {
    for (int i = 0; i < R; ++i) coef[(int64_t)i * B + b] = 0.f;
  }
  loss[b] = l;
  dy[b] = g;
}

// dmask[e] += coef[rel_e][block_e] · ∂a_e/∂vals[e]   (structure-loss gradient, every edge of the sub-graph)
__global__ void structure_grad_kernel(const int32_t* __restrict__ rowptr, const float* __restrict__ vals, int mode,
                                      const float* __restrict__ coef, int64_t N, int64_t B, int R,
                                      float* __restrict__ dvals) {
  const int64_t b = blockIdx.x;
  const int r = blockIdx.y;
  const int64_t NB = N * B;
  const int beg = rowptr[(int64_t)r * NB + b * N], end = rowptr[(int64_t)r * NB + (b + 1) * N];
  const float cf = coef[(int64_t)r * B + b];
  for (int e = beg + threadIdx.x; e < end; e += blockDim.x) {
    float d = 1.f;
    if (mode == IDDGCN_VALS_SIGMOID) {
      const float sg = sigmoidf_(vals[e]);
      d = sg * (1.f - sg);
    }
    dvals[e] += cf * d;
  }
}

}  // namespace iddgcn

using namespace iddgcn;

extern "C" {

// Emits the block-diagonal union of the 1-hop computation graphs of B triples as int64 [M,3] (head, rel, tail)
// rows with node ids shifted by b·N (duplicates included; feed to iddgcn_graph_build with num_entities = B·N).
// Two-phase: pass triples_out = NULL to get *M_out, then call again with a buffer of M rows.
int iddgcn_subgraph_triples(const iddgcn_graph_t* g, const int64_t* head, const int64_t* tail, int64_t B,
                            int64_t* triples_out, int64_t capacity_rows, int64_t* M_out, void* stream) {
  IDD_CHECK(g && head && tail && M_out, IDDGCN_EINVAL, "NULL argument");
  IDD_CHECK(g->n_rows == g->N && g->row_lo == 0, IDDGCN_EINVAL, "needs the whole graph");
  IDD_CHECK(B >= 0 && B * 4 * g->R < (int64_t)0x7fffff00, IDDGCN_ELIMIT, "batch too large");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)B * 4 * g->R;
  if (n == 0) { *M_out = 0; return IDDGCN_OK; }
  DevBuf<uint32_t> counts, scratch;
  IDD_CUDA(counts.ensure(n + 1));
  IDD_CUDA(scratch.ensure(scan_scratch_elems(n + 1)));
  subgraph_count_kernel<<<(unsigned)((n + 1 + 255) / 256), 256, 0, st>>>(g->csr_rowptr, g->csc_rowptr, head, tail, B,
                                                                         g->N, g->R, counts.p);
  IDD_LAUNCH_CHECK();
  IDD_TRY(exclusive_scan_u32(counts.p, counts.p, n + 1, scratch.p, st));
  uint32_t total = 0;
  IDD_CUDA(cudaMemcpyAsync(&total, counts.p + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  IDD_CUDA(cudaStreamSynchronize(st));
  *M_out = (int64_t)total;
  if (!triples_out) return IDDGCN_OK;
  IDD_CHECK(capacity_rows >= (int64_t)total, IDDGCN_EINVAL, "triples_out holds %lld rows, %u needed",
            (long long)capacity_rows, total);
  subgraph_fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g->csr_rowptr, g->csr_col, g->csc_rowptr,
                                                                    g->csc_row, head, tail, B, g->N, g->R, counts.p,
                                                                    triples_out);
  IDD_LAUNCH_CHECK();
  IDD_CUDA(cudaStreamSynchronize(st));  // counts is freed on return
  return IDDGCN_OK;
}

// c[R][B]: per-(relation, block) sums of the edge values of a block-diagonal graph with blocks of N nodes
int iddgcn_block_relation_sums(const iddgcn_graph_t* g, const float* vals, int vals_mode, int64_t N, int64_t B,
                               float* out, void* stream) {
  IDD_CHECK(g && out, IDDGCN_EINVAL, "NULL argument");
  IDD_CHECK(N * B == g->N && g->n_rows == g->N, IDDGCN_EINVAL, "graph is not %lld blocks of %lld nodes", (long long)B,
            (long long)N);
  IDD_CHECK(vals_mode == IDDGCN_VALS_ONES || vals || g->nnz == 0, IDDGCN_EINVAL, "vals is NULL");
  IDD_CHECK(B < 0x7fffffff && g->R < 65536, IDDGCN_ELIMIT, "too many blocks");
  if (B == 0) return IDDGCN_OK;
  block_relation_sums_kernel<<<dim3((unsigned)B, g->R), 128, 0, (cudaStream_t)stream>>>(g->csr_rowptr, vals, vals_mode,
                                                                                       N, B, out);
  IDD_LAUNCH_CHECK();
  return IDDGCN_OK;
}

// loss[B], dy[B], coef[R][B] from before[B], pred[B] and (optionally) c[R][B] + targets[R] (device pointers)
int iddgcn_explain_loss(const float* before, const float* pred, const float* c, const float* targets, int R,
                        int64_t B, float* loss, float* dy, float* coef, void* stream) {
  IDD_CHECK(before && pred && loss && dy, IDDGCN_EINVAL, "NULL argument");
  IDD_CHECK(!targets || (c && coef), IDDGCN_EINVAL, "structure loss needs c and coef");
  if (B == 0) return IDDGCN_OK;
  explain_loss_kernel<<<(unsigned)((B + 127) / 128), 128, 0, (cudaStream_t)stream>>>(before, pred, c, targets, R, B,
                                                                                    loss, dy, coef);
  IDD_LAUNCH_CHECK();
  return IDDGCN_OK;
}

int iddgcn_structure_grad(const iddgcn_graph_t* g, const float* vals, int vals_mode, const float* coef, int64_t N,
                          int64_t B, float* dvals, void* stream) {
  IDD_CHECK(g && coef && dvals, IDDGCN_EINVAL, "NULL argument");
  IDD_CHECK(N * B == g->N && g->n_rows == g->N, IDDGCN_EINVAL, "graph is not %lld blocks of %lld nodes", (long long)B,
            (long long)N);
  if (B == 0) return IDDGCN_OK;
  structure_grad_kernel<<<dim3((unsigned)B, g->R), 64, 0, (cudaStream_t)stream>>>(g->csr_rowptr, vals, vals_mode, coef,
                                                                                 N, B, g->R, dvals);
  IDD_LAUNCH_CHECK();
  return IDDGCN_OK;
}

}  // extern "C"
