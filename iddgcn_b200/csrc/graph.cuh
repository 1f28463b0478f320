// graph.cuh — the opaque graph / plan handles.
#pragma once
#include <vector>
#include "common.cuh"

constexpr int HEAVY_EDGES = 1024;

namespace iddgcn {
uint64_t next_handle_uid();   // core.cu: monotonically increasing ids (a recycled heap address never repeats an id)
}

struct iddgcn_graph {
  uint64_t uid = iddgcn::next_handle_uid();
  int device = 0;
  int64_t N = 0;        // source nodes (columns)
  int64_t n_rows = 0;   // destination rows held here (row_hi - row_lo)
  int64_t row_lo = 0;
  int R = 0;
  int64_t nnz = 0;
  std::vector<int64_t> nnz_rel;
  void* block_ptrs = nullptr;     // one allocation holding the row-pointer arrays below
  void* block_edges = nullptr;    // one allocation holding the six per-edge arrays below
  int32_t* csr_rowptr = nullptr;  // [R*n_rows+1]
  int32_t* csr_col = nullptr;     // [nnz]
  int32_t* csr_row = nullptr;     // [nnz]
  int32_t* csc_rowptr = nullptr;  // [R*N+1]
  int32_t* csc_row = nullptr;     // [nnz]
  int32_t* csc_perm = nullptr;    // [nnz]
  // node-major view of the CSC used by the backward kernel: all in-edges of source node k, relations in order,
  // are one contiguous run, so dX = Σ_r A_rᵀ·dS_r is a plain CSR gather over dS viewed as [(R·n_rows), D]
  int32_t* bwd_rowptr = nullptr;  // [N+1]
  int32_t* bwd_src = nullptr;     // [nnz]  r·n_rows + local head
  int32_t* bwd_perm = nullptr;    // [nnz]  CSR edge id
  // Hub rows (more than HEAVY_EDGES edges) are skipped by the row kernels and handled by a split pass: the row is cut
  // into segments of HEAVY_EDGES edges summed by separate warps, then combined in ascending segment order.
  struct HeavyList {
    int H = 0;                    // number of hub rows
    int total_segs = 0;
    int32_t* rows = nullptr;      // [H]   row ids, ascending
    int32_t* segptr = nullptr;    // [H+1] first segment of each hub row
  };
  HeavyList heavy_fwd, heavy_bwd;
  mutable float* heavy_ws = nullptr;   // [total_segs, D] partial sums: iddgcn_spmm_prepare, or grown on first use (pooled;
                                       // refused while the stream is being captured)
  mutable size_t heavy_ws_elems = 0;
  float* rel_sums_scratch = nullptr;
  // lazily created by iddgcn_spmm_fwd_bwd_host and kept for the next call (mutable cache)
  mutable float* host_ws = nullptr;
  mutable size_t host_ws_elems = 0;
  mutable cudaStream_t host_streams[3] = {nullptr, nullptr, nullptr};
  mutable cudaEvent_t host_events[16] = {};   // per column panel: X in, dS in, S ready, dX ready (up to 4 panels)
};

// Triples of one model call, plus the orderings that make every gather-backward a segment sum.
struct iddgcn_plan {
  uint64_t uid = iddgcn::next_handle_uid();
  int device = 0;
  int64_t T = 0;
  int64_t N = 0;
  int R = 0;
  void* block = nullptr;        // one allocation holding every array below
  int32_t* head = nullptr;      // [T]
  int32_t* rel = nullptr;       // [T]
  int32_t* tail = nullptr;      // [T]
  // triples sorted by head / by tail: perm gives the triple id at sorted position j,
  // ptr[n]..ptr[n+1] is the run of triples whose head (tail) is node n.
  int32_t* by_head_perm = nullptr;  // [T]
  int32_t* by_head_ptr = nullptr;   // [N+1]
  int32_t* by_tail_perm = nullptr;  // [T]
  int32_t* by_tail_ptr = nullptr;   // [N+1]
  int32_t* by_rel_perm = nullptr;   // [T]
  int32_t* by_rel_ptr = nullptr;    // [R+1]
  // node-major merge of the two orderings: entries by_node_ptr[n]..by_node_ptr[n+1] are the rows of a (2,T,D) buffer
  // viewed as [2T, D] that node n receives gradient from — its head-side triples (row k) in by-head order, then its
  // tail-side triples (row T + k) in by-tail order.  One run per node, so the gather backward is a flat stream.
  int32_t* by_node_ptr = nullptr;   // [N+1]
  int32_t* by_node_src = nullptr;   // [2T]
};

namespace iddgcn {
// new plan holding the triples of `a` followed by those of `b` (same N, R); caller owns it (plan_destroy_internal)
int plan_concat(const iddgcn_plan* a, const iddgcn_plan* b, cudaStream_t st, iddgcn_plan** out);
void plan_destroy_internal(iddgcn_plan* p);
}  // namespace iddgcn
