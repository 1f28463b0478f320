// graph.cuh — the opaque graph / plan handles.
#pragma once
#include <vector>
#include "common.cuh"

struct iddgcn_graph {
  int device = 0;
  int64_t N = 0;        // source nodes (columns)
  int64_t n_rows = 0;   // destination rows held here (row_hi - row_lo)
  int64_t row_lo = 0;
  int R = 0;
  int64_t nnz = 0;
  std::vector<int64_t> nnz_rel;
  int32_t* csr_rowptr = nullptr;  // [R*n_rows+1]
  int32_t* csr_col = nullptr;     // [nnz]
  int32_t* csr_row = nullptr;     // [nnz]
  int32_t* csc_rowptr = nullptr;  // [R*N+1]
  int32_t* csc_row = nullptr;     // [nnz]
  int32_t* csc_perm = nullptr;    // [nnz]
  float* rel_sums_scratch = nullptr;
};

// Triples of one model call, plus the orderings that make every gather-backward a segment sum.
struct iddgcn_plan {
  int device = 0;
  int64_t T = 0;
  int64_t N = 0;
  int R = 0;
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
};
