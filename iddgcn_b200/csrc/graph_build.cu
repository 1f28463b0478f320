// graph_build.cu — triples -> relation-segmented CSR + CSC, entirely on device.
//
// Replaces utils1.get_adj_mats / distinct (prediction/utils1.py:415-451): per relation keep
// (head, tail) of rows with rel == i, np.unique(axis=0) (lexicographic sort + dedup).  Here all
// relations are handled at once by sorting packed keys (rel | head | tail); the result is
// bit-exact w.r.t. the reference (integer work, deterministic sort).
#include "graph.cuh"
#include "primitives.cuh"

namespace iddgcn {

struct KeyLayout {
  int bt;  // bits of the low field
  int bh;  // bits of the middle field
  int br;  // bits of rel
  __host__ __device__ int total() const { return bt + bh + br; }
};

// key = rel << (bh+bt) | (head-row_lo) << bt | tail ; invalid rows get bit `total` set.
__global__ void pack_keys_kernel(const int64_t* __restrict__ triples, size_t M, int64_t N, int R, int64_t row_lo,
                                 int64_t row_hi, KeyLayout L, uint64_t* __restrict__ keys, int* __restrict__ err) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  const int64_t h = triples[3 * i + 0];
  const int64_t r = triples[3 * i + 1];
  const int64_t t = triples[3 * i + 2];
  const uint64_t invalid = 1ull << L.total();
  uint64_t key = invalid;
  if (r >= 0 && r < (int64_t)R) {
    if (h < 0 || h >= N || t < 0 || t >= N) {
      *err = 1;  // benign race: every writer stores 1
    } else if (h >= row_lo && h < row_hi) {
      key = ((uint64_t)r << (L.bh + L.bt)) | ((uint64_t)(h - row_lo) << L.bt) | (uint64_t)t;
    }
  }
  keys[i] = key;
}

__global__ void unique_flag_kernel(const uint64_t* __restrict__ keys, size_t M, uint64_t invalid,
                                   uint32_t* __restrict__ flags) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > M) return;
  uint32_t f = 0;
  if (i < M) {
    const uint64_t k = keys[i];
    f = (k < invalid) && (i == 0 || keys[i - 1] != k);
  }
  flags[i] = f;  // flags[M] = 0 so that the scan yields the total at [M]
}

__global__ void unique_compact_kernel(const uint64_t* __restrict__ keys, size_t M, uint64_t invalid,
                                      const uint32_t* __restrict__ pos, uint64_t* __restrict__ ukeys) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  const uint64_t k = keys[i];
  if ((k < invalid) && (i == 0 || keys[i - 1] != k)) ukeys[pos[i]] = k;
}

// CSR edge arrays from the unique sorted keys, and the (rel | tail | head) keys for the CSC sort.
__global__ void decode_csr_kernel(const uint64_t* __restrict__ ukeys, size_t nnz, KeyLayout L,
                                  int32_t* __restrict__ col, int32_t* __restrict__ row,
                                  uint64_t* __restrict__ keys2, uint32_t* __restrict__ eid) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  const uint64_t k = ukeys[e];
  const uint64_t t = k & ((1ull << L.bt) - 1ull);
  const uint64_t h = (k >> L.bt) & ((1ull << L.bh) - 1ull);
  const uint64_t r = k >> (L.bt + L.bh);
  col[e] = (int32_t)t;
  row[e] = (int32_t)h;
  keys2[e] = (r << (L.bh + L.bt)) | (t << L.bh) | h;
  eid[e] = (uint32_t)e;
}

__global__ void decode_csc_kernel(const uint64_t* __restrict__ keys2, const uint32_t* __restrict__ eid, size_t nnz,
                                  KeyLayout L, int32_t* __restrict__ csc_row, int32_t* __restrict__ csc_perm) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  csc_row[e] = (int32_t)(keys2[e] & ((1ull << L.bh) - 1ull));
  csc_perm[e] = (int32_t)eid[e];
}

// ptr[s] = lower_bound(keys, (s / seg_len) << hi_shift | (s % seg_len) << lo_shift); ptr[num_segs] = nnz
__global__ void segment_ptr_kernel(const uint64_t* __restrict__ keys, size_t nnz, size_t num_segs, uint64_t seg_len,
                                   int hi_shift, int lo_shift, int32_t* __restrict__ ptr) {
  const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s > num_segs) return;
  if (s == num_segs) {
    ptr[s] = (int32_t)nnz;
    return;
  }
  const uint64_t target = ((uint64_t)(s / seg_len) << hi_shift) | ((uint64_t)(s % seg_len) << lo_shift);
  size_t lo = 0, hi = nnz;
  while (lo < hi) {
    const size_t mid = (lo + hi) >> 1;
    if (keys[mid] < target) lo = mid + 1; else hi = mid;
  }
  ptr[s] = (int32_t)lo;
}

// in-degree of node k summed over relations (+ a trailing 0 so the scan yields the total)
__global__ void bwd_count_kernel(const int32_t* __restrict__ csc_rowptr, size_t N, int R, uint32_t* __restrict__ cnt) {
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k > N) return;
  uint32_t c = 0;
  if (k < N)
    for (int r = 0; r < R; ++r) c += (uint32_t)(csc_rowptr[(size_t)r * N + k + 1] - csc_rowptr[(size_t)r * N + k]);
  cnt[k] = c;
}

// thread per (relation r, node k): copy its CSC run to its place in the node-major arrays
__global__ void bwd_fill_kernel(const int32_t* __restrict__ csc_rowptr, const int32_t* __restrict__ csc_row,
                                const int32_t* __restrict__ csc_perm, size_t N, int R, int64_t n_rows,
                                const int32_t* __restrict__ bwd_rowptr, int32_t* __restrict__ bwd_src,
                                int32_t* __restrict__ bwd_perm) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * (size_t)R) return;
  const size_t k = i / R;
  const int r = (int)(i % R);
  int dst = bwd_rowptr[k];
  for (int q = 0; q < r; ++q) dst += csc_rowptr[(size_t)q * N + k + 1] - csc_rowptr[(size_t)q * N + k];
  const int beg = csc_rowptr[(size_t)r * N + k], end = csc_rowptr[(size_t)r * N + k + 1];
  for (int e = beg; e < end; ++e, ++dst) {
    bwd_src[dst] = (int32_t)((int64_t)r * n_rows + csc_row[e]);
    bwd_perm[dst] = csc_perm[e];
  }
}

// hub-row detection: flag[i] = number of HEAVY_EDGES-segments of row i if it is a hub row, else 0; hub[i] = 0/1
__global__ void heavy_flag_kernel(const int32_t* __restrict__ ptr, size_t rows, uint32_t* __restrict__ is_hub,
                                  uint32_t* __restrict__ nseg) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > rows) return;
  uint32_t h = 0, s = 0;
  if (i < rows) {
    const int deg = ptr[i + 1] - ptr[i];
    if (deg > HEAVY_EDGES) { h = 1; s = (uint32_t)((deg + HEAVY_EDGES - 1) / HEAVY_EDGES); }
  }
  is_hub[i] = h;
  nseg[i] = s;
}

__global__ void heavy_compact_kernel(const int32_t* __restrict__ ptr, size_t rows, const uint32_t* __restrict__ hub_pos,
                                     const uint32_t* __restrict__ seg_pos, int32_t* __restrict__ out_rows,
                                     int32_t* __restrict__ out_segptr, uint32_t H) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > rows) return;
  if (i == rows) { out_segptr[H] = (int32_t)seg_pos[rows]; return; }
  if (ptr[i + 1] - ptr[i] > HEAVY_EDGES) {
    out_rows[hub_pos[i]] = (int32_t)i;
    out_segptr[hub_pos[i]] = (int32_t)seg_pos[i];
  }
}

static inline unsigned blocks_for(size_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

static inline unsigned blocks_for(size_t n, int threads);

static int build_heavy_list(const int32_t* ptr, size_t rows, uint32_t* hub, uint32_t* seg, uint32_t* scr,
                            cudaStream_t st, iddgcn_graph::HeavyList* out) {
  heavy_flag_kernel<<<blocks_for(rows + 1, 256), 256, 0, st>>>(ptr, rows, hub, seg);
  IDD_LAUNCH_CHECK();
  IDD_TRY(exclusive_scan_u32(hub, hub, rows + 1, scr, st));
  IDD_TRY(exclusive_scan_u32(seg, seg, rows + 1, scr, st));
  uint32_t H = 0, S = 0;
  IDD_CUDA(cudaMemcpyAsync(&H, hub + rows, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  IDD_CUDA(cudaMemcpyAsync(&S, seg + rows, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  IDD_CUDA(cudaStreamSynchronize(st));
  out->H = (int)H;
  out->total_segs = (int)S;
  if (H) {
    IDD_CUDA(pool_alloc((void**)&out->rows, (size_t)H * sizeof(int32_t)));
    IDD_CUDA(pool_alloc((void**)&out->segptr, ((size_t)H + 1) * sizeof(int32_t)));
    heavy_compact_kernel<<<blocks_for(rows + 1, 256), 256, 0, st>>>(ptr, rows, hub, seg, out->rows, out->segptr, H);
    IDD_LAUNCH_CHECK();
  }
  return IDDGCN_OK;
}

static int graph_build_impl(const int64_t* triples_in, int64_t M, int on_device, int64_t N, int R, int64_t row_lo,
                            int64_t row_hi, cudaStream_t st, iddgcn_graph* g) {
  const int64_t n_rows = row_hi - row_lo;
  KeyLayout L;
  L.bt = ceil_log2_u64((uint64_t)(N - 1));
  L.bh = n_rows > 0 ? ceil_log2_u64((uint64_t)(n_rows - 1)) : 1;   // an empty row block keeps no triple at all
  L.br = ceil_log2_u64((uint64_t)(R - 1));
  if (L.bt < 1) L.bt = 1;
  if (L.bh < 1) L.bh = 1;
  if (L.br < 1) L.br = 1;
  IDD_CHECK(L.total() + 1 <= 63, IDDGCN_ELIMIT, "key does not fit 63 bits (N=%lld, R=%d)", (long long)N, R);
  const size_t csr_segs = (size_t)R * (size_t)n_rows;
  const size_t csc_segs = (size_t)R * (size_t)N;
  const size_t Nz = (size_t)N;
  IDD_CHECK(M < (int64_t)0x7fffff00 && csc_segs < (size_t)0x7fffff00 && csr_segs < (size_t)0x7fffff00, IDDGCN_ELIMIT,
            "graph exceeds int32 index range (M=%lld, R*N=%zu)", (long long)M, csc_segs);

  g->N = N; g->n_rows = n_rows; g->row_lo = row_lo; g->R = R;
  g->nnz_rel.assign(R, 0);
  IDD_CUDA(cudaGetDevice(&g->device));

  // persistent block 1: the three row-pointer arrays + the per-relation scratch
  {
    const size_t bytes = Arena::pad((csr_segs + 1) * 4) + Arena::pad((csc_segs + 1) * 4) + Arena::pad((Nz + 1) * 4) +
                         Arena::pad((size_t)(R > 0 ? R : 1) * 4);
    IDD_CUDA(pool_alloc(&g->block_ptrs, bytes));
    Arena a{reinterpret_cast<char*>(g->block_ptrs), bytes, 0};
    g->csr_rowptr = a.get<int32_t>(csr_segs + 1);
    g->csc_rowptr = a.get<int32_t>(csc_segs + 1);
    g->bwd_rowptr = a.get<int32_t>(Nz + 1);
    g->rel_sums_scratch = a.get<float>((size_t)(R > 0 ? R : 1));
  }

  // temporary arena
  const size_t Mz = (size_t)M;
  const size_t Ma = Mz ? Mz : 1;
  const RadixSortPlan plan = radix_sort_plan(Ma);
  const size_t max_rows = (csr_segs > Nz ? csr_segs : Nz);
  size_t scan_n = Mz + 1;
  if (max_rows + 1 > scan_n) scan_n = max_rows + 1;
  size_t scr_elems = scan_scratch_elems(scan_n);
  if (plan.scratch_elems > scr_elems) scr_elems = plan.scratch_elems;
  size_t tmp_bytes = 2 * Arena::pad(Ma * 8) + 2 * Arena::pad(Ma * 4) + Arena::pad(plan.hist_elems * 4) +
                     Arena::pad(scr_elems * 4) + Arena::pad((Mz + 1) * 4) + 2 * Arena::pad((max_rows + 1) * 4) +
                     Arena::pad(256);
  if (!on_device && Mz) tmp_bytes += Arena::pad(Mz * 3 * 8);
  DevBuf<char> tmp;
  IDD_CUDA(tmp.ensure(tmp_bytes));
  Arena ar{tmp.p, tmp_bytes, 0};
  uint64_t* kA = ar.get<uint64_t>(Ma);
  uint64_t* kB = ar.get<uint64_t>(Ma);
  uint32_t* vA = ar.get<uint32_t>(Ma);
  uint32_t* vB = ar.get<uint32_t>(Ma);
  uint32_t* hist = ar.get<uint32_t>(plan.hist_elems);
  uint32_t* scratch = ar.get<uint32_t>(scr_elems);
  uint32_t* flags = ar.get<uint32_t>(Mz + 1);
  uint32_t* hub_tmp = ar.get<uint32_t>(max_rows + 1);
  uint32_t* seg_tmp = ar.get<uint32_t>(max_rows + 1);
  int* err = ar.get<int>(64);
  const int64_t* triples = triples_in;
  if (!on_device && Mz) {
    int64_t* trip_tmp = ar.get<int64_t>(Mz * 3);
    IDD_CHECK(trip_tmp != nullptr, IDDGCN_ENOMEM, "internal: arena too small");
    IDD_CUDA(cudaMemcpyAsync(trip_tmp, triples_in, Mz * 3 * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    triples = trip_tmp;
  }
  IDD_CHECK(err != nullptr && ar.off <= ar.cap, IDDGCN_ENOMEM, "internal: arena too small");

  size_t nnz = 0;
  uint64_t* ukeys = nullptr;
  uint64_t* sorted = nullptr;
  if (Mz) {
    IDD_CUDA(cudaMemsetAsync(err, 0, sizeof(int), st));
    pack_keys_kernel<<<blocks_for(Mz, 256), 256, 0, st>>>(triples, Mz, N, R, row_lo, row_hi, L, kA, err);
    IDD_LAUNCH_CHECK();
    IDD_TRY(radix_sort_u64(kA, kB, nullptr, nullptr, Mz, L.total() + 1, plan, hist, scratch, st, &sorted, nullptr));
    uint64_t* other = (sorted == kA) ? kB : kA;
    const uint64_t invalid = 1ull << L.total();
    unique_flag_kernel<<<blocks_for(Mz + 1, 256), 256, 0, st>>>(sorted, Mz, invalid, flags);
    IDD_LAUNCH_CHECK();
    IDD_TRY(exclusive_scan_u32(flags, flags, Mz + 1, scratch, st));
    unique_compact_kernel<<<blocks_for(Mz, 256), 256, 0, st>>>(sorted, Mz, invalid, flags, other);
    IDD_LAUNCH_CHECK();
    uint32_t nnz32 = 0;
    int herr = 0;
    IDD_CUDA(cudaMemcpyAsync(&nnz32, flags + Mz, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    IDD_CUDA(cudaMemcpyAsync(&herr, err, sizeof(int), cudaMemcpyDeviceToHost, st));
    IDD_CUDA(cudaStreamSynchronize(st));
    IDD_CHECK(herr == 0, IDDGCN_ERANGE, "triples contain an entity index outside [0, %lld)", (long long)N);
    nnz = nnz32;
    ukeys = other;  // `sorted` is now free to be reused
  }

  // persistent block 2: the six per-edge arrays
  {
    const size_t ne = nnz ? nnz : 1;
    const size_t bytes = 6 * Arena::pad(ne * 4);
    IDD_CUDA(pool_alloc(&g->block_edges, bytes));
    Arena a{reinterpret_cast<char*>(g->block_edges), bytes, 0};
    g->csr_col = a.get<int32_t>(ne);
    g->csr_row = a.get<int32_t>(ne);
    g->csc_row = a.get<int32_t>(ne);
    g->csc_perm = a.get<int32_t>(ne);
    g->bwd_src = a.get<int32_t>(ne);
    g->bwd_perm = a.get<int32_t>(ne);
  }

  if (nnz) {
    segment_ptr_kernel<<<blocks_for(csr_segs + 1, 256), 256, 0, st>>>(ukeys, nnz, csr_segs, (uint64_t)n_rows,
                                                                      L.bh + L.bt, L.bt, g->csr_rowptr);
    IDD_LAUNCH_CHECK();
    uint64_t* keys2 = sorted;
    decode_csr_kernel<<<blocks_for(nnz, 256), 256, 0, st>>>(ukeys, nnz, L, g->csr_col, g->csr_row, keys2, vA);
    IDD_LAUNCH_CHECK();
    const RadixSortPlan plan2 = radix_sort_plan(nnz);  // nnz <= M: hist/scratch are large enough
    uint64_t* s2 = nullptr;
    uint32_t* sv = nullptr;
    IDD_TRY(radix_sort_u64(keys2, ukeys, vA, vB, nnz, L.total(), plan2, hist, scratch, st, &s2, &sv));
    decode_csc_kernel<<<blocks_for(nnz, 256), 256, 0, st>>>(s2, sv, nnz, L, g->csc_row, g->csc_perm);
    IDD_LAUNCH_CHECK();
    segment_ptr_kernel<<<blocks_for(csc_segs + 1, 256), 256, 0, st>>>(s2, nnz, csc_segs, (uint64_t)N, L.bh + L.bt, L.bh,
                                                                      g->csc_rowptr);
    IDD_LAUNCH_CHECK();
  } else {
    IDD_CUDA(cudaMemsetAsync(g->csr_rowptr, 0, (csr_segs + 1) * sizeof(int32_t), st));
    IDD_CUDA(cudaMemsetAsync(g->csc_rowptr, 0, (csc_segs + 1) * sizeof(int32_t), st));
  }
  // node-major backward view
  {
    uint32_t* cnt = reinterpret_cast<uint32_t*>(g->bwd_rowptr);
    bwd_count_kernel<<<blocks_for(Nz + 1, 256), 256, 0, st>>>(g->csc_rowptr, Nz, R, cnt);
    IDD_LAUNCH_CHECK();
    IDD_TRY(exclusive_scan_u32(cnt, cnt, Nz + 1, scratch, st));
    if (nnz) {
      bwd_fill_kernel<<<blocks_for(Nz * (size_t)R, 256), 256, 0, st>>>(g->csc_rowptr, g->csc_row, g->csc_perm, Nz, R,
                                                                       n_rows, g->bwd_rowptr, g->bwd_src, g->bwd_perm);
      IDD_LAUNCH_CHECK();
    }
  }
  IDD_TRY(build_heavy_list(g->csr_rowptr, csr_segs, hub_tmp, seg_tmp, scratch, st, &g->heavy_fwd));
  IDD_TRY(build_heavy_list(g->bwd_rowptr, Nz, hub_tmp, seg_tmp, scratch, st, &g->heavy_bwd));
  g->nnz = (int64_t)nnz;
  // per-relation counts: R+1 strided reads of the CSR row pointer
  std::vector<int32_t> bounds(R + 1, 0);
  for (int r = 0; r <= R; ++r)
    IDD_CUDA(cudaMemcpyAsync(&bounds[r], g->csr_rowptr + (size_t)r * n_rows, sizeof(int32_t), cudaMemcpyDeviceToHost,
                             st));
  IDD_CUDA(cudaStreamSynchronize(st));  // also: every kernel using the arena has finished before it is freed
  for (int r = 0; r < R; ++r) g->nnz_rel[r] = (int64_t)bounds[r + 1] - bounds[r];
  return IDDGCN_OK;
}

static void graph_free(iddgcn_graph* g) {
  if (!g) return;
  pool_free(g->block_ptrs);    // csr_rowptr, csc_rowptr, bwd_rowptr, rel_sums_scratch are views into it
  pool_free(g->block_edges);   // csr_col, csr_row, csc_row, csc_perm, bwd_src, bwd_perm likewise
  pool_free(g->heavy_fwd.rows); pool_free(g->heavy_fwd.segptr);
  pool_free(g->heavy_bwd.rows); pool_free(g->heavy_bwd.segptr);
  pool_free(g->heavy_ws);
  pool_free(g->host_ws);
  for (int i = 0; i < 3; ++i)
    if (g->host_streams[i]) cudaStreamDestroy(g->host_streams[i]);
  for (int i = 0; i < 16; ++i)
    if (g->host_events[i]) cudaEventDestroy(g->host_events[i]);
  delete g;
}

// ---------------------------------------------------------------------------
// plan: int32 copies of (head, rel, tail) + sort-by-head / by-tail / by-rel permutations
// ---------------------------------------------------------------------------
__global__ void plan_pack_kernel(const int64_t* __restrict__ head, const int64_t* __restrict__ rel,
                                 const int64_t* __restrict__ tail, size_t T, int64_t N, int R,
                                 int32_t* __restrict__ h32, int32_t* __restrict__ r32, int32_t* __restrict__ t32,
                                 int* __restrict__ err) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T) return;
  const int64_t h = head[i], r = rel[i], t = tail[i];
  if (h < 0 || h >= N || t < 0 || t >= N || r < 0 || r >= R) *err = 1;
  h32[i] = (int32_t)h; r32[i] = (int32_t)r; t32[i] = (int32_t)t;
}

__global__ void plan_keys_kernel(const int32_t* __restrict__ field, size_t T, uint64_t* __restrict__ keys,
                                 uint32_t* __restrict__ ids) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T) return;
  keys[i] = (uint64_t)(uint32_t)field[i];
  ids[i] = (uint32_t)i;
}

__global__ void copy_u32_to_i32_kernel(const uint32_t* __restrict__ in, size_t n, int32_t* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int32_t)in[i];
}

static int plan_sort_by(const int32_t* field, size_t T, size_t num_bins, int bits, uint64_t* kA, uint64_t* kB,
                        uint32_t* vA, uint32_t* vB, const RadixSortPlan& plan, uint32_t* hist, uint32_t* scratch,
                        cudaStream_t st, int32_t* perm_out, int32_t* ptr_out) {
  plan_keys_kernel<<<blocks_for(T, 256), 256, 0, st>>>(field, T, kA, vA);
  IDD_LAUNCH_CHECK();
  uint64_t* ks = nullptr;
  uint32_t* vs = nullptr;
  IDD_TRY(radix_sort_u64(kA, kB, vA, vB, T, bits, plan, hist, scratch, st, &ks, &vs));
  copy_u32_to_i32_kernel<<<blocks_for(T, 256), 256, 0, st>>>(vs, T, perm_out);
  IDD_LAUNCH_CHECK();
  // key == bin index: seg_len = num_bins+1 keeps s / seg_len == 0 for every s <= num_bins
  segment_ptr_kernel<<<blocks_for(num_bins + 1, 256), 256, 0, st>>>(ks, T, num_bins, (uint64_t)num_bins + 1, 40, 0,
                                                                    ptr_out);
  IDD_LAUNCH_CHECK();
  return IDDGCN_OK;
}

// merged node-major list (see graph.cuh): one warp per node copies its two runs
__global__ void plan_node_ptr_kernel(const int32_t* __restrict__ hptr, const int32_t* __restrict__ tptr, size_t N,
                                     int32_t* __restrict__ nptr) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= N) nptr[i] = hptr[i] + tptr[i];
}

__global__ void plan_node_fill_kernel(const int32_t* __restrict__ hptr, const int32_t* __restrict__ hperm,
                                      const int32_t* __restrict__ tptr, const int32_t* __restrict__ tperm, size_t N,
                                      int32_t T, const int32_t* __restrict__ nptr, int32_t* __restrict__ nsrc) {
  const size_t n = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  const int hb = hptr[n], he = hptr[n + 1], tb = tptr[n], te = tptr[n + 1];
  int32_t* dst = nsrc + nptr[n];
  for (int j = hb + lane; j < he; j += 32) dst[j - hb] = hperm[j];
  dst += he - hb;
  for (int j = tb + lane; j < te; j += 32) dst[j - tb] = T + tperm[j];
}

static int plan_create_impl(const int64_t* head, const int64_t* rel, const int64_t* tail, int64_t T, int on_device,
                            int64_t N, int R, cudaStream_t st, iddgcn_plan* p) {
  IDD_CUDA(cudaGetDevice(&p->device));
  p->T = T; p->N = N; p->R = R;
  const size_t Tz = (size_t)T;
  const size_t Ta = Tz ? Tz : 1;
  {  // one persistent allocation for the nine arrays
    const size_t bytes = 6 * Arena::pad(Ta * 4) + 3 * Arena::pad(((size_t)N + 1) * 4) + Arena::pad(((size_t)R + 1) * 4) +
                         Arena::pad(2 * Ta * 4);
    IDD_CUDA(pool_alloc(&p->block, bytes));
    Arena a{reinterpret_cast<char*>(p->block), bytes, 0};
    p->head = a.get<int32_t>(Ta); p->rel = a.get<int32_t>(Ta); p->tail = a.get<int32_t>(Ta);
    p->by_head_perm = a.get<int32_t>(Ta); p->by_tail_perm = a.get<int32_t>(Ta); p->by_rel_perm = a.get<int32_t>(Ta);
    p->by_head_ptr = a.get<int32_t>((size_t)N + 1); p->by_tail_ptr = a.get<int32_t>((size_t)N + 1);
    p->by_rel_ptr = a.get<int32_t>((size_t)R + 1);
    p->by_node_ptr = a.get<int32_t>((size_t)N + 1);
    p->by_node_src = a.get<int32_t>(2 * Ta);
    IDD_CHECK(p->by_node_src != nullptr, IDDGCN_ENOMEM, "internal: plan block too small");
  }
  IDD_CHECK(2 * T < (int64_t)0x7fffff00, IDDGCN_ELIMIT, "plan exceeds the int32 index range (T=%lld)", (long long)T);
  if (!Tz) {
    IDD_CUDA(cudaMemsetAsync(p->by_head_ptr, 0, ((size_t)N + 1) * 4, st));
    IDD_CUDA(cudaMemsetAsync(p->by_tail_ptr, 0, ((size_t)N + 1) * 4, st));
    IDD_CUDA(cudaMemsetAsync(p->by_rel_ptr, 0, ((size_t)R + 1) * 4, st));
    IDD_CUDA(cudaMemsetAsync(p->by_node_ptr, 0, ((size_t)N + 1) * 4, st));
    IDD_CUDA(cudaStreamSynchronize(st));
    return IDDGCN_OK;
  }
  const RadixSortPlan plan = radix_sort_plan(Tz);
  size_t tmp_bytes = 2 * Arena::pad(Tz * 8) + 2 * Arena::pad(Tz * 4) + Arena::pad(plan.hist_elems * 4) +
                     Arena::pad(plan.scratch_elems * 4) + Arena::pad(256);
  if (!on_device) tmp_bytes += Arena::pad(Tz * 3 * 8);
  DevBuf<char> tmp;
  IDD_CUDA(tmp.ensure(tmp_bytes));
  Arena ar{tmp.p, tmp_bytes, 0};
  uint64_t* kA = ar.get<uint64_t>(Tz);
  uint64_t* kB = ar.get<uint64_t>(Tz);
  uint32_t* vA = ar.get<uint32_t>(Tz);
  uint32_t* vB = ar.get<uint32_t>(Tz);
  uint32_t* hist = ar.get<uint32_t>(plan.hist_elems);
  uint32_t* scratch = ar.get<uint32_t>(plan.scratch_elems);
  int* err = ar.get<int>(64);
  const int64_t *dh = head, *dr = rel, *dt = tail;
  if (!on_device) {
    int64_t* t3 = ar.get<int64_t>(Tz * 3);
    IDD_CHECK(t3 != nullptr, IDDGCN_ENOMEM, "internal: arena too small");
    IDD_CUDA(cudaMemcpyAsync(t3, head, Tz * 8, cudaMemcpyHostToDevice, st));
    IDD_CUDA(cudaMemcpyAsync(t3 + Tz, rel, Tz * 8, cudaMemcpyHostToDevice, st));
    IDD_CUDA(cudaMemcpyAsync(t3 + 2 * Tz, tail, Tz * 8, cudaMemcpyHostToDevice, st));
    dh = t3; dr = t3 + Tz; dt = t3 + 2 * Tz;
  }
  IDD_CHECK(err != nullptr && ar.off <= ar.cap, IDDGCN_ENOMEM, "internal: arena too small");
  IDD_CUDA(cudaMemsetAsync(err, 0, sizeof(int), st));
  plan_pack_kernel<<<blocks_for(Tz, 256), 256, 0, st>>>(dh, dr, dt, Tz, N, R, p->head, p->rel, p->tail, err);
  IDD_LAUNCH_CHECK();
  const int bn = ceil_log2_u64((uint64_t)(N > 1 ? N - 1 : 1));
  const int brr = ceil_log2_u64((uint64_t)(R > 1 ? R - 1 : 1));
  IDD_TRY(plan_sort_by(p->head, Tz, (size_t)N, bn, kA, kB, vA, vB, plan, hist, scratch, st, p->by_head_perm,
                       p->by_head_ptr));
  IDD_TRY(plan_sort_by(p->tail, Tz, (size_t)N, bn, kA, kB, vA, vB, plan, hist, scratch, st, p->by_tail_perm,
                       p->by_tail_ptr));
  IDD_TRY(plan_sort_by(p->rel, Tz, (size_t)R, brr, kA, kB, vA, vB, plan, hist, scratch, st, p->by_rel_perm,
                       p->by_rel_ptr));
  plan_node_ptr_kernel<<<blocks_for((size_t)N + 1, 256), 256, 0, st>>>(p->by_head_ptr, p->by_tail_ptr, (size_t)N,
                                                                       p->by_node_ptr);
  IDD_LAUNCH_CHECK();
  plan_node_fill_kernel<<<blocks_for((size_t)N * 32, 256), 256, 0, st>>>(p->by_head_ptr, p->by_head_perm, p->by_tail_ptr,
                                                                         p->by_tail_perm, (size_t)N, (int32_t)T,
                                                                         p->by_node_ptr, p->by_node_src);
  IDD_LAUNCH_CHECK();
  int herr = 0;
  IDD_CUDA(cudaMemcpyAsync(&herr, err, sizeof(int), cudaMemcpyDeviceToHost, st));
  IDD_CUDA(cudaStreamSynchronize(st));  // the arena is freed on return
  IDD_CHECK(herr == 0, IDDGCN_ERANGE, "triple index outside [0,%lld) or relation outside [0,%d)", (long long)N, R);
  return IDDGCN_OK;
}

static void plan_free(iddgcn_plan* p) {
  if (!p) return;
  pool_free(p->block);   // the nine index arrays are views into it
  delete p;
}

__global__ void widen_concat_kernel(const int32_t* __restrict__ a, size_t na, const int32_t* __restrict__ b, size_t nb,
                                    int64_t* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= na + nb) return;
  out[i] = i < na ? (int64_t)a[i] : (int64_t)b[i - na];
}

int plan_concat(const iddgcn_plan* a, const iddgcn_plan* b, cudaStream_t st, iddgcn_plan** out) {
  IDD_CHECK(a && b && out && a->N == b->N && a->R == b->R, IDDGCN_EINVAL, "plans do not match");
  const size_t na = (size_t)a->T, nb = (size_t)b->T, n = na + nb;
  DevBuf<int64_t> tmp;
  IDD_CUDA(tmp.ensure(3 * (n ? n : 1)));
  if (n) {
    widen_concat_kernel<<<blocks_for(n, 256), 256, 0, st>>>(a->head, na, b->head, nb, tmp.p);
    IDD_LAUNCH_CHECK();
    widen_concat_kernel<<<blocks_for(n, 256), 256, 0, st>>>(a->rel, na, b->rel, nb, tmp.p + n);
    IDD_LAUNCH_CHECK();
    widen_concat_kernel<<<blocks_for(n, 256), 256, 0, st>>>(a->tail, na, b->tail, nb, tmp.p + 2 * n);
    IDD_LAUNCH_CHECK();
  }
  iddgcn_plan* p = new iddgcn_plan();
  int rc = plan_create_impl(tmp.p, tmp.p + n, tmp.p + 2 * n, (int64_t)n, 1, a->N, a->R, st, p);
  if (rc != IDDGCN_OK) {
    plan_free(p);
    return rc;
  }
  IDD_CUDA(cudaStreamSynchronize(st));
  *out = p;
  return IDDGCN_OK;
}

void plan_destroy_internal(iddgcn_plan* p) { plan_free(p); }

}  // namespace iddgcn

using namespace iddgcn;

extern "C" {

int iddgcn_graph_build(const int64_t* triples, int64_t M, int triples_on_device, int64_t num_entities,
                       int num_relations, int64_t row_lo, int64_t row_hi, void* stream, iddgcn_graph_t** out) {
  IDD_CHECK(out != nullptr, IDDGCN_EINVAL, "out is NULL");
  *out = nullptr;
  IDD_CHECK(M >= 0 && (M == 0 || triples != nullptr), IDDGCN_EINVAL, "triples is NULL with M=%lld", (long long)M);
  IDD_CHECK(num_entities > 0 && num_relations > 0, IDDGCN_EINVAL, "num_entities and num_relations must be positive");
  // row_lo == row_hi is a valid (empty) block: with N not divisible by the number of row blocks the trailing ranks of a
  // partition own nothing, and they still have to take part in the exchange
  IDD_CHECK(row_lo >= 0 && row_lo <= row_hi && row_hi <= num_entities, IDDGCN_EINVAL,
            "row block [%lld,%lld) is not inside [0,%lld)", (long long)row_lo, (long long)row_hi,
            (long long)num_entities);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("no CUDA device: libiddgcn_b200 has no CPU fallback");
    return IDDGCN_ECUDA;
  }
  iddgcn_graph* g = new iddgcn_graph();
  int rc = graph_build_impl(triples, M, triples_on_device, num_entities, num_relations, row_lo, row_hi,
                            (cudaStream_t)stream, g);
  if (rc != IDDGCN_OK) {
    graph_free(g);
    return rc;
  }
  *out = g;
  return IDDGCN_OK;
}

int iddgcn_graph_destroy(iddgcn_graph_t* g) {
  graph_free(g);
  return IDDGCN_OK;
}

int iddgcn_graph_info(const iddgcn_graph_t* g, int64_t* num_entities, int* num_relations, int64_t* n_rows,
                      int64_t* row_lo, int64_t* nnz, int64_t* nnz_per_rel) {
  IDD_CHECK(g != nullptr, IDDGCN_EINVAL, "graph is NULL");
  if (num_entities) *num_entities = g->N;
  if (num_relations) *num_relations = g->R;
  if (n_rows) *n_rows = g->n_rows;
  if (row_lo) *row_lo = g->row_lo;
  if (nnz) *nnz = g->nnz;
  if (nnz_per_rel)
    for (int r = 0; r < g->R; ++r) nnz_per_rel[r] = g->nnz_rel[r];
  return IDDGCN_OK;
}

int iddgcn_graph_export_host(const iddgcn_graph_t* g, int32_t* csr_rowptr, int32_t* csr_col, int32_t* csr_row,
                             int32_t* csc_rowptr, int32_t* csc_row, int32_t* csc_perm) {
  IDD_CHECK(g != nullptr, IDDGCN_EINVAL, "graph is NULL");
  const size_t e = (size_t)g->nnz * sizeof(int32_t);
  if (csr_rowptr)
    IDD_CUDA(cudaMemcpy(csr_rowptr, g->csr_rowptr, ((size_t)g->R * g->n_rows + 1) * 4, cudaMemcpyDeviceToHost));
  if (csc_rowptr)
    IDD_CUDA(cudaMemcpy(csc_rowptr, g->csc_rowptr, ((size_t)g->R * g->N + 1) * 4, cudaMemcpyDeviceToHost));
  if (e) {
    if (csr_col) IDD_CUDA(cudaMemcpy(csr_col, g->csr_col, e, cudaMemcpyDeviceToHost));
    if (csr_row) IDD_CUDA(cudaMemcpy(csr_row, g->csr_row, e, cudaMemcpyDeviceToHost));
    if (csc_row) IDD_CUDA(cudaMemcpy(csc_row, g->csc_row, e, cudaMemcpyDeviceToHost));
    if (csc_perm) IDD_CUDA(cudaMemcpy(csc_perm, g->csc_perm, e, cudaMemcpyDeviceToHost));
  }
  return IDDGCN_OK;
}

int iddgcn_graph_device_ptrs(const iddgcn_graph_t* g, const int32_t** csr_rowptr, const int32_t** csr_col,
                             const int32_t** csr_row, const int32_t** csc_rowptr, const int32_t** csc_row,
                             const int32_t** csc_perm) {
  IDD_CHECK(g != nullptr, IDDGCN_EINVAL, "graph is NULL");
  if (csr_rowptr) *csr_rowptr = g->csr_rowptr;
  if (csr_col) *csr_col = g->csr_col;
  if (csr_row) *csr_row = g->csr_row;
  if (csc_rowptr) *csc_rowptr = g->csc_rowptr;
  if (csc_row) *csc_row = g->csc_row;
  if (csc_perm) *csc_perm = g->csc_perm;
  return IDDGCN_OK;
}

int iddgcn_plan_create(const int64_t* head, const int64_t* rel, const int64_t* tail, int64_t T, int on_device,
                       int64_t num_entities, int num_relations, void* stream, iddgcn_plan_t** out) {
  IDD_CHECK(out != nullptr, IDDGCN_EINVAL, "out is NULL");
  *out = nullptr;
  IDD_CHECK(T >= 0 && (T == 0 || (head && rel && tail)), IDDGCN_EINVAL, "NULL triple arrays with T=%lld",
            (long long)T);
  IDD_CHECK(T < (int64_t)0x7fffff00, IDDGCN_ELIMIT, "T exceeds int32 range");
  IDD_CHECK(num_entities > 0 && num_relations > 0, IDDGCN_EINVAL, "num_entities and num_relations must be positive");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("no CUDA device: libiddgcn_b200 has no CPU fallback");
    return IDDGCN_ECUDA;
  }
  iddgcn_plan* p = new iddgcn_plan();
  int rc = plan_create_impl(head, rel, tail, T, on_device, num_entities, num_relations, (cudaStream_t)stream, p);
  if (rc != IDDGCN_OK) {
    plan_free(p);
    return rc;
  }
  *out = p;
  return IDDGCN_OK;
}

int iddgcn_plan_destroy(iddgcn_plan_t* p) {
  plan_free(p);
  return IDDGCN_OK;
}

}  // extern "C"
