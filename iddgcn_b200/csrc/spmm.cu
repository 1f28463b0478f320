// spmm.cu — relational SpMM forward, transposed backward, and SDDMM (edge-value gradient).
//
// Replaces tf.sparse.sparse_dense_matmul(adj_i, embeddings) (prediction/IDDGCN.py:70) for all
// relations in one launch, plus its TF gradient w.r.t. the dense operand (adjoint SpMM) and
// w.r.t. the sparse values (explanation/explaiNE.py:17, GnnExplainer.py:51).
//
// HBM-bound gather kernels: one sub-warp of LPR lanes owns one output row; every lane keeps
// float4 accumulators for its 16-byte slice(s) of the feature row; edge indices of a row are
// loaded coalesced (one per lane) and broadcast by shuffle; U independent 128-bit row gathers
// are issued before any is consumed.  The forward walks CSR rows (r, head); the backward
// walks CSC rows (r, tail) with all relations fused into a single dX write, so neither needs
// atomics and the summation order is fixed (ascending neighbour index, as TF-CPU does).
// Edge weights (explicit values or sigmoid(mask)) are applied in registers.
#include <stdlib.h>
#include "dlpack.cuh"
#include "graph.cuh"

namespace iddgcn {

constexpr int SPMM_THREADS = 256;

// Extra addends of an output row: out[o] = gathered sum + Σ_j add.p[j][o].  n == 0: plain store; the classic
// "accumulate" (out += ...) is add.p[0] == out.  The partitioned backward uses it to fold the partial dX rows that
// peer GPUs pushed into this rank's stage buffers into the epilogue of the kernel that computes the rank's own
// partial (fixed order: own partial, then peers in rank order), so the reduce-scatter needs no reduction pass.
constexpr int MAX_ADD = 8;
struct AddSrcs {
  const float* p[MAX_ADD];
  int n;
};
static inline AddSrcs add_none() { AddSrcs a{}; a.n = 0; return a; }
static inline AddSrcs add_self(const float* out) { AddSrcs a{}; a.n = 1; a.p[0] = out; return a; }

// statically indexed (fully unrolled, predicated): a run-time index into the by-value parameter struct would make the
// compiler copy it to local memory and spill around every flush
__device__ __forceinline__ void add_rows(float4& r, const AddSrcs& add, size_t vec_index) {
  if (add.n == 0) return;
#pragma unroll
  for (int j = 0; j < MAX_ADD; ++j) {
    if (j < add.n) {
      const float4 prev = *(reinterpret_cast<const float4*>(add.p[j]) + vec_index);
      r.x += prev.x; r.y += prev.y; r.z += prev.z; r.w += prev.w;
    }
  }
}

// Σ_j add.p[j][vec_index] in the order given (zero when there is none) — the value the epilogue adds to a row
__device__ __forceinline__ float4 addend_sum(const AddSrcs& add, size_t vec_index) {
  // One addend (two row blocks, or `accumulate`) is the common case and gets a single unconditional load: in the
  // general predicated chain below every `t += load_j` reuses the same registers, so the instruction after the first
  // load waits for it even when j >= n — which turned the prefetch back into a dependent load (3.05 ms instead of 1.4 ms
  // for the own-block launch, profiles/r02_addend_epilogue_microbench.json).
  if (add.n == 1) return *(reinterpret_cast<const float4*>(add.p[0]) + vec_index);
  float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
  add_rows(t, add, vec_index);
  return t;
}

template <int MODE>
__device__ __forceinline__ float edge_weight(const float* __restrict__ vals, int e) {
  if (MODE == IDDGCN_VALS_ONES) return 1.0f;
  const float v = __ldg(vals + e);
  if (MODE == IDDGCN_VALS_SIGMOID) return sigmoidf_(v);
  return v;
}

// Generic segmented gather-accumulate.
//   out[o, :] (+)= sum_{gi < G} sum_{e in ptr[gi*seg_stride + o] .. ptr[gi*seg_stride + o + 1]}
//                       w(e) * src[(gi*src_group_rows + idx[e]), :]
// forward : G=1, ptr=csr_rowptr, idx=csr_col, src=X,  o over R*n_rows, weight index = e
// backward: G=R, ptr=csc_rowptr, idx=csc_row, src=dS, o over N,        weight index = perm[e]
template <int LPR, int CH, int U, int MODE, bool USE_PERM>
__global__ void __launch_bounds__(SPMM_THREADS)
    gather_rows_kernel(const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                       const int32_t* __restrict__ perm, const float* __restrict__ vals,
                       const float* __restrict__ src, float* __restrict__ out, int64_t row_begin, int64_t row_end,
                       int G, int64_t seg_stride, int64_t src_group_rows, int nvec /* D/4 */, AddSrcs add,
                       int heavy_t) {
  constexpr int SUBS = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR;
  const int l = lane % LPR;
  const int64_t warp_global = (int64_t)blockIdx.x * (SPMM_THREADS / 32) + (threadIdx.x >> 5);
  const int64_t o = row_begin + warp_global * SUBS + sub;
  if (o >= row_end) return;
  const unsigned mask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (sub * LPR));
  const size_t D = (size_t)nvec * 4;

  float4 acc[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int gi = 0; gi < G; ++gi) {
    const int64_t s = (int64_t)gi * seg_stride + o;
    const int beg = ld_stream_i32(ptr + s);
    const int end = ld_stream_i32(ptr + s + 1);
    if (end - beg > heavy_t) return;  // hub row: written by the split pass (only used with G == 1)
    const float4* __restrict__ sbase = reinterpret_cast<const float4*>(src + (size_t)gi * src_group_rows * D);
    for (int base = beg; base < end; base += LPR) {
      const int my_e = base + l;
      int my_idx = 0;
      float my_w = 0.f;
      if (my_e < end) {
        my_idx = ld_stream_i32(idx + my_e);
        if (MODE != IDDGCN_VALS_ONES) my_w = edge_weight<MODE>(vals, USE_PERM ? ld_stream_i32(perm + my_e) : my_e);
      }
      const int cnt = (end - base < LPR) ? (end - base) : LPR;
      for (int j = 0; j < cnt; j += U) {
        float4 v[U][CH];
        float w[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int jj = j + u;
          const int sidx = __shfl_sync(mask, my_idx, jj & (LPR - 1), LPR);
          if (MODE != IDDGCN_VALS_ONES) w[u] = __shfl_sync(mask, my_w, jj & (LPR - 1), LPR);
          const float4* rowp = sbase + (size_t)sidx * nvec;
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            const int vi = l + c * LPR;
            if (jj < cnt && vi < nvec) v[u][c] = __ldg(rowp + vi);
            else v[u][c] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            if (MODE == IDDGCN_VALS_ONES) {
              acc[c].x += v[u][c].x; acc[c].y += v[u][c].y; acc[c].z += v[u][c].z; acc[c].w += v[u][c].w;
            } else {
              acc[c].x = fmaf(w[u], v[u][c].x, acc[c].x); acc[c].y = fmaf(w[u], v[u][c].y, acc[c].y);
              acc[c].z = fmaf(w[u], v[u][c].z, acc[c].z); acc[c].w = fmaf(w[u], v[u][c].w, acc[c].w);
            }
          }
        }
      }
    }
  }
  float4* orow = reinterpret_cast<float4*>(out + (size_t)o * D);
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int vi = l + c * LPR;
    if (vi < nvec) {
      float4 r = acc[c];
      const float4 t = addend_sum(add, (size_t)o * nvec + vi);
      r.x += t.x; r.y += t.y; r.z += t.z; r.w += t.w;
      st_stream_f4(orow + vi, r);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// v2 gather: edge-index tiles staged in shared memory, edges consumed as one flat stream.
//
// A slice of LPR lanes owns a CHUNK of LPR consecutive output rows.  Lane l fetches the end pointer of row l (one
// coalesced load for the whole chunk); the chunk's index range [ptr[r0], ptr[r0+LPR]) is contiguous, so it is staged
// into a shared-memory slab with coalesced loads (weights too, with the sigmoid applied once per edge).  The slice
// then walks the edges in batches of U regardless of row boundaries — U independent 128-bit row gathers are always in
// flight — and flushes the accumulator whenever the stream crosses a row end (a slice-uniform test).  Per chunk there
// are two dependent index latencies instead of two per row, which is what limited v1 at an average degree of 10.
//   out[o, :] (+)= sum_{e in ptr[o]..ptr[o+1]} w(e) * src[idx[e], :]
// forward : ptr=csr_rowptr, idx=csr_col,  src=X,                  weight index e
// backward: ptr=bwd_rowptr, idx=bwd_src,  src=dS as [(R·n_rows),D], weight index perm[e]
constexpr int SLAB = 512;  // staged indices per warp

template <int LPR, int U, int MODE, bool USE_PERM, bool HUBS, int TPB = SPMM_THREADS, int MINB = 4, bool NA = false,
          bool ADD = false>
__global__ void __launch_bounds__(TPB, MINB)
    stream_rows_kernel(const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                       const int32_t* __restrict__ perm, const float* __restrict__ vals,
                       const float* __restrict__ src, float* __restrict__ out, int64_t row_begin, int64_t row_end,
                       int nvec, AddSrcs add, int heavy_t) {
  constexpr int SUBS = 32 / LPR;
  constexpr int SUB_SLAB = SLAB / SUBS;
  constexpr bool WEIGHTED = MODE != IDDGCN_VALS_ONES;
  // Narrow rows (D <= 32) with epilogue addends: the addend rows of the slice's LPR output rows are contiguous, so the
  // slice stages their sum in shared memory once per chunk (LPR x LPR float4, coalesced) and the flush adds from there.
  // With 4 or 8 slices per warp flushing at different moments a per-flush global load — even one requested two rows
  // ahead — made the own-block launch of the partitioned backward 1.9x slower than the same launch without addends
  // (profiles/r02_addend_epilogue_microbench.json); wide rows (one or two slices per warp) keep the register prefetch.
  constexpr bool ADD_SMEM = ADD && LPR <= 8 && !WEIGHTED && !HUBS;
  __shared__ int32_t s_idx[TPB / 32][SLAB];
  __shared__ float s_w[WEIGHTED ? TPB / 32 : 1][WEIGHTED ? SLAB : 1];
  __shared__ float4 s_add[ADD_SMEM ? TPB / 32 : 1][ADD_SMEM ? 32 * LPR : 1];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR;
  const int l = lane % LPR;
  const int64_t chunk = ((int64_t)blockIdx.x * (TPB / 32) + warp) * SUBS + sub;
  const int64_t r0 = row_begin + chunk * LPR;
  if (r0 >= row_end) return;
  const unsigned mask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (sub * LPR));
  const int nrows = (row_end - r0 < LPR) ? (int)(row_end - r0) : LPR;
  int32_t* const sl_idx = &s_idx[warp][sub * SUB_SLAB];
  float* const sl_w = WEIGHTED ? &s_w[warp][sub * SUB_SLAB] : nullptr;

  const int chunk_beg = ld_stream_i32(ptr + r0);
  const int my_end = (l < nrows) ? ld_stream_i32(ptr + r0 + l + 1) : 0x7fffffff;
  const int chunk_end = __shfl_sync(mask, my_end, nrows - 1, LPR);
  const float4* __restrict__ src4 = reinterpret_cast<const float4*>(src);
  float4* __restrict__ out4 = reinterpret_cast<float4*>(out) + (size_t)r0 * nvec;

  if constexpr (!HUBS) {
    // graph without hub rows: the tight loop (no per-row degree test)
    int row = 0;
    int cur_end = __shfl_sync(mask, my_end, 0, LPR);
    int pos = chunk_beg, slab_lo = chunk_beg, slab_hi = chunk_beg;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    // The addends of a row (staged peer partials) are requested two rows ahead: loaded inside the flush they were a
    // dependent ~1 us load at every row end (the own-block launch of the partitioned backward ran at half speed).
    float4 pre_a = make_float4(0.f, 0.f, 0.f, 0.f), pre_b = pre_a;
    float4* const sl_add = ADD_SMEM ? &s_add[warp][sub * LPR * LPR] : nullptr;
    if constexpr (ADD_SMEM) {
      const int total = nrows * nvec;                        // the chunk's addend rows are one contiguous range
      for (int i = l; i < total; i += LPR) sl_add[i] = addend_sum(add, (size_t)r0 * nvec + i);
      __syncwarp(mask);
    } else if constexpr (ADD) {
      if (l < nvec) {
        pre_a = addend_sum(add, (size_t)r0 * nvec + l);
        if (nrows > 1) pre_b = addend_sum(add, (size_t)(r0 + 1) * nvec + l);
      }
    }

    auto flush = [&]() {
      if (l < nvec) {
        float4 r = acc;
        float4* o = out4 + (size_t)row * nvec + l;
        if constexpr (ADD_SMEM) {
          const float4 t = sl_add[row * nvec + l];
          r.x += t.x; r.y += t.y; r.z += t.z; r.w += t.w;
        } else if constexpr (ADD) {
          r.x += pre_a.x; r.y += pre_a.y; r.z += pre_a.z; r.w += pre_a.w;
          pre_a = pre_b;
          if (row + 2 < nrows) pre_b = addend_sum(add, (size_t)(r0 + row + 2) * nvec + l);
        }
        st_stream_f4(o, r);
      }
      acc = make_float4(0.f, 0.f, 0.f, 0.f);
      ++row;
      cur_end = __shfl_sync(mask, my_end, row & (LPR - 1), LPR);
      if (row >= nrows) cur_end = 0x7fffffff;
    };

    while (pos < chunk_end) {
      if (pos + U > slab_hi && slab_hi < chunk_end) {  // refill the slab starting at the current position
        __syncwarp(mask);
        slab_lo = pos;
        slab_hi = (chunk_end - pos < SUB_SLAB) ? chunk_end : pos + SUB_SLAB;
        for (int q = l; q < slab_hi - slab_lo; q += LPR) {
          const int e = slab_lo + q;
          sl_idx[q] = ld_stream_i32(idx + e);
          if (WEIGHTED) sl_w[q] = edge_weight<MODE>(vals, USE_PERM ? ld_stream_i32(perm + e) : e);
        }
        __syncwarp(mask);
      }
      const int n = (slab_hi - pos < U) ? (slab_hi - pos) : U;
      float4 v[U];
      float w[U];
  #pragma unroll
      for (int u = 0; u < U; ++u) {
        if (u < n) {
          const int sidx = sl_idx[pos + u - slab_lo];
          if (WEIGHTED) w[u] = sl_w[pos + u - slab_lo];
          v[u] = (l < nvec) ? (NA ? ld_stream_f4(src4 + (size_t)sidx * nvec + l) : __ldg(src4 + (size_t)sidx * nvec + l))
                           : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
  #pragma unroll
      for (int u = 0; u < U; ++u) {
        if (u < n) {
          while (cur_end == pos + u) flush();
          if (WEIGHTED) {
            acc.x = fmaf(w[u], v[u].x, acc.x); acc.y = fmaf(w[u], v[u].y, acc.y);
            acc.z = fmaf(w[u], v[u].z, acc.z); acc.w = fmaf(w[u], v[u].w, acc.w);
          } else {
            acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
          }
        }
      }
      pos += n;
    }
    while (row < nrows) flush();
  } else {
    int row = 0, row_beg = chunk_beg;
    int cur_end = __shfl_sync(mask, my_end, 0, LPR);
    bool heavy = cur_end - row_beg > heavy_t;  // hub rows are skipped here and written by the split pass
    int pos = chunk_beg, slab_lo = chunk_beg, slab_hi = chunk_beg;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);

    auto store_row = [&]() {
      if (l < nvec) {
        float4 r = acc;
        float4* o = out4 + (size_t)row * nvec + l;
        if constexpr (ADD) {
          const float4 t = addend_sum(add, (size_t)(r0 + row) * nvec + l);
          r.x += t.x; r.y += t.y; r.z += t.z; r.w += t.w;
        }
        st_stream_f4(o, r);
      }
      acc = make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto next_row = [&]() {
      row_beg = cur_end;
      ++row;
      cur_end = __shfl_sync(mask, my_end, row & (LPR - 1), LPR);
      heavy = false;
      if (row >= nrows) cur_end = 0x7fffffff;
      else heavy = cur_end - row_beg > heavy_t;
    };

    while (row < nrows) {
      if (heavy) {  // jump over the hub row's edges
        pos = cur_end;
        next_row();
        continue;
      }
      if (pos == cur_end) {  // row complete (or empty)
        store_row();
        next_row();
        continue;
      }
      if (pos + U > slab_hi && slab_hi < chunk_end) {  // refill the slab starting at the current position
        __syncwarp(mask);
        slab_lo = pos;
        slab_hi = (chunk_end - pos < SUB_SLAB) ? chunk_end : pos + SUB_SLAB;
        for (int q = l; q < slab_hi - slab_lo; q += LPR) {
          const int e = slab_lo + q;
          sl_idx[q] = ld_stream_i32(idx + e);
          if (WEIGHTED) sl_w[q] = edge_weight<MODE>(vals, USE_PERM ? ld_stream_i32(perm + e) : e);
        }
        __syncwarp(mask);
      }
      const int n = (slab_hi - pos < U) ? (slab_hi - pos) : U;
      float4 v[U];
      float w[U];
  #pragma unroll
      for (int u = 0; u < U; ++u) {
        if (u < n) {
          const int sidx = sl_idx[pos + u - slab_lo];
          if (WEIGHTED) w[u] = sl_w[pos + u - slab_lo];
          v[u] = (l < nvec) ? (NA ? ld_stream_f4(src4 + (size_t)sidx * nvec + l) : __ldg(src4 + (size_t)sidx * nvec + l))
                           : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      int consumed = 0;
      bool stop = false;
  #pragma unroll
      for (int u = 0; u < U; ++u) {
        if (u < n && !stop) {
          if (pos + u == cur_end) {  // the stream crosses into the next row(s)
            store_row();
            next_row();
            while (row < nrows && !heavy && cur_end == pos + u) {  // empty rows
              store_row();
              next_row();
            }
            stop = heavy || row >= nrows;  // a hub row starts here: drop the rest of this batch
          }
          if (!stop) {
            if (WEIGHTED) {
              acc.x = fmaf(w[u], v[u].x, acc.x); acc.y = fmaf(w[u], v[u].y, acc.y);
              acc.z = fmaf(w[u], v[u].z, acc.z); acc.w = fmaf(w[u], v[u].w, acc.w);
            } else {
              acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
            }
            ++consumed;
          }
        }
      }
      pos += consumed;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Split pass for hub rows (degree > HEAVY_EDGES): one lane slice per HEAVY_EDGES-edge segment writes a partial sum,
// then one slice per hub row adds its partials in ascending segment order.  Deterministic; runs only when the graph
// has hub rows (none in the uniform benchmark graphs, many in Zipf-skewed ones).
template <int LPR, int CH, int U, int MODE, bool USE_PERM>
__global__ void __launch_bounds__(SPMM_THREADS)
    heavy_partial_kernel(const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                         const int32_t* __restrict__ perm, const float* __restrict__ vals,
                         const float* __restrict__ src, const int32_t* __restrict__ hub_rows,
                         const int32_t* __restrict__ hub_segptr, int H, int total_segs, float* __restrict__ scratch,
                         int nvec) {
  constexpr int SUBS = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR;
  const int l = lane % LPR;
  const int64_t sg = ((int64_t)blockIdx.x * (SPMM_THREADS / 32) + (threadIdx.x >> 5)) * SUBS + sub;
  if (sg >= total_segs) return;
  const unsigned mask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (sub * LPR));
  int lo = 0, hi = H;  // last hub row whose first segment is <= sg
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (hub_segptr[mid] <= sg) lo = mid; else hi = mid;
  }
  const int64_t row = hub_rows[lo];
  const int k = (int)(sg - hub_segptr[lo]);
  const int rbeg = ptr[row], rend = ptr[row + 1];
  const int beg = rbeg + k * HEAVY_EDGES;
  const int end = (beg + HEAVY_EDGES < rend) ? beg + HEAVY_EDGES : rend;
  const float4* __restrict__ sbase = reinterpret_cast<const float4*>(src);
  float4 acc[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int base = beg; base < end; base += LPR) {
    const int my_e = base + l;
    int my_idx = 0;
    float my_w = 0.f;
    if (my_e < end) {
      my_idx = ld_stream_i32(idx + my_e);
      if (MODE != IDDGCN_VALS_ONES) my_w = edge_weight<MODE>(vals, USE_PERM ? ld_stream_i32(perm + my_e) : my_e);
    }
    const int cnt = (end - base < LPR) ? (end - base) : LPR;
    for (int j = 0; j < cnt; j += U) {
      float4 v[U][CH];
      float w[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int jj = j + u;
        const int sidx = __shfl_sync(mask, my_idx, jj & (LPR - 1), LPR);
        if (MODE != IDDGCN_VALS_ONES) w[u] = __shfl_sync(mask, my_w, jj & (LPR - 1), LPR);
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          const int vi = l + c * LPR;
          v[u][c] = (jj < cnt && vi < nvec) ? __ldg(sbase + (size_t)sidx * nvec + vi) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          const float ww = (MODE == IDDGCN_VALS_ONES) ? 1.0f : w[u];
          acc[c].x = fmaf(ww, v[u][c].x, acc[c].x); acc[c].y = fmaf(ww, v[u][c].y, acc[c].y);
          acc[c].z = fmaf(ww, v[u][c].z, acc[c].z); acc[c].w = fmaf(ww, v[u][c].w, acc[c].w);
        }
    }
  }
  float4* o = reinterpret_cast<float4*>(scratch) + (size_t)sg * nvec;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int vi = l + c * LPR;
    if (vi < nvec) o[vi] = acc[c];
  }
}

__global__ void __launch_bounds__(256)
    heavy_reduce_kernel(const int32_t* __restrict__ hub_rows, const int32_t* __restrict__ hub_segptr, int H,
                        const float* __restrict__ scratch, float* __restrict__ out, int64_t row_begin, int64_t row_end,
                        int nvec, AddSrcs add) {
  const int h = blockIdx.x;
  const int64_t row = hub_rows[h];
  if (row < row_begin || row >= row_end) return;
  const float4* s4 = reinterpret_cast<const float4*>(scratch);
  float4* o4 = reinterpret_cast<float4*>(out) + (size_t)row * nvec;
  for (int vi = threadIdx.x; vi < nvec; vi += blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int sg = hub_segptr[h]; sg < hub_segptr[h + 1]; ++sg) {
      const float4 p = s4[(size_t)sg * nvec + vi];
      acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
    }
    const float4 t = addend_sum(add, (size_t)row * nvec + vi);   // same association as the row kernels: acc + (Σ_j a_j)
    acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    o4[vi] = acc;
  }
}

// SDDMM: dvals[e] = <dS[seg(e), :], X[col[e], :]> (· σ'(vals[e]) in SIGMOID mode), CSR edge order.
// One lane slice walks the edge range [beg, end) of output row o: the dS row stays in registers, every edge costs one
// 128-bit gather of X[col] per lane and a shuffle reduction, results are written LPR at a time (coalesced).
template <int LPR, int CH, int U, int MODE>
__device__ __forceinline__ void sddmm_range(const int32_t* __restrict__ idx, const float* __restrict__ vals,
                                            const float* __restrict__ dS, const float* __restrict__ X,
                                            float* __restrict__ dvals, int64_t o, int beg, int end, int nvec,
                                            unsigned mask, int l) {
  const float4* __restrict__ grow = reinterpret_cast<const float4*>(dS) + (size_t)o * nvec;
  const float4* __restrict__ xb = reinterpret_cast<const float4*>(X);
  float4 g[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int vi = l + c * LPR;
    g[c] = (vi < nvec) ? ld_stream_f4(grow + vi) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int base = beg; base < end; base += LPR) {
    const int my_e = base + l;
    const int my_idx = (my_e < end) ? ld_stream_i32(idx + my_e) : 0;
    const int cnt = (end - base < LPR) ? (end - base) : LPR;
    float my_out = 0.f;
    for (int j = 0; j < cnt; j += U) {
      float4 v[U][CH];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int jj = j + u;
        const int sidx = __shfl_sync(mask, my_idx, jj & (LPR - 1), LPR);
        const float4* rowp = xb + (size_t)sidx * nvec;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          const int vi = l + c * LPR;
          if (jj < cnt && vi < nvec) v[u][c] = __ldg(rowp + vi);
          else v[u][c] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float p = 0.f;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          p = fmaf(g[c].x, v[u][c].x, p); p = fmaf(g[c].y, v[u][c].y, p);
          p = fmaf(g[c].z, v[u][c].z, p); p = fmaf(g[c].w, v[u][c].w, p);
        }
#pragma unroll
        for (int off = LPR / 2; off > 0; off >>= 1) p += __shfl_xor_sync(mask, p, off, LPR);
        if (l == j + u) my_out = p;  // lane l keeps the result of edge base+l
      }
    }
    if (my_e < end) {
      float r = my_out;
      if (MODE == IDDGCN_VALS_SIGMOID) {
        const float sg = sigmoidf_(__ldg(vals + my_e));
        r *= sg * (1.0f - sg);
      }
      dvals[my_e] = r;
    }
  }
}

template <int LPR, int CH, int U, int MODE>
__global__ void __launch_bounds__(SPMM_THREADS)
    sddmm_rows_kernel(const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx, const float* __restrict__ vals,
                      const float* __restrict__ dS, const float* __restrict__ X, float* __restrict__ dvals,
                      int64_t num_rows, int nvec, int heavy_t) {
  constexpr int SUBS = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR;
  const int l = lane % LPR;
  const int64_t o = ((int64_t)blockIdx.x * (SPMM_THREADS / 32) + (threadIdx.x >> 5)) * SUBS + sub;
  if (o >= num_rows) return;
  const unsigned mask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (sub * LPR));
  const int beg = ld_stream_i32(ptr + o);
  const int end = ld_stream_i32(ptr + o + 1);
  if (beg >= end || end - beg > heavy_t) return;  // hub rows: sddmm_hub_kernel
  sddmm_range<LPR, CH, U, MODE>(idx, vals, dS, X, dvals, o, beg, end, nvec, mask, l);
}

// hub rows: one lane slice per HEAVY_EDGES-edge segment (edge outputs are independent, so no combine step)
template <int LPR, int CH, int U, int MODE>
__global__ void __launch_bounds__(SPMM_THREADS)
    sddmm_hub_kernel(const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx, const float* __restrict__ vals,
                     const float* __restrict__ dS, const float* __restrict__ X, float* __restrict__ dvals,
                     const int32_t* __restrict__ hub_rows, const int32_t* __restrict__ hub_segptr, int H,
                     int total_segs, int nvec) {
  constexpr int SUBS = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR;
  const int l = lane % LPR;
  const int64_t sg = ((int64_t)blockIdx.x * (SPMM_THREADS / 32) + (threadIdx.x >> 5)) * SUBS + sub;
  if (sg >= total_segs) return;
  const unsigned mask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (sub * LPR));
  int lo = 0, hi = H;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (hub_segptr[mid] <= sg) lo = mid; else hi = mid;
  }
  const int64_t o = hub_rows[lo];
  const int rbeg = ptr[o], rend = ptr[o + 1];
  const int beg = rbeg + (int)(sg - hub_segptr[lo]) * HEAVY_EDGES;
  const int end = (beg + HEAVY_EDGES < rend) ? beg + HEAVY_EDGES : rend;
  sddmm_range<LPR, CH, U, MODE>(idx, vals, dS, X, dvals, o, beg, end, nvec, mask, l);
}

// ---------------------------------------------------------------------------------------------
// v2 SDDMM: the flat-edge-stream pattern of stream_rows_kernel applied to the edge-value gradient.
//
// A slice of LPR lanes owns a chunk of LPR consecutive CSR rows; the chunk's index range is staged in a shared-memory
// slab and consumed in batches of U edges regardless of row boundaries.  Per batch the slice first walks the row ends
// (registers + shuffles only) to learn which edges start a new row, then issues U independent 128-bit gathers of
// X[col] together with the (predicated) loads of the dS rows that begin inside the batch — an edge that continues the
// previous edge's row reuses its register copy — so a row switch costs no dependent memory latency.  The U dot
// products are reduced with a transpose reduction (U/2 + U/4 + ... shuffles instead of U·log2(LPR)) that leaves edge
// i of the batch in lane i·(LPR/U), and those lanes write U consecutive floats (one 32-byte sector).
//   dvals[e] = <dS[row(e), :], X[idx[e], :]>  (· σ'(vals[e]) in SIGMOID mode, staged once per edge with the indices)
template <int LPR, int NV, int OFF>
struct SliceReduce {
  static __device__ __forceinline__ void run(float* v, int l, unsigned mask) {
    if constexpr (OFF >= 1) {
      if constexpr (NV > 1) {
        constexpr int H = NV / 2;
        const bool hi = (l & OFF) != 0;
#pragma unroll
        for (int i = 0; i < H; ++i) {
          const float send = hi ? v[i] : v[i + H];
          const float keep = hi ? v[i + H] : v[i];
          v[i] = keep + __shfl_xor_sync(mask, send, OFF, LPR);
        }
        SliceReduce<LPR, H, OFF / 2>::run(v, l, mask);
      } else {
        v[0] += __shfl_xor_sync(mask, v[0], OFF, LPR);
        SliceReduce<LPR, 1, OFF / 2>::run(v, l, mask);
      }
    }
  }
};

template <int LPR, int U, int MODE, bool HUBS, int TPB = SPMM_THREADS, int MINB = 3>
__global__ void __launch_bounds__(TPB, MINB)
    sddmm_stream_kernel(const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                        const float* __restrict__ vals, const float* __restrict__ dS, const float* __restrict__ X,
                        float* __restrict__ dvals, int64_t num_rows, int nvec, int heavy_t) {
  constexpr int SUBS = 32 / LPR;
  constexpr int SUB_SLAB = SLAB / SUBS;
  constexpr bool SIG = MODE == IDDGCN_VALS_SIGMOID;
  constexpr int KEEP = (U > LPR) ? U / LPR : 1;          // results left per lane by the transpose reduction
  constexpr int STRIDE = (LPR > U) ? LPR / U : 1;        // lane spacing of the result holders
  __shared__ int32_t s_idx[TPB / 32][SLAB];
  __shared__ float s_w[SIG ? TPB / 32 : 1][SIG ? SLAB : 1];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR;
  const int l = lane % LPR;
  const int64_t chunk = ((int64_t)blockIdx.x * (TPB / 32) + warp) * SUBS + sub;
  const int64_t r0 = chunk * LPR;
  if (r0 >= num_rows) return;
  const unsigned mask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (sub * LPR));
  const int nrows = (num_rows - r0 < LPR) ? (int)(num_rows - r0) : LPR;
  int32_t* const sl_idx = &s_idx[warp][sub * SUB_SLAB];
  float* const sl_w = SIG ? &s_w[warp][sub * SUB_SLAB] : nullptr;

  const int chunk_beg = ld_stream_i32(ptr + r0);
  const int my_end = (l < nrows) ? ld_stream_i32(ptr + r0 + l + 1) : 0x7fffffff;
  const int chunk_end = __shfl_sync(mask, my_end, nrows - 1, LPR);
  const float4* __restrict__ X4 = reinterpret_cast<const float4*>(X);
  const float4* __restrict__ G4 = reinterpret_cast<const float4*>(dS) + (size_t)r0 * nvec;
  const bool live = l < nvec;

  int row = 0, row_beg = chunk_beg;
  int cur_end = __shfl_sync(mask, my_end, 0, LPR);
  bool heavy_cur = HUBS && (cur_end - row_beg > heavy_t);
  bool fresh = true;                       // the first edge of the next batch starts a row not yet loaded
  int pos = chunk_beg, slab_lo = chunk_beg, slab_hi = chunk_beg;
  float4 g_carry = make_float4(0.f, 0.f, 0.f, 0.f);

  while (true) {
    // skip finished / empty rows and hub rows (those are written by sddmm_hub_kernel)
    while (row < nrows && (pos == cur_end || heavy_cur)) {
      pos = cur_end;
      row_beg = cur_end;
      ++row;
      heavy_cur = false;
      fresh = true;
      if (row < nrows) {
        cur_end = __shfl_sync(mask, my_end, row & (LPR - 1), LPR);
        heavy_cur = HUBS && (cur_end - row_beg > heavy_t);
      }
    }
    if (row >= nrows) break;
    if (pos + U > slab_hi && slab_hi < chunk_end) {  // refill the slab starting at the current position
      __syncwarp(mask);
      slab_lo = pos;
      slab_hi = (chunk_end - pos < SUB_SLAB) ? chunk_end : pos + SUB_SLAB;
      for (int q = l; q < slab_hi - slab_lo; q += LPR) {
        const int e = slab_lo + q;
        sl_idx[q] = ld_stream_i32(idx + e);
        if (SIG) {
          const float sg = sigmoidf_(__ldg(vals + e));
          sl_w[q] = sg * (1.0f - sg);
        }
      }
      __syncwarp(mask);
    }
    int n = (slab_hi - pos < U) ? (slab_hi - pos) : U;
    float4 v[U], gs[U];
    // issue phase: row walk in registers, then every load of the batch
    int r = row, ce = cur_end, rb = row_beg;
    bool hv = false;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      bool nw = (u == 0) ? fresh : false;
      if (u < n) {
        while (ce == pos + u) {          // edge u starts a later row (rows in between are empty)
          rb = ce;
          ++r;
          ce = (r < nrows) ? __shfl_sync(mask, my_end, r & (LPR - 1), LPR) : 0x7fffffff;
          nw = true;
        }
        if (HUBS && nw && u > 0 && ce - rb > heavy_t) {  // a hub row starts here: end the batch before it
          n = u;
          hv = true;
        }
      }
      if (u < n) {
        const int sidx = sl_idx[pos + u - slab_lo];
        v[u] = live ? __ldg(X4 + (size_t)sidx * nvec + l) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (nw) gs[u] = live ? ld_stream_f4(G4 + (size_t)r * nvec + l) : make_float4(0.f, 0.f, 0.f, 0.f);
        else gs[u] = (u == 0) ? g_carry : gs[u > 0 ? u - 1 : 0];
      } else {
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        gs[u] = (u == 0) ? g_carry : gs[u > 0 ? u - 1 : 0];
      }
    }
    float p[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float a = gs[u].x * v[u].x;
      a = fmaf(gs[u].y, v[u].y, a);
      a = fmaf(gs[u].z, v[u].z, a);
      a = fmaf(gs[u].w, v[u].w, a);
      p[u] = a;
    }
    SliceReduce<LPR, U, LPR / 2>::run(p, l, mask);
    if ((l % STRIDE) == 0) {
      const int base = (l / STRIDE) * KEEP;
#pragma unroll
      for (int i = 0; i < KEEP; ++i) {
        if (base + i < n) {
          float out = p[i];
          if (SIG) out *= sl_w[pos + base + i - slab_lo];
          dvals[pos + base + i] = out;
        }
      }
    }
    g_carry = gs[U - 1];
    fresh = false;
    pos += n;
    row = r; cur_end = ce; row_beg = rb;
    heavy_cur = hv;
    if (hv) fresh = true;
  }
}

// one block per relation, fixed-order reduction
template <int MODE>
__global__ void __launch_bounds__(256) edge_value_sums_kernel(const int32_t* __restrict__ rowptr, int64_t n_rows,
                                                              const float* __restrict__ vals,
                                                              float* __restrict__ out) {
  __shared__ float s_part[256];
  const int r = blockIdx.x;
  const int beg = rowptr[(int64_t)r * n_rows], end = rowptr[(int64_t)(r + 1) * n_rows];
  float acc = 0.f;
  for (int e = beg + threadIdx.x; e < end; e += 256) acc += edge_weight<MODE>(vals, e);
  s_part[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s_part[threadIdx.x] += s_part[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[r] = s_part[0];
}

// ---------------------------------------------------------------------------
// dispatch
// ---------------------------------------------------------------------------
struct GatherArgs {
  const int32_t* ptr; const int32_t* idx; const int32_t* perm; const float* vals; const float* src; float* out;
  int64_t row_begin, row_end; int G; int64_t seg_stride, src_group_rows; int nvec; AddSrcs add; int heavy_t;
};

template <int LPR, int CH, int U, bool USE_PERM>
static int launch_gather_mode(const GatherArgs& a, int mode, cudaStream_t st) {
  const int64_t rows = a.row_end - a.row_begin;
  if (rows <= 0) return IDDGCN_OK;
  constexpr int rows_per_block = (SPMM_THREADS / 32) * (32 / LPR);
  const int64_t blocks = (rows + rows_per_block - 1) / rows_per_block;
  IDD_CHECK(blocks < (int64_t)0x7fffffff, IDDGCN_ELIMIT, "too many rows for one launch");
#define IDD_GATHER(MODE_)                                                                                        \
  gather_rows_kernel<LPR, CH, U, MODE_, USE_PERM><<<(unsigned)blocks, SPMM_THREADS, 0, st>>>(                    \
      a.ptr, a.idx, a.perm, a.vals, a.src, a.out, a.row_begin, a.row_end, a.G, a.seg_stride, a.src_group_rows,  \
      a.nvec, a.add, a.heavy_t)
  if (mode == IDDGCN_VALS_ONES) IDD_GATHER(IDDGCN_VALS_ONES);
  else if (mode == IDDGCN_VALS_EXPLICIT) IDD_GATHER(IDDGCN_VALS_EXPLICIT);
  else IDD_GATHER(IDDGCN_VALS_SIGMOID);
#undef IDD_GATHER
  IDD_LAUNCH_CHECK();
  return IDDGCN_OK;
}

template <bool USE_PERM>
static int launch_gather(const GatherArgs& a, int mode, cudaStream_t st) {
  const int nvec = a.nvec;
  if (nvec <= 8) return launch_gather_mode<8, 1, 8, USE_PERM>(a, mode, st);
  if (nvec <= 16) return launch_gather_mode<16, 1, 8, USE_PERM>(a, mode, st);
  if (nvec <= 32) return launch_gather_mode<32, 1, 8, USE_PERM>(a, mode, st);
  if (nvec <= 64) return launch_gather_mode<32, 2, 4, USE_PERM>(a, mode, st);
  if (nvec <= 128) return launch_gather_mode<32, 4, 2, USE_PERM>(a, mode, st);
  set_error("feature width D=%d > 512 is not supported", nvec * 4);
  return IDDGCN_ELIMIT;
}

template <int LPR, int CH, int U>
static int launch_sddmm_mode(const iddgcn_graph* g, const float* dS, const float* X, const float* vals, int mode,
                             float* dvals, int nvec, cudaStream_t st) {
  const int64_t rows = (int64_t)g->R * g->n_rows;
  constexpr int rows_per_block = (SPMM_THREADS / 32) * (32 / LPR);
  const int64_t blocks = (rows + rows_per_block - 1) / rows_per_block;
  IDD_CHECK(blocks < (int64_t)0x7fffffff, IDDGCN_ELIMIT, "too many rows for one launch");
  const iddgcn_graph::HeavyList& hl = g->heavy_fwd;
  const int heavy_t = hl.H > 0 ? HEAVY_EDGES : 0x7fffffff;
  const unsigned hblocks = (unsigned)((hl.total_segs + rows_per_block - 1) / rows_per_block);
  if (mode == IDDGCN_VALS_SIGMOID) {
    sddmm_rows_kernel<LPR, CH, U, IDDGCN_VALS_SIGMOID><<<(unsigned)blocks, SPMM_THREADS, 0, st>>>(
        g->csr_rowptr, g->csr_col, vals, dS, X, dvals, rows, nvec, heavy_t);
    IDD_LAUNCH_CHECK();
    if (hl.H) {
      sddmm_hub_kernel<LPR, CH, U, IDDGCN_VALS_SIGMOID><<<hblocks, SPMM_THREADS, 0, st>>>(
          g->csr_rowptr, g->csr_col, vals, dS, X, dvals, hl.rows, hl.segptr, hl.H, hl.total_segs, nvec);
      IDD_LAUNCH_CHECK();
    }
  } else {
    sddmm_rows_kernel<LPR, CH, U, IDDGCN_VALS_EXPLICIT><<<(unsigned)blocks, SPMM_THREADS, 0, st>>>(
        g->csr_rowptr, g->csr_col, vals, dS, X, dvals, rows, nvec, heavy_t);
    IDD_LAUNCH_CHECK();
    if (hl.H) {
      sddmm_hub_kernel<LPR, CH, U, IDDGCN_VALS_EXPLICIT><<<hblocks, SPMM_THREADS, 0, st>>>(
          g->csr_rowptr, g->csr_col, vals, dS, X, dvals, hl.rows, hl.segptr, hl.H, hl.total_segs, nvec);
      IDD_LAUNCH_CHECK();
    }
  }
  return IDDGCN_OK;
}

struct StreamArgs {
  const int32_t* ptr; const int32_t* idx; const int32_t* perm; const float* vals; const float* src; float* out;
  int64_t row_begin, row_end; int nvec; AddSrcs add; int heavy_t;
};

template <int LPR, bool USE_PERM>
static int launch_stream_mode(const StreamArgs& a, int mode, cudaStream_t st) {
  const int64_t rows = a.row_end - a.row_begin;
  if (rows <= 0) return IDDGCN_OK;
  constexpr int rows_per_block = (SPMM_THREADS / 32) * 32;  // every warp covers 32 rows (SUBS chunks of LPR)
  const int64_t blocks = (rows + rows_per_block - 1) / rows_per_block;
  IDD_CHECK(blocks < (int64_t)0x7fffffff, IDDGCN_ELIMIT, "too many rows for one launch");
  // blocks per SM of the addend instantiation: the shared-memory variant (narrow rows, see the kernel) needs no extra
  // registers and keeps 4; the register-prefetch variant needs ~70 registers -> 3.  IDDGCN_ADD_MINB=3/4 forces (A/B).
  const char* am = getenv("IDDGCN_ADD_MINB");
  const bool smem_variant = LPR <= 8 && mode == IDDGCN_VALS_ONES && a.heavy_t == 0x7fffffff;
  const int add_minb = (am && (am[0] == '3' || am[0] == '4')) ? (am[0] - '0') : (smem_variant ? 4 : 3);
  // the epilogue addends (accumulate / staged peer partials) exist only on the backward side (USE_PERM) and are a
  // separate instantiation (3 blocks per SM: the prefetched addends need ~70 registers), so the plain kernels keep the
  // register allocation they were tuned with (64 registers, 4 blocks, no spills)
#define IDD_STREAM_ARGS a.ptr, a.idx, a.perm, a.vals, a.src, a.out, a.row_begin, a.row_end, a.nvec, a.add, a.heavy_t
#define IDD_STREAM2(MODE_, HUBS_)                                                                                    \
  do {                                                                                                               \
    if (USE_PERM && a.add.n > 0 && add_minb == 4)                                                                    \
      stream_rows_kernel<LPR, 8, MODE_, USE_PERM, HUBS_, SPMM_THREADS, 4, false, USE_PERM>                           \
          <<<(unsigned)blocks, SPMM_THREADS, 0, st>>>(IDD_STREAM_ARGS);                                              \
    else if (USE_PERM && a.add.n > 0)                                                                                \
      stream_rows_kernel<LPR, 8, MODE_, USE_PERM, HUBS_, SPMM_THREADS, 3, false, USE_PERM>                           \
          <<<(unsigned)blocks, SPMM_THREADS, 0, st>>>(IDD_STREAM_ARGS);                                              \
    else                                                                                                             \
      stream_rows_kernel<LPR, 8, MODE_, USE_PERM, HUBS_><<<(unsigned)blocks, SPMM_THREADS, 0, st>>>(IDD_STREAM_ARGS); \
  } while (0)
#define IDD_STREAM(MODE_)                                       \
  do {                                                          \
    if (a.heavy_t == 0x7fffffff) IDD_STREAM2(MODE_, false);     \
    else IDD_STREAM2(MODE_, true);                              \
  } while (0)
  // IDDGCN_TUNE=<k>: launch-shape variants of the headline instantiation (D=128, unweighted, no hub rows) for A/B runs
  if ((LPR == 32 || LPR == 8) && mode == IDDGCN_VALS_ONES && a.heavy_t == 0x7fffffff && a.add.n == 0) {
    const char* tv = getenv("IDDGCN_TUNE");
    const int k = tv ? atoi(tv) : 0;
    if (k > 0) {
#define IDD_TUNE(U_, TPB_, MINB_, NA_)                                                                              \
      do {                                                                                                          \
        const int64_t nb = (rows + (TPB_ / 32) * 32 - 1) / ((TPB_ / 32) * 32);                                      \
        stream_rows_kernel<(LPR == 8 ? 8 : 32), U_, IDDGCN_VALS_ONES, USE_PERM, false, TPB_, MINB_, NA_>            \
            <<<(unsigned)nb, TPB_, 0, st>>>(                                                                        \
            a.ptr, a.idx, a.perm, a.vals, a.src, a.out, a.row_begin, a.row_end, a.nvec, a.add, a.heavy_t);   \
      } while (0)
      switch (k) {
        case 1: IDD_TUNE(8, 256, 4, true); break;
        case 2: IDD_TUNE(16, 256, 1, false); break;
        case 3: IDD_TUNE(4, 256, 6, false); break;
        case 4: IDD_TUNE(8, 128, 1, false); break;
        case 5: IDD_TUNE(8, 128, 8, false); break;
        case 6: IDD_TUNE(8, 256, 5, false); break;
        case 7: IDD_TUNE(12, 256, 1, false); break;
        case 8: IDD_TUNE(6, 256, 5, false); break;
        case 9: IDD_TUNE(8, 512, 2, false); break;
        case 10: IDD_TUNE(8, 256, 1, false); break;
        case 11: IDD_TUNE(8, 256, 3, false); break;
        case 12: IDD_TUNE(10, 256, 3, false); break;
        case 13: IDD_TUNE(4, 256, 8, false); break;
        case 14: IDD_TUNE(2, 256, 8, false); break;
        default: IDD_TUNE(8, 256, 4, false); break;
      }
#undef IDD_TUNE
      IDD_LAUNCH_CHECK();
      return IDDGCN_OK;
    }
  }
  if (mode == IDDGCN_VALS_ONES) IDD_STREAM(IDDGCN_VALS_ONES);
  else if (mode == IDDGCN_VALS_EXPLICIT) IDD_STREAM(IDDGCN_VALS_EXPLICIT);
  else IDD_STREAM(IDDGCN_VALS_SIGMOID);
#undef IDD_STREAM
#undef IDD_STREAM2
#undef IDD_STREAM_ARGS
  IDD_LAUNCH_CHECK();
  return IDDGCN_OK;
}

template <bool USE_PERM>
static int launch_stream(const StreamArgs& a, int mode, cudaStream_t st) {
  if (a.nvec <= 4) return launch_stream_mode<4, USE_PERM>(a, mode, st);
  if (a.nvec <= 8) return launch_stream_mode<8, USE_PERM>(a, mode, st);
  if (a.nvec <= 16) return launch_stream_mode<16, USE_PERM>(a, mode, st);
  return launch_stream_mode<32, USE_PERM>(a, mode, st);
}

// Kernel choice.  v2 gives every warp 32 consecutive rows, so it needs >= ~64k rows to fill 148 SMs; below that (the
// shipped 845-node graph, explainer sub-graphs) the row-at-a-time v1 kernel exposes 32/LPR times more parallelism and
// is the faster one.  D > 128 always uses v1.  IDDGCN_SPMM_V1=1 / =0 forces v1 / v2 for A/B measurements.
constexpr int64_t V2_MIN_ROWS = 65536;
static bool use_v1(int64_t rows) {
  const char* e = getenv("IDDGCN_SPMM_V1");  // read on every call so tests can flip it at run time
  if (e && (e[0] == '0' || e[0] == '1')) return e[0] == '1';
  return rows < V2_MIN_ROWS;
}

static int check_common(const iddgcn_graph* g, int64_t D, const float* vals, int mode) {
  IDD_CHECK(g != nullptr, IDDGCN_EINVAL, "graph is NULL");
  IDD_CHECK(D > 0 && (D % 4) == 0, IDDGCN_EINVAL, "feature width D=%lld must be a positive multiple of 4",
            (long long)D);
  IDD_CHECK(mode >= IDDGCN_VALS_ONES && mode <= IDDGCN_VALS_SIGMOID, IDDGCN_EINVAL, "bad vals_mode %d", mode);
  IDD_CHECK(mode == IDDGCN_VALS_ONES || vals != nullptr || g->nnz == 0, IDDGCN_EINVAL,
            "vals is NULL but vals_mode=%d", mode);
  return IDDGCN_OK;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- hub-row split pass ---------------------------------------------------------------------
// The workspace of the split pass ([total_segs, D] partial sums) belongs to the graph handle and comes from the pool.
// It is sized by iddgcn_spmm_prepare (called by the model before it captures its train step) or grown here on first
// use; growing is refused while the stream is being captured (an allocation would invalidate the capture).
static int ensure_heavy_ws(const iddgcn_graph* g, int total_segs, int64_t D, cudaStream_t st) {
  const size_t need = (size_t)total_segs * (size_t)D;
  if (need <= g->heavy_ws_elems) return IDDGCN_OK;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone) {
    set_error("hub-row workspace (%zu floats) is not allocated and the stream is being captured: call "
              "iddgcn_spmm_prepare(graph, D) before the capture", need);
    return IDDGCN_EINVAL;
  }
  if (g->heavy_ws) pool_free(g->heavy_ws);
  g->heavy_ws = nullptr;
  g->heavy_ws_elems = 0;
  IDD_CUDA(pool_alloc((void**)&g->heavy_ws, need * sizeof(float)));
  g->heavy_ws_elems = need;
  return IDDGCN_OK;
}

template <int LPR, int CH, int U, bool USE_PERM>
static int launch_heavy_mode(const iddgcn_graph::HeavyList& hl, const int32_t* ptr, const int32_t* idx,
                             const int32_t* perm, const float* vals, int mode, const float* src, float* scratch,
                             int nvec, cudaStream_t st) {
  constexpr int per_block = (SPMM_THREADS / 32) * (32 / LPR);
  const unsigned blocks = (unsigned)((hl.total_segs + per_block - 1) / per_block);
#define IDD_HEAVY(MODE_)                                                                                   \
  heavy_partial_kernel<LPR, CH, U, MODE_, USE_PERM><<<blocks, SPMM_THREADS, 0, st>>>(                      \
      ptr, idx, perm, vals, src, hl.rows, hl.segptr, hl.H, hl.total_segs, scratch, nvec)
  if (mode == IDDGCN_VALS_ONES) IDD_HEAVY(IDDGCN_VALS_ONES);
  else if (mode == IDDGCN_VALS_EXPLICIT) IDD_HEAVY(IDDGCN_VALS_EXPLICIT);
  else IDD_HEAVY(IDDGCN_VALS_SIGMOID);
#undef IDD_HEAVY
  IDD_LAUNCH_CHECK();
  return IDDGCN_OK;
}

template <bool USE_PERM>
static int heavy_pass(const iddgcn_graph* g, const iddgcn_graph::HeavyList& hl, const int32_t* ptr,
                      const int32_t* idx, const int32_t* perm, const float* vals, int mode, const float* src,
                      float* out, int64_t row_begin, int64_t row_end, int64_t D, const AddSrcs& add, cudaStream_t st) {
  if (hl.H == 0) return IDDGCN_OK;
  IDD_TRY(ensure_heavy_ws(g, hl.total_segs, D, st));
  const int nvec = (int)(D / 4);
  if (nvec <= 8) IDD_TRY((launch_heavy_mode<8, 1, 8, USE_PERM>(hl, ptr, idx, perm, vals, mode, src, g->heavy_ws, nvec, st)));
  else if (nvec <= 16) IDD_TRY((launch_heavy_mode<16, 1, 8, USE_PERM>(hl, ptr, idx, perm, vals, mode, src, g->heavy_ws, nvec, st)));
  else if (nvec <= 32) IDD_TRY((launch_heavy_mode<32, 1, 8, USE_PERM>(hl, ptr, idx, perm, vals, mode, src, g->heavy_ws, nvec, st)));
  else if (nvec <= 64) IDD_TRY((launch_heavy_mode<32, 2, 4, USE_PERM>(hl, ptr, idx, perm, vals, mode, src, g->heavy_ws, nvec, st)));
  else IDD_TRY((launch_heavy_mode<32, 4, 2, USE_PERM>(hl, ptr, idx, perm, vals, mode, src, g->heavy_ws, nvec, st)));
  heavy_reduce_kernel<<<hl.H, 128, 0, st>>>(hl.rows, hl.segptr, hl.H, g->heavy_ws, out, row_begin, row_end, nvec, add);
  IDD_LAUNCH_CHECK();
  return IDDGCN_OK;
}

// Both directions are the same generic gather  out[o] = Σ_{e in ptr[o]..ptr[o+1]} w(e)·src[idx[e]] (+ Σ_j add_j[o]):
//   forward  on (csr_rowptr, csr_col)            over R·n_rows rows, weight index e
//   backward on (bwd_rowptr, bwd_src, bwd_perm)  over N rows (node-major CSC view), weight index perm[e]
// `out` and the addends are indexed by the GLOBAL row id o (callers that hold only the rows of a range pass pointers
// biased by -row_begin·D).
template <bool USE_PERM>
static int generic_gather(const iddgcn_graph* g, const iddgcn_graph::HeavyList& hl, const int32_t* ptr,
                          const int32_t* idx, const int32_t* perm, int64_t total_rows, const float* vals, int mode,
                          const float* src, float* out, int64_t D, const AddSrcs& add, int64_t row_begin,
                          int64_t row_end, cudaStream_t st) {
  const int heavy_t = hl.H > 0 ? HEAVY_EDGES : 0x7fffffff;
  if (D <= 128 && !use_v1(total_rows)) {
    StreamArgs a{ptr, idx, perm, vals, src, out, row_begin, row_end, (int)(D / 4), add, heavy_t};
    IDD_TRY(launch_stream<USE_PERM>(a, mode, st));
  } else {
    GatherArgs a{ptr, idx, perm, vals, src, out, row_begin, row_end, 1, 0, 0, (int)(D / 4), add, heavy_t};
    IDD_TRY(launch_gather<USE_PERM>(a, mode, st));
  }
  return heavy_pass<USE_PERM>(g, hl, ptr, idx, perm, vals, mode, src, out, row_begin, row_end, D, add, st);
}

int spmm_fwd_range(const iddgcn_graph* g, const float* X, int64_t D, const float* vals, int mode, float* S,
                   int64_t row_begin, int64_t row_end, cudaStream_t st) {
  return generic_gather<false>(g, g->heavy_fwd, g->csr_rowptr, g->csr_col, nullptr, (int64_t)g->R * g->n_rows, vals,
                               mode, X, S, D, add_none(), row_begin, row_end, st);
}

static int spmm_bwd_x_add(const iddgcn_graph* g, const float* dS, int64_t D, const float* vals, int mode, float* dX,
                          const AddSrcs& add, int64_t row_begin, int64_t row_end, cudaStream_t st) {
  return generic_gather<true>(g, g->heavy_bwd, g->bwd_rowptr, g->bwd_src, g->bwd_perm, g->N, vals, mode, dS, dX, D,
                              add, row_begin, row_end, st);
}

int spmm_bwd_x_range(const iddgcn_graph* g, const float* dS, int64_t D, const float* vals, int mode, float* dX,
                     int accumulate, int64_t row_begin, int64_t row_end, cudaStream_t st) {
  return spmm_bwd_x_add(g, dS, D, vals, mode, dX, accumulate ? add_self(dX) : add_none(), row_begin, row_end, st);
}

int spmm_prepare(const iddgcn_graph* g, int64_t D, cudaStream_t st) {
  const int segs = g->heavy_fwd.total_segs > g->heavy_bwd.total_segs ? g->heavy_fwd.total_segs : g->heavy_bwd.total_segs;
  if (segs == 0) return IDDGCN_OK;
  return ensure_heavy_ws(g, segs, D, st);
}

// v2 (flat edge stream) for graphs large enough to fill the machine with 32-row chunks, v1 (row at a time) otherwise
template <int LPR, int U, int MINB>
static int launch_sddmm_stream(const iddgcn_graph* g, const float* dS, const float* X, const float* vals, int mode,
                               float* dvals, int nvec, cudaStream_t st) {
  const int64_t rows = (int64_t)g->R * g->n_rows;
  constexpr int rows_per_block = (SPMM_THREADS / 32) * 32;
  const int64_t blocks = (rows + rows_per_block - 1) / rows_per_block;
  IDD_CHECK(blocks < (int64_t)0x7fffffff, IDDGCN_ELIMIT, "too many rows for one launch");
  const iddgcn_graph::HeavyList& hl = g->heavy_fwd;
  const int heavy_t = hl.H > 0 ? HEAVY_EDGES : 0x7fffffff;
#define IDD_SDDMM(MODE_, HUBS_)                                                                              \
  sddmm_stream_kernel<LPR, U, MODE_, HUBS_, SPMM_THREADS, MINB><<<(unsigned)blocks, SPMM_THREADS, 0, st>>>( \
      g->csr_rowptr, g->csr_col, vals, dS, X, dvals, rows, nvec, heavy_t)
  if (mode == IDDGCN_VALS_SIGMOID) {
    if (hl.H) IDD_SDDMM(IDDGCN_VALS_SIGMOID, true); else IDD_SDDMM(IDDGCN_VALS_SIGMOID, false);
  } else {
    if (hl.H) IDD_SDDMM(IDDGCN_VALS_EXPLICIT, true); else IDD_SDDMM(IDDGCN_VALS_EXPLICIT, false);
  }
#undef IDD_SDDMM
  IDD_LAUNCH_CHECK();
  if (hl.H) {
    constexpr int per_block = (SPMM_THREADS / 32) * (32 / LPR);
    const unsigned hblocks = (unsigned)((hl.total_segs + per_block - 1) / per_block);
    if (mode == IDDGCN_VALS_SIGMOID)
      sddmm_hub_kernel<LPR, 1, 8, IDDGCN_VALS_SIGMOID><<<hblocks, SPMM_THREADS, 0, st>>>(
          g->csr_rowptr, g->csr_col, vals, dS, X, dvals, hl.rows, hl.segptr, hl.H, hl.total_segs, nvec);
    else
      sddmm_hub_kernel<LPR, 1, 8, IDDGCN_VALS_EXPLICIT><<<hblocks, SPMM_THREADS, 0, st>>>(
          g->csr_rowptr, g->csr_col, vals, dS, X, dvals, hl.rows, hl.segptr, hl.H, hl.total_segs, nvec);
    IDD_LAUNCH_CHECK();
  }
  return IDDGCN_OK;
}

int sddmm_dispatch(const iddgcn_graph* g, const float* dS, const float* X, int64_t D, const float* vals, int mode,
                   float* dvals, cudaStream_t st) {
  if (g->nnz == 0) return IDDGCN_OK;
  const int nvec = (int)(D / 4);
  if (D <= 128 && !use_v1((int64_t)g->R * g->n_rows)) {
    // launch shape measured at config 3 (profiles/r02_sddmm_launch_shape_sweep.txt): occupancy beats batch depth —
    // (U gathers in flight, blocks/SM): (2,8) 2.02 ms, (4,6) 2.35, (4,5) 2.54, (4,4) 2.76, (8,3) 3.55, (8,2) 4.97 ms.
    // 8 blocks x 256 threads is the SM's thread limit (32 registers per thread).
    if (nvec <= 4) return launch_sddmm_stream<4, 2, 8>(g, dS, X, vals, mode, dvals, nvec, st);
    if (nvec <= 8) return launch_sddmm_stream<8, 2, 8>(g, dS, X, vals, mode, dvals, nvec, st);
    if (nvec <= 16) return launch_sddmm_stream<16, 2, 8>(g, dS, X, vals, mode, dvals, nvec, st);
    if (const char* tv = getenv("IDDGCN_SDDMM_TUNE")) {   // launch-shape variants of the D=128 instantiation (A/B runs)
      switch (atoi(tv)) {
        case 1: return launch_sddmm_stream<32, 4, 6>(g, dS, X, vals, mode, dvals, nvec, st);
        case 2: return launch_sddmm_stream<32, 1, 8>(g, dS, X, vals, mode, dvals, nvec, st);
        case 3: return launch_sddmm_stream<32, 4, 4>(g, dS, X, vals, mode, dvals, nvec, st);
        case 4: return launch_sddmm_stream<32, 8, 3>(g, dS, X, vals, mode, dvals, nvec, st);
        default: break;
      }
    }
    return launch_sddmm_stream<32, 2, 8>(g, dS, X, vals, mode, dvals, nvec, st);
  }
  if (nvec <= 8) return launch_sddmm_mode<8, 1, 8>(g, dS, X, vals, mode, dvals, nvec, st);
  if (nvec <= 16) return launch_sddmm_mode<16, 1, 8>(g, dS, X, vals, mode, dvals, nvec, st);
  if (nvec <= 32) return launch_sddmm_mode<32, 1, 8>(g, dS, X, vals, mode, dvals, nvec, st);
  if (nvec <= 64) return launch_sddmm_mode<32, 2, 4>(g, dS, X, vals, mode, dvals, nvec, st);
  if (nvec <= 128) return launch_sddmm_mode<32, 4, 2>(g, dS, X, vals, mode, dvals, nvec, st);
  set_error("feature width D=%lld > 512 is not supported", (long long)D);
  return IDDGCN_ELIMIT;
}

}  // namespace iddgcn

using namespace iddgcn;

extern "C" {

int iddgcn_spmm_fwd(const iddgcn_graph_t* g, const float* X, int64_t D, const float* vals, int vals_mode, float* S,
                    void* stream) {
  IDD_TRY(check_common(g, D, vals, vals_mode));
  if (g->n_rows == 0) return IDDGCN_OK;   // empty row block of a partition: nothing to write
  IDD_CHECK(X && S, IDDGCN_EINVAL, "X or S is NULL");
  IDD_CHECK(aligned16(X) && aligned16(S), IDDGCN_EINVAL, "X and S must be 16-byte aligned");
  return spmm_fwd_range(g, X, D, vals, vals_mode, S, 0, (int64_t)g->R * g->n_rows, (cudaStream_t)stream);
}

int iddgcn_spmm_bwd_x(const iddgcn_graph_t* g, const float* dS, int64_t D, const float* vals, int vals_mode,
                      float* dX, int accumulate, void* stream) {
  IDD_TRY(check_common(g, D, vals, vals_mode));
  IDD_CHECK((dS || g->nnz == 0) && dX, IDDGCN_EINVAL, "dS or dX is NULL");   // an empty row block has no dS rows
  IDD_CHECK(aligned16(dS) && aligned16(dX), IDDGCN_EINVAL, "dS and dX must be 16-byte aligned");
  return spmm_bwd_x_range(g, dS, D, vals, vals_mode, dX, accumulate, 0, g->N, (cudaStream_t)stream);
}

int iddgcn_spmm_prepare(const iddgcn_graph_t* g, int64_t D, void* stream) {
  IDD_TRY(check_common(g, D, nullptr, IDDGCN_VALS_ONES));
  return spmm_prepare(g, D, (cudaStream_t)stream);
}

int iddgcn_spmm_fwd_rows(const iddgcn_graph_t* g, const float* X, int64_t D, const float* vals, int vals_mode,
                         float* S_rows, int64_t row_begin, int64_t row_end, void* stream) {
  IDD_TRY(check_common(g, D, vals, vals_mode));
  if (row_begin == row_end) return IDDGCN_OK;
  IDD_CHECK(X && S_rows, IDDGCN_EINVAL, "X or S_rows is NULL");
  IDD_CHECK(aligned16(X) && aligned16(S_rows), IDDGCN_EINVAL, "X and S_rows must be 16-byte aligned");
  IDD_CHECK(0 <= row_begin && row_begin <= row_end && row_end <= (int64_t)g->R * g->n_rows, IDDGCN_EINVAL,
            "row range [%lld, %lld) outside [0, R*n_rows)", (long long)row_begin, (long long)row_end);
  return spmm_fwd_range(g, X, D, vals, vals_mode, S_rows - (size_t)row_begin * D, row_begin, row_end,
                        (cudaStream_t)stream);
}

int iddgcn_spmm_bwd_x_rows(const iddgcn_graph_t* g, const float* dS, int64_t D, const float* vals, int vals_mode,
                           float* dX_rows, int64_t node_begin, int64_t node_end, const float* const* addends,
                           int n_addends, void* stream) {
  IDD_TRY(check_common(g, D, vals, vals_mode));
  if (node_begin == node_end) return IDDGCN_OK;
  IDD_CHECK((dS || g->nnz == 0) && dX_rows, IDDGCN_EINVAL, "dS or dX_rows is NULL");
  IDD_CHECK(aligned16(dS) && aligned16(dX_rows), IDDGCN_EINVAL, "dS and dX_rows must be 16-byte aligned");
  IDD_CHECK(0 <= node_begin && node_begin <= node_end && node_end <= g->N, IDDGCN_EINVAL,
            "node range [%lld, %lld) outside [0, N)", (long long)node_begin, (long long)node_end);
  IDD_CHECK(n_addends >= 0 && n_addends <= MAX_ADD && (n_addends == 0 || addends != nullptr), IDDGCN_EINVAL,
            "n_addends must be in [0, %d]", MAX_ADD);
  AddSrcs add = add_none();
  for (int j = 0; j < n_addends; ++j) {
    IDD_CHECK(addends[j] != nullptr && aligned16(addends[j]), IDDGCN_EINVAL, "addend %d is NULL or misaligned", j);
    add.p[j] = addends[j] - (size_t)node_begin * D;
  }
  add.n = n_addends;
  return spmm_bwd_x_add(g, dS, D, vals, vals_mode, dX_rows - (size_t)node_begin * D, add, node_begin, node_end,
                        (cudaStream_t)stream);
}

int iddgcn_spmm_bwd_vals(const iddgcn_graph_t* g, const float* dS, const float* X, int64_t D, const float* vals,
                         int vals_mode, float* dvals, void* stream) {
  IDD_TRY(check_common(g, D, vals, vals_mode));
  IDD_CHECK(dS && X && dvals, IDDGCN_EINVAL, "dS, X or dvals is NULL");
  IDD_CHECK(aligned16(dS) && aligned16(X), IDDGCN_EINVAL, "dS and X must be 16-byte aligned");
  return sddmm_dispatch(g, dS, X, D, vals, vals_mode, dvals, (cudaStream_t)stream);
}

int iddgcn_edge_value_sums(const iddgcn_graph_t* g, const float* vals, int vals_mode, float* out_per_rel,
                           void* stream) {
  IDD_CHECK(g != nullptr && out_per_rel != nullptr, IDDGCN_EINVAL, "graph or out is NULL");
  IDD_CHECK(vals_mode >= IDDGCN_VALS_ONES && vals_mode <= IDDGCN_VALS_SIGMOID, IDDGCN_EINVAL, "bad vals_mode");
  IDD_CHECK(vals_mode == IDDGCN_VALS_ONES || vals != nullptr || g->nnz == 0, IDDGCN_EINVAL, "vals is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  if (vals_mode == IDDGCN_VALS_ONES)
    edge_value_sums_kernel<IDDGCN_VALS_ONES><<<g->R, 256, 0, st>>>(g->csr_rowptr, g->n_rows, vals, out_per_rel);
  else if (vals_mode == IDDGCN_VALS_EXPLICIT)
    edge_value_sums_kernel<IDDGCN_VALS_EXPLICIT><<<g->R, 256, 0, st>>>(g->csr_rowptr, g->n_rows, vals, out_per_rel);
  else
    edge_value_sums_kernel<IDDGCN_VALS_SIGMOID><<<g->R, 256, 0, st>>>(g->csr_rowptr, g->n_rows, vals, out_per_rel);
  IDD_LAUNCH_CHECK();
  return IDDGCN_OK;
}

// ---- host-buffer end-to-end path ------------------------------------------------------------
// The step is PCIe-bound (X + dS in, S + dX out: 3 GB at config 3 against ~3 ms of kernels), so what matters is that both
// directions of the link stay busy.  Feature columns are independent in an SpMM — S[:, c] needs X[:, c] only and
// dX[:, c] needs dS[:, c] only — so the buffers travel as P column panels (strided 2-D copies; on the device every panel
// is a dense [rows, D/P] array the kernels run on directly): S panels start flowing back after the FIRST X panel has
// landed instead of after all of X, and the last dX panel is 1/P of dX instead of all of it.
static int host_panels(int64_t D) {
  // measured at config 3 (D = 128, profiles/README.md): 2 panels of 256-byte row segments 37.3 ms, 1 panel 38.8 ms, 4 panels
  // of 128-byte segments 49.0 ms — the copy engines lose their rate below 256 bytes per row, so never go narrower
  int p = D >= 128 ? 2 : 1;
  if (const char* e = getenv("IDDGCN_HOST_PANELS")) {   // A/B runs
    const int v = atoi(e);
    if (v == 1 || v == 2 || v == 4) p = v;
  }
  while (p > 1 && (D % (4 * p)) != 0) p >>= 1;
  return p;
}

int iddgcn_spmm_fwd_bwd_host(const iddgcn_graph_t* g, const float* X_host, const float* dS_host, int64_t D,
                             float* S_host, float* dX_host) {
  IDD_TRY(check_common(g, D, nullptr, IDDGCN_VALS_ONES));
  IDD_CHECK(X_host && dS_host, IDDGCN_EINVAL, "X_host or dS_host is NULL");
  const int64_t srows = (int64_t)g->R * g->n_rows;
  const size_t xN = (size_t)g->N * D, sN = (size_t)srows * D;
  const size_t need = 2 * xN + 2 * sN;
  if (g->host_ws_elems < need) {
    if (g->host_ws) pool_free(g->host_ws);
    g->host_ws = nullptr;
    g->host_ws_elems = 0;
    IDD_CUDA(pool_alloc((void**)&g->host_ws, need * sizeof(float)));
    g->host_ws_elems = need;
  }
  float* X_dev = g->host_ws;                 // [P][N][Dp]
  float* dS_dev = X_dev + xN;                // [P][srows][Dp]
  float* S_dev = dS_dev + sN;                // [P][srows][Dp]
  float* dX_dev = S_dev + sN;                // [P][N][Dp]
  for (int i = 0; i < 3; ++i)
    if (!g->host_streams[i]) IDD_CUDA(cudaStreamCreateWithFlags(&g->host_streams[i], cudaStreamNonBlocking));
  for (int i = 0; i < 16; ++i)
    if (!g->host_events[i]) IDD_CUDA(cudaEventCreateWithFlags(&g->host_events[i], cudaEventDisableTiming));
  cudaStream_t s_in = g->host_streams[0], s_cmp = g->host_streams[1], s_out = g->host_streams[2];
  const int P = host_panels(D);
  const int64_t Dp = D / P;
  const size_t wbytes = (size_t)Dp * 4, hpitch = (size_t)D * 4;
  auto ev = [&](int kind, int c) { return g->host_events[kind * 4 + c]; };   // 0: X panel c landed, 1: dS panel c landed
  auto body = [&]() -> int {
    for (int c = 0; c < P; ++c) {
      IDD_CUDA(cudaMemcpy2DAsync(X_dev + (size_t)c * g->N * Dp, wbytes, X_host + c * Dp, hpitch, wbytes, (size_t)g->N,
                                 cudaMemcpyHostToDevice, s_in));
      IDD_CUDA(cudaEventRecord(ev(0, c), s_in));
    }
    for (int c = 0; c < P; ++c) {
      IDD_CUDA(cudaMemcpy2DAsync(dS_dev + (size_t)c * srows * Dp, wbytes, dS_host + c * Dp, hpitch, wbytes,
                                 (size_t)srows, cudaMemcpyHostToDevice, s_in));
      IDD_CUDA(cudaEventRecord(ev(1, c), s_in));
    }
    // inside a panel the output rows are cut into chunks so that the copy back starts while the kernel still runs
    const int NCHUNK = 8 / P;
    cudaEvent_t ev_chunk = g->host_events[15];
    for (int c = 0; c < P; ++c) {
      IDD_CUDA(cudaStreamWaitEvent(s_cmp, ev(0, c), 0));
      float* Sp = S_dev + (size_t)c * srows * Dp;
      for (int q = 0; q < NCHUNK; ++q) {
        const int64_t lo = srows * q / NCHUNK, hi = srows * (q + 1) / NCHUNK;
        IDD_TRY(spmm_fwd_range(g, X_dev + (size_t)c * g->N * Dp, Dp, nullptr, IDDGCN_VALS_ONES, Sp, lo, hi, s_cmp));
        if (S_host && hi > lo) {
          IDD_CUDA(cudaEventRecord(ev_chunk, s_cmp));
          IDD_CUDA(cudaStreamWaitEvent(s_out, ev_chunk, 0));
          IDD_CUDA(cudaMemcpy2DAsync(S_host + (size_t)lo * D + c * Dp, hpitch, Sp + (size_t)lo * Dp, wbytes, wbytes,
                                     (size_t)(hi - lo), cudaMemcpyDeviceToHost, s_out));
        }
      }
    }
    for (int c = 0; c < P; ++c) {
      IDD_CUDA(cudaStreamWaitEvent(s_cmp, ev(1, c), 0));
      float* dXp = dX_dev + (size_t)c * g->N * Dp;
      for (int q = 0; q < NCHUNK; ++q) {
        const int64_t lo = g->N * q / NCHUNK, hi = g->N * (q + 1) / NCHUNK;
        IDD_TRY(spmm_bwd_x_range(g, dS_dev + (size_t)c * srows * Dp, Dp, nullptr, IDDGCN_VALS_ONES, dXp, 0, lo, hi, s_cmp));
        if (dX_host && hi > lo) {
          IDD_CUDA(cudaEventRecord(ev_chunk, s_cmp));
          IDD_CUDA(cudaStreamWaitEvent(s_out, ev_chunk, 0));
          IDD_CUDA(cudaMemcpy2DAsync(dX_host + (size_t)lo * D + c * Dp, hpitch, dXp + (size_t)lo * Dp, wbytes, wbytes,
                                     (size_t)(hi - lo), cudaMemcpyDeviceToHost, s_out));
        }
      }
    }
    IDD_CUDA(cudaStreamSynchronize(s_cmp));
    IDD_CUDA(cudaStreamSynchronize(s_out));
    IDD_CUDA(cudaStreamSynchronize(s_in));
    return IDDGCN_OK;
  };
  const int rc = body();
  if (rc != IDDGCN_OK) cudaDeviceSynchronize();
  return rc;
}

// ---- DLPack flavour (argument unpacking: dlpack.cuh) ----------------------------------------
int iddgcn_spmm_fwd_dl(const iddgcn_graph_t* g, void* X_dl, void* vals_dl, int vals_mode, void* S_dl, void* stream) {
  IDD_CHECK(g != nullptr, IDDGCN_EINVAL, "graph is NULL");
  float *X, *S, *vals = nullptr;
  int64_t nx, ns, nv, dx, ds, dv;
  IDD_TRY(dl_f32_cuda(X_dl, "X", g->device, &X, &nx, &dx));
  IDD_TRY(dl_f32_cuda(S_dl, "S", g->device, &S, &ns, &ds));
  IDD_CHECK(dx == ds, IDDGCN_EINVAL, "X and S disagree on the feature width (%lld vs %lld)", (long long)dx,
            (long long)ds);
  IDD_CHECK(nx == g->N * dx, IDDGCN_EINVAL, "X must be [N=%lld, D]", (long long)g->N);
  IDD_CHECK(ns == (int64_t)g->R * g->n_rows * dx, IDDGCN_EINVAL, "S must be [R, n_rows, D]");
  if (vals_dl) {
    IDD_TRY(dl_f32_cuda(vals_dl, "vals", g->device, &vals, &nv, &dv));
    IDD_CHECK(nv == g->nnz, IDDGCN_EINVAL, "vals must have nnz=%lld elements", (long long)g->nnz);
  }
  return iddgcn_spmm_fwd(g, X, dx, vals, vals_mode, S, stream);
}

int iddgcn_spmm_bwd_x_dl(const iddgcn_graph_t* g, void* dS_dl, void* vals_dl, int vals_mode, void* dX_dl,
                         int accumulate, void* stream) {
  IDD_CHECK(g != nullptr, IDDGCN_EINVAL, "graph is NULL");
  float *dS, *dX, *vals = nullptr;
  int64_t nx, ns, nv, dx, ds, dv;
  IDD_TRY(dl_f32_cuda(dS_dl, "dS", g->device, &dS, &ns, &ds));
  IDD_TRY(dl_f32_cuda(dX_dl, "dX", g->device, &dX, &nx, &dx));
  IDD_CHECK(dx == ds, IDDGCN_EINVAL, "dS and dX disagree on the feature width");
  IDD_CHECK(nx == g->N * dx, IDDGCN_EINVAL, "dX must be [N=%lld, D]", (long long)g->N);
  IDD_CHECK(ns == (int64_t)g->R * g->n_rows * dx, IDDGCN_EINVAL, "dS must be [R, n_rows, D]");
  if (vals_dl) {
    IDD_TRY(dl_f32_cuda(vals_dl, "vals", g->device, &vals, &nv, &dv));
    IDD_CHECK(nv == g->nnz, IDDGCN_EINVAL, "vals must have nnz=%lld elements", (long long)g->nnz);
  }
  return iddgcn_spmm_bwd_x(g, dS, dx, vals, vals_mode, dX, accumulate, stream);
}

}  // extern "C"
