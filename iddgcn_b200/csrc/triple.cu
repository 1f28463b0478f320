// triple.cu — per-triple stage of IDDGCN_Layer.call (prediction/IDDGCN.py:60-79), DistMult (:103-109),
// the BCE loss of train_step (:161-168) and Keras Adam (:174), forward and backward.
//
// Layout: every per-triple activation is a (2, T, D) buffer — plane 0 = head side, plane 1 = tail side —
// so the two self-kernel matmuls of a layer are ONE GEMM with M = 2T.  One warp owns one triple; each lane
// keeps NV float4 slices of the D-wide row (D <= 128*NV).  Gather backward passes are segment sums over
// the plan's by-head / by-tail orderings: no atomics, fixed summation order.
#include <stdlib.h>
#include "triple.cuh"

namespace iddgcn {

constexpr int TW = 8;  // warps per block
constexpr int SS_U = 4;  // row gathers in flight per lane in the segment sums

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void kahan_add(float& sum, float& comp, float x) {
  const float y = x - comp;
  const float t = sum + y;
  comp = (t - sum) - y;
  sum = t;
}
__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float f4_dot(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ float4 f4_fma(float s, float4 a, float4 c) {
  return make_float4(fmaf(s, a.x, c.x), fmaf(s, a.y, c.y), fmaf(s, a.z, c.z), fmaf(s, a.w, c.w));
}
__device__ __forceinline__ float4 f4_sigmoid(float4 a) {
  return make_float4(sigmoidf_(a.x), sigmoidf_(a.y), sigmoidf_(a.z), sigmoidf_(a.w));
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gather_emb_kernel(const int32_t* __restrict__ head,
                                                         const int32_t* __restrict__ tail,
                                                         const float4* __restrict__ emb, int64_t T, int nvec,
                                                         float4* __restrict__ HT) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = 2 * T * nvec;
  if (i >= total) return;
  const int v = (int)(i % nvec);
  const int64_t row = i / nvec;  // [0, 2T)
  const int32_t n = row < T ? head[row] : tail[row - T];
  HT[i] = __ldg(emb + (int64_t)n * nvec + v);
}

int gather_embeddings(const iddgcn_plan* p, const float* emb, int64_t D, float* HT, cudaStream_t st) {
  if (p->T == 0) return IDDGCN_OK;
  const int nvec = (int)(D / 4);
  const int64_t total = 2 * p->T * nvec;
  gather_emb_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
      p->head, p->tail, reinterpret_cast<const float4*>(emb), p->T, nvec, reinterpret_cast<float4*>(HT));
  IDD_LAUNCH_CHECK();
  return IDDGCN_OK;
}

// ---------------------------------------------------------------------------------------------
// forward epilogue: gate + gathers of Z_i[head], Z_i[tail] + sigmoid
template <int NV, int RPRE>   // RPRE > 0: rows of up to RPRE relations are prefetched (NV == 1 only)
__global__ void __launch_bounds__(TW * 32)
    layer_combine_kernel(const int32_t* __restrict__ head, const int32_t* __restrict__ tail,
                         const float4* __restrict__ Hin, const float4* __restrict__ Z,
                         const float* __restrict__ W_alpha, const float* __restrict__ b_alpha, int64_t T, int64_t N,
                         int nvec, int R, float4* __restrict__ P, float* __restrict__ alpha_out) {
  extern __shared__ float s_wa[];  // W_alpha [D][R]
  const int D = nvec * 4;
  for (int i = threadIdx.x; i < D * R; i += blockDim.x) s_wa[i] = W_alpha[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t k = (int64_t)blockIdx.x * TW + (threadIdx.x >> 5);
  if (k >= T) return;
  // D <= 128: every row this triple touches (its two pre-activation rows and the 2R gathered Z rows) is requested up
  // front, so those 2 + 2R independent 512-byte loads are in flight while the gate (a dot product, a warp reduction
  // and a softmax on the Hin row) is computed; the arithmetic below is unchanged.
  constexpr bool PRE = RPRE > 0 && NV == 1;
  float4 pre_p[PRE ? 2 : 1], pre_z[PRE ? 2 : 1][PRE ? RPRE : 1];
  const int32_t nh0 = head[k], nt0 = tail[k];
  if (PRE && lane < nvec) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int64_t node = half == 0 ? nh0 : nt0;
      pre_p[half] = P[((int64_t)half * T + k) * nvec + lane];
#pragma unroll
      for (int i = 0; i < (PRE ? RPRE : 1); ++i)
        if (i < R) pre_z[half][i] = __ldg(Z + ((int64_t)i * N + node) * nvec + lane);
    }
  }
  // logits z_i = Hin[k]·W_alpha[:, i] + b_i
  float z[RMAX];
#pragma unroll
  for (int i = 0; i < RMAX; ++i) z[i] = 0.f;
#pragma unroll
  for (int c = 0; c < NV; ++c) {
    const int v = lane + c * 32;
    if (v < nvec) {
      const float4 h = __ldg(Hin + k * nvec + v);
      const float hv[4] = {h.x, h.y, h.z, h.w};
      if (R == 2) {
        // the lane's 4 x R weights are contiguous in W_alpha [D][R]: two 128-bit shared loads instead of eight scalar
        // ones with an 8-way bank conflict (lane stride 8 floats); same fma order per relation as the loop below
        const float4 w01 = reinterpret_cast<const float4*>(s_wa)[v * 2];
        const float4 w23 = reinterpret_cast<const float4*>(s_wa)[v * 2 + 1];
        z[0] = fmaf(hv[3], w23.z, fmaf(hv[2], w23.x, fmaf(hv[1], w01.z, fmaf(hv[0], w01.x, z[0]))));
        z[1] = fmaf(hv[3], w23.w, fmaf(hv[2], w23.y, fmaf(hv[1], w01.w, fmaf(hv[0], w01.y, z[1]))));
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float* w = s_wa + (v * 4 + q) * R;
#pragma unroll
          for (int i = 0; i < RMAX; ++i)
            if (i < R) z[i] = fmaf(hv[q], w[i], z[i]);
        }
      }
    }
  }
  float zmax = -INFINITY;
#pragma unroll
  for (int i = 0; i < RMAX; ++i)
    if (i < R) {
      z[i] = warp_sum(z[i]) + b_alpha[i];
      zmax = fmaxf(zmax, z[i]);
    }
  float den = 0.f;
#pragma unroll
  for (int i = 0; i < RMAX; ++i)
    if (i < R) {
      z[i] = expf(z[i] - zmax);
      den += z[i];
    }
  float gate[RMAX];
#pragma unroll
  for (int i = 0; i < RMAX; ++i)
    if (i < R) {
      const float a = z[i] / den;  // softmax (IDDGCN.py:66)
      if (alpha_out && lane == 0) alpha_out[k * R + i] = a;
      gate[i] = sigmoidf_(a);      // relation_weight (IDDGCN.py:75)
    }
  const int32_t nh = nh0, nt = nt0;
  if (PRE) {
    if (lane < nvec) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float4 acc = pre_p[half];
#pragma unroll
        for (int i = 0; i < (PRE ? RPRE : 1); ++i)
          if (i < R) acc = f4_fma(gate[i], pre_z[half][i], acc);
        P[((int64_t)half * T + k) * nvec + lane] = f4_sigmoid(acc);
      }
    }
    return;
  }
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int64_t node = half == 0 ? nh : nt;
    float4* prow = P + ((int64_t)half * T + k) * nvec;
#pragma unroll
    for (int c = 0; c < NV; ++c) {
      const int v = lane + c * 32;
      if (v < nvec) {
        float4 acc = prow[v];
#pragma unroll
        for (int i = 0; i < RMAX; ++i)
          if (i < R) acc = f4_fma(gate[i], __ldg(Z + ((int64_t)i * N + node) * nvec + v), acc);
        prow[v] = f4_sigmoid(acc);
      }
    }
  }
}

int layer_combine(const iddgcn_plan* p, const float* Hin, const float* Z, const float* W_alpha, const float* b_alpha,
                  int64_t D, float* P_inout, float* alpha_out, cudaStream_t st) {
  if (p->T == 0) return IDDGCN_OK;
  IDD_CHECK(p->R <= RMAX, IDDGCN_ELIMIT, "triple stage supports at most %d relations (got %d)", RMAX, p->R);
  IDD_CHECK(D % 4 == 0 && D <= 512, IDDGCN_ELIMIT, "triple stage needs D %% 4 == 0 and D <= 512 (got %lld)",
            (long long)D);
  const int nvec = (int)(D / 4);
  const unsigned grid = (unsigned)((p->T + TW - 1) / TW);
  const size_t smem = (size_t)D * p->R * sizeof(float);
#define IDD_ARGS p->head, p->tail, reinterpret_cast<const float4*>(Hin), reinterpret_cast<const float4*>(Z), W_alpha, \
                 b_alpha, p->T, p->N, nvec, p->R, reinterpret_cast<float4*>(P_inout), alpha_out
  if (nvec <= 32 && p->R <= 2) layer_combine_kernel<1, 2><<<grid, TW * 32, smem, st>>>(IDD_ARGS);
  else if (nvec <= 32 && p->R <= 4) layer_combine_kernel<1, 4><<<grid, TW * 32, smem, st>>>(IDD_ARGS);
  else if (nvec <= 32) layer_combine_kernel<1, 0><<<grid, TW * 32, smem, st>>>(IDD_ARGS);
  else if (nvec <= 64) layer_combine_kernel<2, 0><<<grid, TW * 32, smem, st>>>(IDD_ARGS);
  else layer_combine_kernel<4, 0><<<grid, TW * 32, smem, st>>>(IDD_ARGS);
#undef IDD_ARGS
  IDD_LAUNCH_CHECK();
  return IDDGCN_OK;
}

// ---------------------------------------------------------------------------------------------
// backward, part 1: through the sigmoid, the gate and the softmax (SURVEY App. A eq. 5-7)
template <int NV>
__global__ void __launch_bounds__(TW * 32)
    layer_bwd_pre_kernel(const int32_t* __restrict__ head, const int32_t* __restrict__ tail,
                         const float4* __restrict__ HT_out, const float* __restrict__ alpha,
                         const float4* __restrict__ Z, int64_t T, int64_t N, int nvec, int R,
                         float4* __restrict__ dHT, float* __restrict__ dz) {
  const int lane = threadIdx.x & 31;
  const int64_t k = (int64_t)blockIdx.x * TW + (threadIdx.x >> 5);
  if (k >= T) return;
  float dgate[RMAX];
#pragma unroll
  for (int i = 0; i < RMAX; ++i) dgate[i] = 0.f;
  const int32_t nh = head[k], nt = tail[k];
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int64_t node = half == 0 ? nh : nt;
    const int64_t off = ((int64_t)half * T + k) * nvec;
#pragma unroll
    for (int c = 0; c < NV; ++c) {
      const int v = lane + c * 32;
      if (v < nvec) {
        const float4 o = __ldg(HT_out + off + v);
        const float4 g = dHT[off + v];
        const float4 dp = make_float4(g.x * o.x * (1.f - o.x), g.y * o.y * (1.f - o.y), g.z * o.z * (1.f - o.z),
                                      g.w * o.w * (1.f - o.w));  // eq. 5
        dHT[off + v] = dp;
#pragma unroll
        for (int i = 0; i < RMAX; ++i)
          if (i < R) dgate[i] += f4_dot(dp, __ldg(Z + ((int64_t)i * N + node) * nvec + v));  // eq. 6
      }
    }
  }
  float dal[RMAX], al[RMAX];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < RMAX; ++i)
    if (i < R) {
      const float dg = warp_sum(dgate[i]);
      al[i] = alpha[k * R + i];
      const float gt = sigmoidf_(al[i]);
      dal[i] = dg * gt * (1.f - gt);
      s += dal[i] * al[i];
    }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < RMAX; ++i)
      if (i < R) dz[k * R + i] = al[i] * (dal[i] - s);  // eq. 7
  }
}

int layer_bwd_pre(const iddgcn_plan* p, const float* HT_out, const float* alpha, const float* Z, int64_t D,
                  float* dHT_inout, float* dz, cudaStream_t st) {
  if (p->T == 0) return IDDGCN_OK;
  IDD_CHECK(p->R <= RMAX && D % 4 == 0 && D <= 512, IDDGCN_ELIMIT, "unsupported R=%d / D=%lld", p->R, (long long)D);
  const int nvec = (int)(D / 4);
  const unsigned grid = (unsigned)((p->T + TW - 1) / TW);
#define IDD_ARGS p->head, p->tail, reinterpret_cast<const float4*>(HT_out), alpha, reinterpret_cast<const float4*>(Z), \
                 p->T, p->N, nvec, p->R, reinterpret_cast<float4*>(dHT_inout), dz
  if (nvec <= 32) layer_bwd_pre_kernel<1><<<grid, TW * 32, 0, st>>>(IDD_ARGS);
  else if (nvec <= 64) layer_bwd_pre_kernel<2><<<grid, TW * 32, 0, st>>>(IDD_ARGS);
  else layer_bwd_pre_kernel<4><<<grid, TW * 32, 0, st>>>(IDD_ARGS);
#undef IDD_ARGS
  IDD_LAUNCH_CHECK();
  return IDDGCN_OK;
}

// backward, part 2: dHin += dz · W_alphaᵀ (second term of eq. 9)
__global__ void __launch_bounds__(256) layer_bwd_post_kernel(const float* __restrict__ dz,
                                                             const float* __restrict__ W_alpha, int64_t T, int D,
                                                             int R, float* __restrict__ dHin) {
  // one thread per float4 of a row; same per-element arithmetic and order as a scalar loop over r
  const int nvec = D >> 2;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T * nvec) return;
  const int64_t k = i / nvec;
  const int d = (int)(i - k * nvec) * 4;
  float4* p = reinterpret_cast<float4*>(dHin) + i;
  float4 acc = *p;
  if (R == 2) {
    // the 4 x R weights of this float4 are contiguous in W_alpha [D][R]: two 128-bit loads instead of eight scalar
    // ones; same products, same order (r = 0 then r = 1 per element)
    const float2 z = *reinterpret_cast<const float2*>(dz + k * 2);
    const float4 w01 = __ldg(reinterpret_cast<const float4*>(W_alpha + (size_t)d * 2));
    const float4 w23 = __ldg(reinterpret_cast<const float4*>(W_alpha + (size_t)d * 2) + 1);
    acc.x = fmaf(z.y, w01.y, fmaf(z.x, w01.x, acc.x));
    acc.y = fmaf(z.y, w01.w, fmaf(z.x, w01.z, acc.y));
    acc.z = fmaf(z.y, w23.y, fmaf(z.x, w23.x, acc.z));
    acc.w = fmaf(z.y, w23.w, fmaf(z.x, w23.z, acc.w));
    *p = acc;
    return;
  }
  for (int r = 0; r < R; ++r) {
    const float z = dz[k * R + r];
    acc.x = fmaf(z, __ldg(W_alpha + (d + 0) * R + r), acc.x);
    acc.y = fmaf(z, __ldg(W_alpha + (d + 1) * R + r), acc.y);
    acc.z = fmaf(z, __ldg(W_alpha + (d + 2) * R + r), acc.z);
    acc.w = fmaf(z, __ldg(W_alpha + (d + 3) * R + r), acc.w);
  }
  *p = acc;
}

int layer_bwd_post(const float* dz, const float* W_alpha, int64_t T, int64_t D, int R, float* dHin, cudaStream_t st) {
  if (T == 0) return IDDGCN_OK;
  IDD_CHECK(D % 4 == 0, IDDGCN_ELIMIT, "unsupported D=%lld", (long long)D);
  layer_bwd_post_kernel<<<(unsigned)((T * (D / 4) + 255) / 256), 256, 0, st>>>(dz, W_alpha, T, (int)D, R, dHin);
  IDD_LAUNCH_CHECK();
  return IDDGCN_OK;
}

// ---------------------------------------------------------------------------------------------
// segment sums over the plan's by-head / by-tail orderings (gather backward without atomics)
template <int NV, bool GATED, int RCMAX>
__global__ void __launch_bounds__(TW * 32)
    segment_scatter_kernel(const int32_t* __restrict__ ptr_h, const int32_t* __restrict__ perm_h,
                           const int32_t* __restrict__ ptr_t, const int32_t* __restrict__ perm_t,
                           const float4* __restrict__ dP, const float* __restrict__ alpha, int64_t T, int64_t N,
                           int nvec, int R, float4* __restrict__ out, int accumulate) {
  const int lane = threadIdx.x & 31;
  // relations accumulated per sweep over the node's triples (fewer at large D to bound registers).  blockIdx.y picks
  // a group of RC relations: with RCMAX = 1 (small N) every (node, relation) pair gets its own warp.
  constexpr int RC = GATED ? (RCMAX < (NV >= 4 ? 2 : 4) ? RCMAX : (NV >= 4 ? 2 : 4)) : 1;
  const int64_t n = (int64_t)blockIdx.x * TW + (threadIdx.x >> 5);
  if (n >= N) return;
  const int planes = GATED ? R : 1;
  const int r_first = blockIdx.y * RC * (int)((planes + RC * gridDim.y - 1) / (RC * gridDim.y));
  const int r_last = (blockIdx.y + 1) * RC * (int)((planes + RC * gridDim.y - 1) / (RC * gridDim.y));
  for (int r0 = r_first; r0 < planes && r0 < r_last; r0 += RC) {
    // Kahan-compensated sums: a hub node can own thousands of triples, and a plain sequential fp32 sum of that
    // many terms is what would break the 1e-5 gradient tolerance.
    float4 acc[RC][NV], cmp[RC][NV];
#pragma unroll
    for (int i = 0; i < RC; ++i)
#pragma unroll
      for (int c = 0; c < NV; ++c) { acc[i][c] = f4_zero(); cmp[i][c] = f4_zero(); }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int32_t* ptr = half == 0 ? ptr_h : ptr_t;
      const int32_t* perm = half == 0 ? perm_h : perm_t;
      const int beg = ptr[n], end = ptr[n + 1];
      // 32 triples per batch: lane l fetches triple id and gate weights of entry j0+l, rows are then gathered
      // SS_U at a time (independent 128-bit loads in flight) and accumulated in list order.
      for (int j0 = beg; j0 < end; j0 += 32) {
        const int cnt = (end - j0 < 32) ? (end - j0) : 32;
        int my_k = 0;
        float my_w[RC];
#pragma unroll
        for (int i = 0; i < RC; ++i) my_w[i] = GATED ? 0.f : 1.f;
        if (lane < cnt) {
          my_k = perm[j0 + lane];
          if (GATED) {
#pragma unroll
            for (int i = 0; i < RC; ++i)
              if (r0 + i < R) my_w[i] = sigmoidf_(__ldg(alpha + (int64_t)my_k * R + r0 + i));
          }
        }
        for (int u0 = 0; u0 < cnt; u0 += SS_U) {
          float4 g[SS_U][NV];
          float w[SS_U][RC];
#pragma unroll
          for (int u = 0; u < SS_U; ++u) {
            const int src = (u0 + u) & 31;
            const int64_t k = __shfl_sync(0xffffffffu, my_k, src);
#pragma unroll
            for (int i = 0; i < RC; ++i) w[u][i] = __shfl_sync(0xffffffffu, my_w[i], src);
#pragma unroll
            for (int c = 0; c < NV; ++c) {
              const int v = lane + c * 32;
              g[u][c] = (u0 + u < cnt && v < nvec) ? __ldg(dP + ((int64_t)half * T + k) * nvec + v) : f4_zero();
            }
          }
#pragma unroll
          for (int u = 0; u < SS_U; ++u) {
            if (u0 + u >= cnt) break;
#pragma unroll
            for (int c = 0; c < NV; ++c)
#pragma unroll
              for (int i = 0; i < RC; ++i) {
                kahan_add(acc[i][c].x, cmp[i][c].x, w[u][i] * g[u][c].x);
                kahan_add(acc[i][c].y, cmp[i][c].y, w[u][i] * g[u][c].y);
                kahan_add(acc[i][c].z, cmp[i][c].z, w[u][i] * g[u][c].z);
                kahan_add(acc[i][c].w, cmp[i][c].w, w[u][i] * g[u][c].w);
              }
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < RC; ++i) {
      if (r0 + i >= planes) continue;
      float4* orow = out + ((int64_t)(r0 + i) * N + n) * nvec;
#pragma unroll
      for (int c = 0; c < NV; ++c) {
        const int v = lane + c * 32;
        if (v < nvec) {
          float4 r = acc[i][c];
          if (accumulate) {
            const float4 pv = orow[v];
            r.x += pv.x; r.y += pv.y; r.z += pv.z; r.w += pv.w;
          }
          orow[v] = r;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// The same segment sums in the flat-stream pattern of spmm.cu's stream_rows_kernel, for plans large enough to fill the
// machine (the warp-per-node kernel above pays three dependent latencies — pointer, permutation, row — for the ~5 rows
// an average node owns).  A slice of LPR lanes owns a chunk of LPR consecutive nodes; the chunk's run of the plan's
// merged node-major list (by_node_src: rows of the (2,T,D) buffer viewed as [2T, D]) is staged in shared memory
// together with the RC gate values of every entry (sigmoid applied once), and consumed in batches of U independent
// 128-bit row gathers regardless of node boundaries.  Kahan-compensated like the kernel above; the list order is the
// same (head-side triples, then tail-side), so the two kernels agree to rounding of the compensation terms.
//   out[(r, n), :] (+)= Σ_{j in ptr[n]..ptr[n+1]} gate_r(src[j] mod T) · dP[src[j], :]     r in [r0, r0+RC)
constexpr int SEG_SLAB = 256;   // staged entries per warp

template <int LPR, int U, int RC, bool GATED>
__global__ void __launch_bounds__(TW * 32, 4)
    segment_stream_kernel(const int32_t* __restrict__ ptr, const int32_t* __restrict__ src,
                          const float4* __restrict__ dP, const float* __restrict__ alpha, int64_t T, int64_t N, int nvec,
                          int R, float4* __restrict__ out, int accumulate) {
  constexpr int SUBS = 32 / LPR;
  constexpr int SUB_SLAB = SEG_SLAB / SUBS;
  __shared__ int32_t s_idx[TW][SEG_SLAB];
  __shared__ float s_w[GATED ? RC : 1][TW][GATED ? SEG_SLAB : 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / LPR, l = lane % LPR;
  const int64_t chunk = ((int64_t)blockIdx.x * TW + warp) * SUBS + sub;
  const int64_t n0 = chunk * LPR;
  if (n0 >= N) return;
  const unsigned mask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (sub * LPR));
  const int nrows = (N - n0 < LPR) ? (int)(N - n0) : LPR;
  const int r0 = blockIdx.y * RC;
  const int planes = GATED ? R : 1;
  int32_t* const sl_idx = &s_idx[warp][sub * SUB_SLAB];

  const int chunk_beg = ld_stream_i32(ptr + n0);
  const int my_end = (l < nrows) ? ld_stream_i32(ptr + n0 + l + 1) : 0x7fffffff;
  const int chunk_end = __shfl_sync(mask, my_end, nrows - 1, LPR);
  const bool live = l < nvec;

  int row = 0;
  int cur_end = __shfl_sync(mask, my_end, 0, LPR);
  int pos = chunk_beg, slab_lo = chunk_beg, slab_hi = chunk_beg;
  float4 acc[RC], cmp[RC];
#pragma unroll
  for (int i = 0; i < RC; ++i) { acc[i] = f4_zero(); cmp[i] = f4_zero(); }

  auto flush = [&]() {
    if (live) {
#pragma unroll
      for (int i = 0; i < RC; ++i) {
        if (r0 + i < planes) {
          float4* o = out + ((int64_t)(r0 + i) * N + n0 + row) * nvec + l;
          float4 r = acc[i];
          if (accumulate) {
            const float4 pv = *o;
            r.x += pv.x; r.y += pv.y; r.z += pv.z; r.w += pv.w;
          }
          *o = r;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < RC; ++i) { acc[i] = f4_zero(); cmp[i] = f4_zero(); }
    ++row;
    cur_end = __shfl_sync(mask, my_end, row & (LPR - 1), LPR);
    if (row >= nrows) cur_end = 0x7fffffff;
  };

  while (pos < chunk_end) {
    if (pos + U > slab_hi && slab_hi < chunk_end) {   // refill the slab starting at the current position
      __syncwarp(mask);
      slab_lo = pos;
      slab_hi = (chunk_end - pos < SUB_SLAB) ? chunk_end : pos + SUB_SLAB;
      for (int q = l; q < slab_hi - slab_lo; q += LPR) {
        const int s_ = ld_stream_i32(src + slab_lo + q);
        sl_idx[q] = s_;
        if (GATED) {
          const int64_t k = s_ >= T ? s_ - T : s_;
#pragma unroll
          for (int i = 0; i < RC; ++i)
            s_w[i][warp][sub * SUB_SLAB + q] = (r0 + i < R) ? sigmoidf_(__ldg(alpha + k * R + r0 + i)) : 0.f;
        }
      }
      __syncwarp(mask);
    }
    const int n = (slab_hi - pos < U) ? (slab_hi - pos) : U;
    float4 v[U];
    float w[U][RC];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (u < n) {
        const int q = pos + u - slab_lo;
        const int sidx = sl_idx[q];
        if (GATED) {
#pragma unroll
          for (int i = 0; i < RC; ++i) w[u][i] = s_w[i][warp][sub * SUB_SLAB + q];
        }
        v[u] = live ? __ldg(dP + (int64_t)sidx * nvec + l) : f4_zero();
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (u < n) {
        while (cur_end == pos + u) flush();
#pragma unroll
        for (int i = 0; i < RC; ++i) {
          const float ww = GATED ? w[u][i] : 1.0f;
          kahan_add(acc[i].x, cmp[i].x, ww * v[u].x);
          kahan_add(acc[i].y, cmp[i].y, ww * v[u].y);
          kahan_add(acc[i].z, cmp[i].z, ww * v[u].z);
          kahan_add(acc[i].w, cmp[i].w, ww * v[u].w);
        }
      }
    }
    pos += n;
  }
  while (row < nrows) flush();
}

constexpr int64_t SEG_STREAM_MIN_NODES = 65536;
static bool use_segment_stream(const iddgcn_plan* p, int nvec) {
  const char* e = getenv("IDDGCN_SEG_STREAM");   // "0" / "1" force the choice (A/B runs and tests)
  if (e && (e[0] == '0' || e[0] == '1')) return e[0] == '1' && nvec <= 32;
  return p->N >= SEG_STREAM_MIN_NODES && nvec <= 32;
}

template <int LPR>
static int launch_segment_stream(const iddgcn_plan* p, const float* dP, const float* alpha, int nvec, float* out,
                                 int accumulate, cudaStream_t st) {
  const int64_t blocks = (p->N + TW * 32 - 1) / (TW * 32);   // every warp covers 32 nodes (SUBS chunks of LPR)
  IDD_CHECK(blocks < (int64_t)0x7fffffff, IDDGCN_ELIMIT, "too many nodes for one launch");
  const float4* d4 = reinterpret_cast<const float4*>(dP);
  float4* o4 = reinterpret_cast<float4*>(out);
  if (alpha) {
    if (p->R >= 2) {
      const dim3 grid((unsigned)blocks, (unsigned)((p->R + 1) / 2));
      segment_stream_kernel<LPR, 4, 2, true><<<grid, TW * 32, 0, st>>>(p->by_node_ptr, p->by_node_src, d4, alpha, p->T,
                                                                       p->N, nvec, p->R, o4, accumulate);
    } else {
      segment_stream_kernel<LPR, 8, 1, true><<<(unsigned)blocks, TW * 32, 0, st>>>(p->by_node_ptr, p->by_node_src, d4,
                                                                                   alpha, p->T, p->N, nvec, p->R, o4,
                                                                                   accumulate);
    }
  } else {
    segment_stream_kernel<LPR, 8, 1, false><<<(unsigned)blocks, TW * 32, 0, st>>>(p->by_node_ptr, p->by_node_src, d4,
                                                                                  nullptr, p->T, p->N, nvec, p->R, o4,
                                                                                  accumulate);
  }
  IDD_LAUNCH_CHECK();
  return IDDGCN_OK;
}

int segment_scatter(const iddgcn_plan* p, const float* dP, const float* alpha, int64_t D, float* out, int accumulate,
                    cudaStream_t st) {
  IDD_CHECK(D % 4 == 0 && D <= 512, IDDGCN_ELIMIT, "unsupported D=%lld", (long long)D);
  const int nvec = (int)(D / 4);
  if (p->T > 0 && use_segment_stream(p, nvec)) {
    if (nvec <= 4) return launch_segment_stream<4>(p, dP, alpha, nvec, out, accumulate, st);
    if (nvec <= 8) return launch_segment_stream<8>(p, dP, alpha, nvec, out, accumulate, st);
    if (nvec <= 16) return launch_segment_stream<16>(p, dP, alpha, nvec, out, accumulate, st);
    return launch_segment_stream<32>(p, dP, alpha, nvec, out, accumulate, st);
  }
  const unsigned gx = (unsigned)((p->N + TW - 1) / TW);
  // few nodes (the shipped graph has 845): one warp per (node, relation) instead of per node
  const bool split_rel = alpha != nullptr && p->N < 32768 && p->R > 1;
  const dim3 grid(gx, split_rel ? (unsigned)p->R : 1u);
#define IDD_ARGS p->by_head_ptr, p->by_head_perm, p->by_tail_ptr, p->by_tail_perm, reinterpret_cast<const float4*>(dP), \
                 alpha, p->T, p->N, nvec, p->R, reinterpret_cast<float4*>(out), accumulate
  if (alpha) {
    if (split_rel) {
      if (nvec <= 32) segment_scatter_kernel<1, true, 1><<<grid, TW * 32, 0, st>>>(IDD_ARGS);
      else if (nvec <= 64) segment_scatter_kernel<2, true, 1><<<grid, TW * 32, 0, st>>>(IDD_ARGS);
      else segment_scatter_kernel<4, true, 1><<<grid, TW * 32, 0, st>>>(IDD_ARGS);
    } else {
      // relation planes summed per sweep: no more than there are (R = 2 would otherwise carry two idle planes of
      // accumulators and compensation terms, and the registers cost a resident block)
      if (p->R <= 2) {
        if (nvec <= 32) segment_scatter_kernel<1, true, 2><<<grid, TW * 32, 0, st>>>(IDD_ARGS);
        else if (nvec <= 64) segment_scatter_kernel<2, true, 2><<<grid, TW * 32, 0, st>>>(IDD_ARGS);
        else segment_scatter_kernel<4, true, 2><<<grid, TW * 32, 0, st>>>(IDD_ARGS);
      } else {
        if (nvec <= 32) segment_scatter_kernel<1, true, 4><<<grid, TW * 32, 0, st>>>(IDD_ARGS);
        else if (nvec <= 64) segment_scatter_kernel<2, true, 4><<<grid, TW * 32, 0, st>>>(IDD_ARGS);
        else segment_scatter_kernel<4, true, 4><<<grid, TW * 32, 0, st>>>(IDD_ARGS);
      }
    }
  } else {
    if (nvec <= 32) segment_scatter_kernel<1, false, 1><<<grid, TW * 32, 0, st>>>(IDD_ARGS);
    else if (nvec <= 64) segment_scatter_kernel<2, false, 1><<<grid, TW * 32, 0, st>>>(IDD_ARGS);
    else segment_scatter_kernel<4, false, 1><<<grid, TW * 32, 0, st>>>(IDD_ARGS);
  }
#undef IDD_ARGS
  IDD_LAUNCH_CHECK();
  return IDDGCN_OK;
}

// ---------------------------------------------------------------------------------------------
// column sums (db_alpha), two fixed-order stages
constexpr int CS_BLOCK = 256;
constexpr int CS_ROWS_PER_BLOCK = 4096;

__global__ void __launch_bounds__(CS_BLOCK) colsum_partial_kernel(const float* __restrict__ in, int64_t rows, int cols,
                                                                  float* __restrict__ partial) {
  __shared__ float s[CS_BLOCK];
  const int64_t r0 = (int64_t)blockIdx.x * CS_ROWS_PER_BLOCK;
  const int64_t r1 = (r0 + CS_ROWS_PER_BLOCK < rows) ? r0 + CS_ROWS_PER_BLOCK : rows;
  for (int c = 0; c < cols; ++c) {
    float acc = 0.f;
    for (int64_t r = r0 + threadIdx.x; r < r1; r += CS_BLOCK) acc += in[r * cols + c];
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = CS_BLOCK / 2; o > 0; o >>= 1) {
      if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) partial[(int64_t)blockIdx.x * cols + c] = s[0];
    __syncthreads();
  }
}

__global__ void colsum_final_kernel(const float* __restrict__ partial, int64_t nblocks, int cols,
                                    float* __restrict__ out, int accumulate) {
  const int c = threadIdx.x;
  if (c >= cols) return;
  float acc = 0.f;
  for (int64_t b = 0; b < nblocks; ++b) acc += partial[b * cols + c];
  out[c] = accumulate ? out[c] + acc : acc;
}

size_t colsum_scratch_elems(int64_t rows, int cols) {
  return (size_t)((rows + CS_ROWS_PER_BLOCK - 1) / CS_ROWS_PER_BLOCK + 1) * cols;
}

int colsum(const float* in, int64_t rows, int cols, float* out, int accumulate, float* scratch, cudaStream_t st) {
  IDD_CHECK(cols <= 32, IDDGCN_ELIMIT, "colsum supports at most 32 columns");
  const int64_t nb = (rows + CS_ROWS_PER_BLOCK - 1) / CS_ROWS_PER_BLOCK;
  if (nb > 0) {
    colsum_partial_kernel<<<(unsigned)nb, CS_BLOCK, 0, st>>>(in, rows, cols, scratch);
    IDD_LAUNCH_CHECK();
  }
  colsum_final_kernel<<<1, 32, 0, st>>>(scratch, nb, cols, out, accumulate);
  IDD_LAUNCH_CHECK();
  return IDDGCN_OK;
}

// ---------------------------------------------------------------------------------------------
// DistMult decoder
template <int NV>
__global__ void __launch_bounds__(TW * 32)
    distmult_fwd_kernel(const int32_t* __restrict__ rel, const float4* __restrict__ HT,
                        const float4* __restrict__ rel_emb, int64_t T, int nvec, float* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int64_t k = (int64_t)blockIdx.x * TW + (threadIdx.x >> 5);
  if (k >= T) return;
  const int r = rel[k];
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < NV; ++c) {
    const int v = lane + c * 32;
    if (v < nvec) {
      const float4 h = __ldg(HT + k * nvec + v), t = __ldg(HT + (T + k) * nvec + v);
      const float4 e = __ldg(rel_emb + (int64_t)r * nvec + v);
      acc += h.x * e.x * t.x + h.y * e.y * t.y + h.z * e.z * t.z + h.w * e.w * t.w;  // IDDGCN.py:108
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) y[k] = sigmoidf_(acc);
}

int distmult_fwd(const iddgcn_plan* p, const float* HT, const float* rel_emb, int64_t D, float* y, cudaStream_t st) {
  if (p->T == 0) return IDDGCN_OK;
  IDD_CHECK(D % 4 == 0 && D <= 512, IDDGCN_ELIMIT, "unsupported D=%lld", (long long)D);
  const int nvec = (int)(D / 4);
  const unsigned grid = (unsigned)((p->T + TW - 1) / TW);
#define IDD_ARGS p->rel, reinterpret_cast<const float4*>(HT), reinterpret_cast<const float4*>(rel_emb), p->T, nvec, y
  if (nvec <= 32) distmult_fwd_kernel<1><<<grid, TW * 32, 0, st>>>(IDD_ARGS);
  else if (nvec <= 64) distmult_fwd_kernel<2><<<grid, TW * 32, 0, st>>>(IDD_ARGS);
  else distmult_fwd_kernel<4><<<grid, TW * 32, 0, st>>>(IDD_ARGS);
#undef IDD_ARGS
  IDD_LAUNCH_CHECK();
  return IDDGCN_OK;
}

template <int NV>
__global__ void __launch_bounds__(TW * 32)
    distmult_bwd_kernel(const int32_t* __restrict__ rel, const float4* __restrict__ HT,
                        const float4* __restrict__ rel_emb, const float* __restrict__ y,
                        const float* __restrict__ dy, int64_t T, int nvec, float4* __restrict__ dHT) {
  const int lane = threadIdx.x & 31;
  const int64_t k = (int64_t)blockIdx.x * TW + (threadIdx.x >> 5);
  if (k >= T) return;
  const int r = rel[k];
  const float yk = y[k];
  const float dq = dy[k] * yk * (1.f - yk);  // eq. 14
#pragma unroll
  for (int c = 0; c < NV; ++c) {
    const int v = lane + c * 32;
    if (v < nvec) {
      const float4 h = __ldg(HT + k * nvec + v), t = __ldg(HT + (T + k) * nvec + v);
      const float4 e = __ldg(rel_emb + (int64_t)r * nvec + v);
      dHT[k * nvec + v] = make_float4(dq * e.x * t.x, dq * e.y * t.y, dq * e.z * t.z, dq * e.w * t.w);
      dHT[(T + k) * nvec + v] = make_float4(dq * e.x * h.x, dq * e.y * h.y, dq * e.z * h.z, dq * e.w * h.w);
    }
  }
}

// d_rel_emb partials: block (chunk c, relation r) sums its slice of the by-rel run, thread per feature
constexpr int DR_CHUNK = 128;
__global__ void __launch_bounds__(128)
    distmult_drel_kernel(const int32_t* __restrict__ ptr_r, const int32_t* __restrict__ perm_r,
                         const float* __restrict__ HT, const float* __restrict__ y, const float* __restrict__ dy,
                         int64_t T, int D, int nchunk, float* __restrict__ partial /* [R][nchunk][D] */) {
  const int r = blockIdx.y, c = blockIdx.x;
  const int beg = ptr_r[r], end = ptr_r[r + 1];
  const int len = end - beg;
  const int per = (len + nchunk - 1) / nchunk;
  const int j0 = beg + c * per;
  const int j1 = (j0 + per < end) ? j0 + per : end;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float acc = 0.f;
    for (int j = j0; j < j1; ++j) {
      const int64_t k = perm_r[j];
      const float yk = y[k];
      const float dq = dy[k] * yk * (1.f - yk);
      acc = fmaf(dq, HT[k * D + d] * HT[(T + k) * D + d], acc);
    }
    partial[((int64_t)r * nchunk + c) * D + d] = acc;
  }
}

__global__ void drel_final_kernel(const float* __restrict__ partial, int nchunk, int D, float* __restrict__ out) {
  const int r = blockIdx.y;
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  float acc = 0.f;
  for (int c = 0; c < nchunk; ++c) acc += partial[((int64_t)r * nchunk + c) * D + d];
  out[(int64_t)r * D + d] += acc;
}

static int drel_chunks(int64_t T) {
  int64_t n = (T + DR_CHUNK - 1) / DR_CHUNK;
  if (n < 1) n = 1;
  if (n > 4096) n = 4096;
  return (int)n;
}

size_t distmult_scratch_elems(const iddgcn_plan* p, int64_t D) { return (size_t)p->R * drel_chunks(p->T) * D; }

int distmult_bwd(const iddgcn_plan* p, const float* HT, const float* rel_emb, const float* y, const float* dy,
                 int64_t D, float* dHT, float* d_rel_emb, float* scratch, cudaStream_t st) {
  if (p->T == 0) return IDDGCN_OK;
  IDD_CHECK(D % 4 == 0 && D <= 512, IDDGCN_ELIMIT, "unsupported D=%lld", (long long)D);
  const int nvec = (int)(D / 4);
  const unsigned grid = (unsigned)((p->T + TW - 1) / TW);
#define IDD_ARGS p->rel, reinterpret_cast<const float4*>(HT), reinterpret_cast<const float4*>(rel_emb), y, dy, p->T, \
                 nvec, reinterpret_cast<float4*>(dHT)
  if (nvec <= 32) distmult_bwd_kernel<1><<<grid, TW * 32, 0, st>>>(IDD_ARGS);
  else if (nvec <= 64) distmult_bwd_kernel<2><<<grid, TW * 32, 0, st>>>(IDD_ARGS);
  else distmult_bwd_kernel<4><<<grid, TW * 32, 0, st>>>(IDD_ARGS);
#undef IDD_ARGS
  IDD_LAUNCH_CHECK();
  if (d_rel_emb) {
    const int nchunk = drel_chunks(p->T);
    distmult_drel_kernel<<<dim3(nchunk, p->R), 128, 0, st>>>(p->by_rel_ptr, p->by_rel_perm, HT, y, dy, p->T, (int)D,
                                                             nchunk, scratch);
    IDD_LAUNCH_CHECK();
    drel_final_kernel<<<dim3((unsigned)((D + 127) / 128), p->R), 128, 0, st>>>(scratch, nchunk, (int)D, d_rel_emb);
    IDD_LAUNCH_CHECK();
  }
  return IDDGCN_OK;
}

// ---------------------------------------------------------------------------------------------
// Keras BinaryCrossentropy (from_logits=False): clip to [eps, 1-eps], log(p + eps)  
constexpr int BCE_BLOCK = 256;
constexpr int BCE_PER_BLOCK = 4096;
__global__ void __launch_bounds__(BCE_BLOCK)
    bce_kernel(const float* __restrict__ y, int64_t T_pos, int64_t T_total, float grad_scale, float* __restrict__ dy,
               float* __restrict__ partial) {
  __shared__ float s[BCE_BLOCK];
  const float eps = 1e-7f;
  const int64_t i0 = (int64_t)blockIdx.x * BCE_PER_BLOCK;
  const int64_t i1 = (i0 + BCE_PER_BLOCK < T_total) ? i0 + BCE_PER_BLOCK : T_total;
  const float inv_n = 1.0f / (float)T_total;
  float acc = 0.f;
  for (int64_t i = i0 + threadIdx.x; i < i1; i += BCE_BLOCK) {
    const float p = y[i];
    const float pc = fminf(fmaxf(p, eps), 1.0f - eps);
    const bool inside = (p >= eps) && (p <= 1.0f - eps);
    float term, g;
    if (i < T_pos) {  // label 1
      term = -logf(pc + eps);
      g = -1.0f / (pc + eps);
    } else {          // label 0
      term = -logf(1.0f - pc + eps);
      g = 1.0f / (1.0f - pc + eps);
    }
    acc += term;
    dy[i] = inside ? g * inv_n * grad_scale : 0.f;  // eq. 15
  }
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = BCE_BLOCK / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = s[0];
}

__global__ void bce_final_kernel(const float* __restrict__ partial, int64_t nb, int64_t T_total,
                                 float* __restrict__ loss_out) {
  if (threadIdx.x != 0) return;
  float acc = 0.f;
  for (int64_t b = 0; b < nb; ++b) acc += partial[b];
  loss_out[0] = acc / (float)T_total;
}

size_t bce_scratch_elems(int64_t T_total) { return (size_t)((T_total + BCE_PER_BLOCK - 1) / BCE_PER_BLOCK + 1); }

int bce_loss_grad(const float* y, int64_t T_pos, int64_t T_total, float grad_scale, float* dy, float* loss_out,
                  float* scratch, cudaStream_t st) {
  IDD_CHECK(T_total > 0, IDDGCN_EINVAL, "BCE over zero predictions");
  const int64_t nb = (T_total + BCE_PER_BLOCK - 1) / BCE_PER_BLOCK;
  bce_kernel<<<(unsigned)nb, BCE_BLOCK, 0, st>>>(y, T_pos, T_total, grad_scale, dy, scratch);
  IDD_LAUNCH_CHECK();
  bce_final_kernel<<<1, 32, 0, st>>>(scratch, nb, T_total, loss_out);
  IDD_LAUNCH_CHECK();
  return IDDGCN_OK;
}

// ---------------------------------------------------------------------------------------------
// Keras Adam (OptimizerV2): lr_t = lr·sqrt(1-b2^t)/(1-b1^t); theta -= lr_t·m/(sqrt(v)+eps)   
__global__ void __launch_bounds__(256)
    adam_kernel(float* __restrict__ theta, const float* __restrict__ grad, float* __restrict__ m,
                float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps, int64_t step,
                const int64_t* __restrict__ step_dev) {
  const int64_t t = step_dev ? *step_dev : step;
  // double keeps b^t accurate for t in the thousands; the two pow() are paid once per thread, the elements are then
  // visited with a grid stride (a million-row embedding table would otherwise spend its time in double-precision pow)
  const double c1 = 1.0 - pow((double)b1, (double)t), c2 = 1.0 - pow((double)b2, (double)t);
  const float lr_t = (float)((double)lr * sqrt(c2) / c1);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float g = grad[i];
    const float mi = b1 * m[i] + (1.0f - b1) * g;
    const float vi = b2 * v[i] + (1.0f - b2) * g * g;
    m[i] = mi;
    v[i] = vi;
    theta[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

int adam_apply(float* theta, const float* grad, float* m, float* v, int64_t n, float lr, float b1, float b2,
               float eps, int64_t step, const int64_t* step_dev, cudaStream_t st) {
  if (n <= 0) return IDDGCN_OK;
  const int64_t blocks = (n + 255) / 256, cap = 16 * (int64_t)sm_count();
  adam_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(theta, grad, m, v, n, lr, b1, b2, eps, step,
                                                                     step_dev);
  IDD_LAUNCH_CHECK();
  return IDDGCN_OK;
}

__global__ void increment_kernel(int64_t* c) { *c += 1; }
int increment_counter(int64_t* counter, cudaStream_t st) {
  increment_kernel<<<1, 1, 0, st>>>(counter);
  IDD_LAUNCH_CHECK();
  return IDDGCN_OK;
}

}  // namespace iddgcn
