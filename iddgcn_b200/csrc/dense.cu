// dense.cu — fp32 SIMT GEMMs for the triple stage (see dense.cuh).
//
// Call sites they replace: tf.matmul(head_e, self_kernel) / tf.matmul(head_update, relation_kernel[i])
// (prediction/IDDGCN.py:62-63,76-77) and the matmul gradients TF autodiff emits for them (IDDGCN.py:172).
#include "dense.cuh"

namespace iddgcn {

constexpr int GK = 16;       // K step
constexpr int GPAD = 4;      // keeps rows 16-byte aligned and spreads banks
constexpr int GTHREADS = 256;

// 256 threads as a 16 x 16 grid, each owning a TM x TN register micro-tile: the C tile is (16·TM) x (16·TN).
//   <4,4>: 64 x 64   — narrow outputs (D = 64, W_alpha with N = R)
//   <8,8>: 128 x 128 — D >= 128: 64 FMAs per 4 shared-memory vector loads instead of 16 per 2
template <bool TA, bool TB, int TM, int TN>
__global__ void __launch_bounds__(GTHREADS, (TM * TN > 16 ? 2 : 1))
    gemm_tile_kernel(int64_t M, int64_t N, int64_t K, const float* __restrict__ A, int64_t lda, int64_t strideA,
                     const float* __restrict__ B, int64_t ldb, int64_t strideB, float* __restrict__ C, int64_t ldc,
                     int64_t strideC, int accumulate, int64_t k_chunk, int64_t ws_split_stride) {
  // grid: x = tile over M, y = tile over N, z = batch * splits (splits > 1 only for the split-K variant,
  // where C points at the workspace and ws_split_stride separates the partial results).
  constexpr int BM = 16 * TM, BN = 16 * TN;
  __shared__ __align__(16) float As[GK][BM + GPAD];
  __shared__ __align__(16) float Bs[GK][BN + GPAD];
  const int splits = (k_chunk > 0 && K > 0) ? (int)((K + k_chunk - 1) / k_chunk) : 1;
  const int b = blockIdx.z / splits;
  const int sp = blockIdx.z % splits;
  const int64_t k_begin = (k_chunk > 0) ? (int64_t)sp * k_chunk : 0;
  const int64_t k_end = (k_chunk > 0) ? ((k_begin + k_chunk < K) ? k_begin + k_chunk : K) : K;
  A += (int64_t)b * strideA;
  B += (int64_t)b * strideB;
  C += (int64_t)b * strideC + (int64_t)sp * ws_split_stride;
  const int64_t m0 = (int64_t)blockIdx.x * BM, n0 = (int64_t)blockIdx.y * BN;
  const int tid = threadIdx.x;
  const int ty = tid / 16, tx = tid % 16;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = k_begin; k0 < k_end; k0 += GK) {
#pragma unroll
    for (int i = 0; i < (BM * GK) / GTHREADS; ++i) {
      const int idx = tid + i * GTHREADS;
      int m, k;
      if (TA) { m = idx % BM; k = idx / BM; } else { k = idx % GK; m = idx / GK; }
      const int64_t gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < k_end) v = TA ? A[gk * lda + gm] : A[gm * lda + gk];
      As[k][m] = v;
    }
#pragma unroll
    for (int i = 0; i < (BN * GK) / GTHREADS; ++i) {
      const int idx = tid + i * GTHREADS;
      int n, k;
      if (TB) { k = idx % GK; n = idx / GK; } else { n = idx % BN; k = idx / BN; }
      const int64_t gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < N && gk < k_end) v = TB ? B[gn * ldb + gk] : B[gk * ldb + gn];
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      // micro-tile rows/cols come in groups of 4 spaced 64 apart (ty*4 + 64*g), so that the 16 threads of a row of the
      // thread grid read 16 consecutive float4 from shared memory (conflict-free) instead of strided ones
      float a[TM], bb[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4 + (i / 4) * 64]);
        a[i] = a4.x; a[i + 1] = a4.y; a[i + 2] = a4.z; a[i + 3] = a4.w;
      }
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4 + (j / 4) * 64]);
        bb[j] = b4.x; bb[j + 1] = b4.y; bb[j + 2] = b4.z; bb[j + 3] = b4.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t gm = m0 + ty * 4 + (i / 4) * 64 + (i % 4);
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int64_t gn = n0 + tx * 4 + (j / 4) * 64 + (j % 4);
      if (gn >= N) continue;
      float* c = C + gm * ldc + gn;
      *c = accumulate ? (*c + acc[i][j]) : acc[i][j];
    }
  }
}

// C[b][m][n] (+)= sum_{s < splits} ws[b][s][m][n].  256 threads = 8 split-lanes x 32 consecutive outputs: split-lane j
// sums partials j, j+8, ... in ascending order, then the 8 lane sums are added in ascending j (fixed order).
__global__ void __launch_bounds__(256)
    splitk_reduce_kernel(const float* __restrict__ ws, int splits, int64_t MN, int64_t N, float* __restrict__ C,
                         int64_t ldc, int64_t strideC, int accumulate) {
  __shared__ float s_part[8][32];
  const int o = threadIdx.x & 31, j = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 32 + o;
  const int b = blockIdx.y;
  float s = 0.f;
  if (i < MN) {
    const float* w = ws + (int64_t)b * splits * MN + i;
    for (int k = j; k < splits; k += 8) s += w[(int64_t)k * MN];
  }
  s_part[j][o] = s;
  __syncthreads();
  if (j == 0 && i < MN) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += s_part[q][o];
    float* c = C + (int64_t)b * strideC + (i / N) * ldc + (i % N);
    *c = accumulate ? (*c + t) : t;
  }
}

template <int TM, int TN>
static int gemm_launch(bool transA, bool transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda,
                       int64_t strideA, const float* B, int64_t ldb, int64_t strideB, float* C, int64_t ldc,
                       int64_t strideC, int batch, int acc, cudaStream_t st) {
  const int64_t gm = (M + 16 * TM - 1) / (16 * TM), gn = (N + 16 * TN - 1) / (16 * TN);
  IDD_CHECK(gm < 0x7fffffff && gn < 65536 && batch < 65536, IDDGCN_ELIMIT, "GEMM grid too large");
  dim3 grid((unsigned)gm, (unsigned)gn, (unsigned)batch);
  if (!transA && !transB)
    gemm_tile_kernel<false, false, TM, TN><<<grid, GTHREADS, 0, st>>>(M, N, K, A, lda, strideA, B, ldb, strideB, C, ldc, strideC, acc, 0, 0);
  else if (!transA && transB)
    gemm_tile_kernel<false, true, TM, TN><<<grid, GTHREADS, 0, st>>>(M, N, K, A, lda, strideA, B, ldb, strideB, C, ldc, strideC, acc, 0, 0);
  else if (transA && !transB)
    gemm_tile_kernel<true, false, TM, TN><<<grid, GTHREADS, 0, st>>>(M, N, K, A, lda, strideA, B, ldb, strideB, C, ldc, strideC, acc, 0, 0);
  else
    gemm_tile_kernel<true, true, TM, TN><<<grid, GTHREADS, 0, st>>>(M, N, K, A, lda, strideA, B, ldb, strideB, C, ldc, strideC, acc, 0, 0);
  IDD_LAUNCH_CHECK();
  return IDDGCN_OK;
}

int gemm_f32(bool transA, bool transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, int64_t strideA,
             const float* B, int64_t ldb, int64_t strideB, float* C, int64_t ldc, int64_t strideC, int batch,
             bool accumulate, cudaStream_t st) {
  if (M <= 0 || N <= 0 || batch <= 0) return IDDGCN_OK;
  // long-M, weight-sized products go to the tensor cores (3xTF32, tc_gemm.cu)
  if (tc_gemm_eligible(transA, M, N, K, A, lda, B, ldb, C, ldc) && (batch == 1 || ((strideA | strideB | strideC) & 3) == 0) &&
      tc_gemm_enabled())
    return gemm_tf32x3(transB, M, N, K, A, lda, strideA, B, strideB, C, ldc, strideC, batch, accumulate, st);
  return gemm_f32_simt(transA, transB, M, N, K, A, lda, strideA, B, ldb, strideB, C, ldc, strideC, batch, accumulate, st);
}

int gemm_f32_simt(bool transA, bool transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda,
                  int64_t strideA, const float* B, int64_t ldb, int64_t strideB, float* C, int64_t ldc, int64_t strideC,
                  int batch, bool accumulate, cudaStream_t st) {
  if (M <= 0 || N <= 0 || batch <= 0) return IDDGCN_OK;
  const int acc = accumulate ? 1 : 0;
  // big tiles only when they are filled and there are enough of them to cover the SMs
  if (N >= 128 && M * batch >= 128 * (int64_t)sm_count())
    return gemm_launch<8, 8>(transA, transB, M, N, K, A, lda, strideA, B, ldb, strideB, C, ldc, strideC, batch, acc, st);
  return gemm_launch<4, 4>(transA, transB, M, N, K, A, lda, strideA, B, ldb, strideB, C, ldc, strideC, batch, acc, st);
}

// ---- tall-skinny reduction: C[M,N] (+)= A[K,M]ᵀ · B[K,N] with N <= 8 (the gate gradient dW_alpha = Hᵀ·dz, N = R) -------
// A 64x64 register tile wastes 7/8 of its work on N = 2..8, and the product is a weighted column sum of A anyway: every
// warp streams whole rows of A (lane = one float4 of columns, coalesced), multiplies by the row's N weights (broadcast
// loads) and keeps N x 4 running sums in registers.  HBM-bound on A.  Warps of a block are combined in a fixed order and
// the blocks by splitk_reduce_kernel: deterministic.
constexpr int SK_WARPS = 8;
constexpr int SK_NMAX = 8;

template <int NV, int NMAX = SK_NMAX>   // NMAX: columns of B kept in registers (>= N); 2 for the gate gradient at R = 2
__global__ void __launch_bounds__(SK_WARPS * 32)
    tn_skinny_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb, int64_t K, int M,
                     int N, int64_t rows_per_block, float* __restrict__ ws) {
  extern __shared__ float sk_smem[];                         // [SK_WARPS][M][N]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = M >> 2;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < K ? r0 + rows_per_block : K;
  float acc[NV][NMAX][4];
#pragma unroll
  for (int c = 0; c < NV; ++c)
#pragma unroll
    for (int n = 0; n < NMAX; ++n) acc[c][n][0] = acc[c][n][1] = acc[c][n][2] = acc[c][n][3] = 0.f;
  // SK_UNROLL rows of this warp are requested before any is consumed (one 512-byte row per iteration left the warp
  // with a single load in flight); rows are still accumulated in the same order, so results are bit-identical
  constexpr int SK_UNROLL = NV == 1 ? (NMAX <= 2 ? 8 : 4) : 2;   // rows in flight per warp
  for (int64_t rb = r0 + warp; rb < r1; rb += (int64_t)SK_WARPS * SK_UNROLL) {
    float b[SK_UNROLL][NMAX];
    float4 a[SK_UNROLL][NV];
#pragma unroll
    for (int u = 0; u < SK_UNROLL; ++u) {
      const int64_t r = rb + (int64_t)u * SK_WARPS;
#pragma unroll
      for (int n = 0; n < NMAX; ++n) b[u][n] = (r < r1 && n < N) ? __ldg(B + r * ldb + n) : 0.f;
#pragma unroll
      for (int c = 0; c < NV; ++c) {
        const int v = lane + c * 32;
        a[u][c] = (r < r1 && v < nvec) ? ld_stream_f4(reinterpret_cast<const float4*>(A + r * lda) + v)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < SK_UNROLL; ++u) {
      if (rb + (int64_t)u * SK_WARPS >= r1) break;
#pragma unroll
      for (int c = 0; c < NV; ++c) {
        if (lane + c * 32 < nvec) {
#pragma unroll
          for (int n = 0; n < NMAX; ++n) {
            acc[c][n][0] += a[u][c].x * b[u][n]; acc[c][n][1] += a[u][c].y * b[u][n];
            acc[c][n][2] += a[u][c].z * b[u][n]; acc[c][n][3] += a[u][c].w * b[u][n];
          }
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < NV; ++c) {
    const int v = lane + c * 32;
    if (v < nvec)
#pragma unroll
      for (int n = 0; n < NMAX; ++n)
        if (n < N)
#pragma unroll
          for (int e = 0; e < 4; ++e) sk_smem[((size_t)warp * M + v * 4 + e) * N + n] = acc[c][n][e];
  }
  __syncthreads();
  const int MN = M * N;
  for (int i = threadIdx.x; i < MN; i += SK_WARPS * 32) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < SK_WARPS; ++w) t += sk_smem[(size_t)w * MN + i];
    ws[(size_t)blockIdx.x * MN + i] = t;
  }
}

static int skinny_blocks(int64_t K) {
  const int64_t want = 2 * (int64_t)sm_count(), by_k = (K + 255) / 256;
  return (int)(want < by_k ? want : (by_k < 1 ? 1 : by_k));
}
static bool skinny_eligible(int64_t M, int64_t N, int64_t K, const float* A, int64_t lda) {
  return N <= SK_NMAX && M <= 512 && (M % 4) == 0 && (lda % 4) == 0 && K >= 1024 &&
         (size_t)SK_WARPS * M * N * sizeof(float) <= 48 * 1024 && (reinterpret_cast<uintptr_t>(A) & 15u) == 0;
}

// split-K tile edge: 128 when both output dimensions fill it (A and B are then read exactly once), else 64
static inline int splitk_tile(int64_t M, int64_t N) { return (M >= 128 && N >= 128) ? 128 : 64; }

static void splitk_shape(int64_t M, int64_t N, int64_t K, int batch, int* splits, int64_t* k_chunk) {
  const int GT = splitk_tile(M, N);
  const int64_t tiles = ((M + GT - 1) / GT) * ((N + GT - 1) / GT) * batch;
  int64_t want = (2 * (int64_t)sm_count() + tiles - 1) / tiles;   // ~2 waves of blocks
  int64_t max_by_k = (K + 4 * GK - 1) / (4 * GK);                  // at least 64 k per split
  if (want > max_by_k) want = max_by_k;
  if (want < 1) want = 1;
  if (want > 1024) want = 1024;
  int64_t chunk = (K + want - 1) / want;
  chunk = (chunk + GK - 1) / GK * GK;
  if (chunk < GK) chunk = GK;
  *k_chunk = chunk;
  *splits = (int)((K + chunk - 1) / chunk);
  if (*splits < 1) *splits = 1;
}

size_t gemm_tn_workspace_elems(int64_t M, int64_t N, int64_t K, int batch) {
  int splits; int64_t chunk;
  splitk_shape(M, N, K, batch, &splits, &chunk);
  if (M == N && (M == 64 || M == 128) && tc_gemm_tn_splits(K) > splits) splits = tc_gemm_tn_splits(K);   // either path fits
  if (N <= SK_NMAX && skinny_blocks(K) > splits) splits = skinny_blocks(K);
  return (size_t)splits * (size_t)M * (size_t)N * (size_t)batch;
}

int gemm_tn_splitk_f32(int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, int64_t strideA, const float* B,
                       int64_t ldb, int64_t strideB, float* C, int64_t ldc, int64_t strideC, int batch,
                       bool accumulate, float* workspace, cudaStream_t st) {
  if (M <= 0 || N <= 0 || batch <= 0) return IDDGCN_OK;
  const int64_t MN = M * N;
  // long reductions into a weight-sized square go to the tensor cores (3xTF32, tc_gemm.cu), same workspace + reduction
  if (tc_gemm_tn_eligible(M, N, K, A, lda, B, ldb) && (batch == 1 || ((strideA | strideB) & 3) == 0) && tc_gemm_enabled())
    return gemm_tn_tc(M, N, K, A, lda, strideA, B, ldb, strideB, C, ldc, strideC, batch, accumulate, workspace, st);
  if (skinny_eligible(M, N, K, A, lda) && (batch == 1 || (strideA & 3) == 0)) {
    const int blocks = skinny_blocks(K);
    const int64_t per = (K + blocks - 1) / blocks;
    const size_t smem = (size_t)SK_WARPS * MN * sizeof(float);
    for (int b = 0; b < batch; ++b) {
      const float* Ab = A + (int64_t)b * strideA;
      const float* Bb = B + (int64_t)b * strideB;
      float* wsb = workspace + (int64_t)b * blocks * MN;
      if (M <= 128 && N <= 2) tn_skinny_kernel<1, 2><<<blocks, SK_WARPS * 32, smem, st>>>(Ab, lda, Bb, ldb, K, (int)M, (int)N, per, wsb);
      else if (M <= 128) tn_skinny_kernel<1><<<blocks, SK_WARPS * 32, smem, st>>>(Ab, lda, Bb, ldb, K, (int)M, (int)N, per, wsb);
      else if (M <= 256) tn_skinny_kernel<2><<<blocks, SK_WARPS * 32, smem, st>>>(Ab, lda, Bb, ldb, K, (int)M, (int)N, per, wsb);
      else tn_skinny_kernel<4><<<blocks, SK_WARPS * 32, smem, st>>>(Ab, lda, Bb, ldb, K, (int)M, (int)N, per, wsb);
      IDD_LAUNCH_CHECK();
    }
    dim3 rgrid((unsigned)((MN + 31) / 32), (unsigned)batch);
    splitk_reduce_kernel<<<rgrid, 256, 0, st>>>(workspace, blocks, MN, N, C, ldc, strideC, accumulate ? 1 : 0);
    IDD_LAUNCH_CHECK();
    return IDDGCN_OK;
  }
  return gemm_tn_splitk_simt(M, N, K, A, lda, strideA, B, ldb, strideB, C, ldc, strideC, batch, accumulate, workspace, st);
}

int gemm_tn_tc(int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, int64_t strideA, const float* B,
               int64_t ldb, int64_t strideB, float* C, int64_t ldc, int64_t strideC, int batch, bool accumulate,
               float* workspace, cudaStream_t st) {
  const int64_t MN = M * N;
  const int splits = tc_gemm_tn_splits(K);
  for (int b = 0; b < batch; ++b)
    IDD_TRY(gemm_tn_tf32x3_partials(M, K, A + (int64_t)b * strideA, lda, B + (int64_t)b * strideB, ldb,
                                    workspace + (int64_t)b * splits * MN, st));
  dim3 rgrid((unsigned)((MN + 31) / 32), (unsigned)batch);
  splitk_reduce_kernel<<<rgrid, 256, 0, st>>>(workspace, splits, MN, N, C, ldc, strideC, accumulate ? 1 : 0);
  IDD_LAUNCH_CHECK();
  return IDDGCN_OK;
}

int gemm_tn_splitk_simt(int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, int64_t strideA, const float* B,
                        int64_t ldb, int64_t strideB, float* C, int64_t ldc, int64_t strideC, int batch,
                        bool accumulate, float* workspace, cudaStream_t st) {
  if (M <= 0 || N <= 0 || batch <= 0) return IDDGCN_OK;
  int splits; int64_t chunk;
  splitk_shape(M, N, K, batch, &splits, &chunk);
  if (K <= 0) splits = 1;
  const int64_t MN = M * N;
  const int GT = splitk_tile(M, N);
  dim3 grid((unsigned)((M + GT - 1) / GT), (unsigned)((N + GT - 1) / GT), (unsigned)(batch * splits));
  IDD_CHECK(grid.y < 65536 && grid.z < 65536, IDDGCN_ELIMIT, "split-K GEMM grid too large");
  // partials: ws[b][s][M][N]  (strideC = splits*MN per batch, ws_split_stride = MN)
  if (GT == 128)
    gemm_tile_kernel<true, false, 8, 8><<<grid, GTHREADS, 0, st>>>(M, N, K, A, lda, strideA, B, ldb, strideB, workspace,
                                                                   N, (int64_t)splits * MN, 0, chunk, MN);
  else
    gemm_tile_kernel<true, false, 4, 4><<<grid, GTHREADS, 0, st>>>(M, N, K, A, lda, strideA, B, ldb, strideB, workspace,
                                                                   N, (int64_t)splits * MN, 0, chunk, MN);
  IDD_LAUNCH_CHECK();
  dim3 rgrid((unsigned)((MN + 31) / 32), (unsigned)batch);
  splitk_reduce_kernel<<<rgrid, 256, 0, st>>>(workspace, splits, MN, N, C, ldc, strideC, accumulate ? 1 : 0);
  IDD_LAUNCH_CHECK();
  return IDDGCN_OK;
}

// out[b][j][i] = in[b][i][j] for small square-ish matrices (weights): lets big-M products with a transposed weight
// run through the faster non-transposed load path
__global__ void transpose_small_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const int r = i / cols, c = i % cols;
  out[(size_t)b * rows * cols + (size_t)c * rows + r] = in[(size_t)b * rows * cols + i];
}

int transpose_small(const float* in, float* out, int rows, int cols, int batch, cudaStream_t st) {
  if (rows <= 0 || cols <= 0 || batch <= 0) return IDDGCN_OK;
  transpose_small_kernel<<<dim3((unsigned)((rows * cols + 255) / 256), (unsigned)batch), 256, 0, st>>>(in, out, rows,
                                                                                                      cols);
  IDD_LAUNCH_CHECK();
  return IDDGCN_OK;
}

}  // namespace iddgcn
