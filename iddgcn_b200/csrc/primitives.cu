// primitives.cu — exclusive scan + stable LSD radix sort (see primitives.cuh).
#include "primitives.cuh"

namespace iddgcn {

// ---------------------------------------------------------------------------
// exclusive scan (reduce -> recurse on block sums -> apply)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t block_exclusive_scan_u32(uint32_t thread_sum, uint32_t* block_total) {
  __shared__ uint32_t s_warp[SCAN_THREADS / 32];
  __shared__ uint32_t s_total;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t incl = thread_sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    constexpr int NW = SCAN_THREADS / 32;
    uint32_t w = lane < NW ? s_warp[lane] : 0u;
    uint32_t wi = w;
#pragma unroll
    for (int o = 1; o < NW; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    if (lane < NW) s_warp[lane] = wi - w;
    if (lane == NW - 1) s_total = wi;
  }
  __syncthreads();
  if (block_total) *block_total = s_total;
  return s_warp[wid] + incl - thread_sum;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const uint32_t* __restrict__ in, size_t n,
                                                                    uint32_t* __restrict__ block_sums) {
  const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j)
    if (base + j < n) s += in[base + j];
  uint32_t total;
  (void)block_exclusive_scan_u32(s, &total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const uint32_t* in, uint32_t* out, size_t n,
                                                                   const uint32_t* __restrict__ block_offsets) {
  const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS];
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    v[j] = (base + j < n) ? in[base + j] : 0u;
    s += v[j];
  }
  uint32_t run = block_exclusive_scan_u32(s, nullptr) + (block_offsets ? block_offsets[blockIdx.x] : 0u);
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    if (base + j < n) out[base + j] = run;
    run += v[j];
  }
}

size_t scan_scratch_elems(size_t n) {
  size_t total = 1;
  while (n > (size_t)SCAN_TILE) {
    size_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    total += nb;
    n = nb;
  }
  return total;
}

int exclusive_scan_u32(const uint32_t* in, uint32_t* out, size_t n, uint32_t* scratch, cudaStream_t st) {
  if (n == 0) return IDDGCN_OK;
  const size_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
  IDD_CHECK(nb < (size_t)0x7fffffff, IDDGCN_ELIMIT, "scan too large");
  if (nb == 1) {
    scan_apply_kernel<<<1, SCAN_THREADS, 0, st>>>(in, out, n, nullptr);
    IDD_LAUNCH_CHECK();
    return IDDGCN_OK;
  }
  scan_reduce_kernel<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, n, scratch);
  IDD_LAUNCH_CHECK();
  IDD_TRY(exclusive_scan_u32(scratch, scratch, nb, scratch + nb, st));
  scan_apply_kernel<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, out, n, scratch);
  IDD_LAUNCH_CHECK();
  return IDDGCN_OK;
}

// ---------------------------------------------------------------------------
// stable LSD radix sort, 8-bit digits, warp-chunked
// ---------------------------------------------------------------------------
constexpr int RS_WARPS = 8;   // warps per block
constexpr int RS_UNROLL = 4;  // 32-key steps whose loads are issued together

RadixSortPlan radix_sort_plan(size_t n) {
  RadixSortPlan p;
  size_t max_warps = (size_t)sm_count() * 32;
  size_t want = (n + 1023) / 1024;
  if (want < 1) want = 1;
  if (want > max_warps) want = max_warps;
  size_t chunk = (n + want - 1) / want;
  chunk = (chunk + 31) / 32 * 32;
  if (chunk == 0) chunk = 32;
  p.chunk = chunk;
  p.num_warps = (int)((n + chunk - 1) / chunk);
  if (p.num_warps < 1) p.num_warps = 1;
  p.hist_elems = (size_t)256 * p.num_warps;
  p.scratch_elems = scan_scratch_elems(p.hist_elems);
  return p;
}

__global__ void __launch_bounds__(RS_WARPS * 32) rs_hist_kernel(const uint64_t* __restrict__ keys, size_t n, int shift,
                                                                 size_t chunk, uint32_t* __restrict__ hist,
                                                                 int num_warps) {
  __shared__ uint32_t s_cnt[RS_WARPS][256];
  const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
  const int w = blockIdx.x * RS_WARPS + wl;
  if (w >= num_warps) return;  // whole warp exits together; no block-level sync below
#pragma unroll
  for (int d = lane; d < 256; d += 32) s_cnt[wl][d] = 0;
  __syncwarp();
  const size_t begin = (size_t)w * chunk;
  const size_t end = (begin + chunk < n) ? begin + chunk : n;
  for (size_t i = begin + lane; i < end; i += 32) {
    const uint32_t d = (uint32_t)(keys[i] >> shift) & 255u;
    atomicAdd(&s_cnt[wl][d], 1u);
  }
  __syncwarp();
#pragma unroll
  for (int d = lane; d < 256; d += 32) hist[(size_t)d * num_warps + w] = s_cnt[wl][d];
}

template <bool HAS_VAL>
__global__ void __launch_bounds__(RS_WARPS * 32)
    rs_scatter_kernel(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                      uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, size_t n, int shift,
                      size_t chunk, const uint32_t* __restrict__ hist_scanned, int num_warps) {
  __shared__ uint32_t s_base[RS_WARPS][256];
  const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
  const int w = blockIdx.x * RS_WARPS + wl;
  if (w >= num_warps) return;
#pragma unroll
  for (int d = lane; d < 256; d += 32) s_base[wl][d] = hist_scanned[(size_t)d * num_warps + w];
  __syncwarp();
  const size_t begin = (size_t)w * chunk;
  const size_t end = (begin + chunk < n) ? begin + chunk : n;
  const uint32_t lt_mask = (1u << lane) - 1u;
  for (size_t i = begin; i < end; i += 32 * RS_UNROLL) {
    uint64_t k[RS_UNROLL];
    uint32_t v[RS_UNROLL];
#pragma unroll
    for (int u = 0; u < RS_UNROLL; ++u) {
      const size_t idx = i + (size_t)u * 32 + lane;
      k[u] = (idx < end) ? keys_in[idx] : 0ull;
      if (HAS_VAL) v[u] = (idx < end) ? vals_in[idx] : 0u;
    }
#pragma unroll
    for (int u = 0; u < RS_UNROLL; ++u) {
      const size_t idx = i + (size_t)u * 32 + lane;
      const bool valid = idx < end;
      // invalid lanes get a private pseudo-digit so they match nobody
      const uint32_t d = valid ? ((uint32_t)(k[u] >> shift) & 255u) : (256u + (uint32_t)lane);
      const uint32_t peers = __match_any_sync(0xffffffffu, d);
      const uint32_t rank = __popc(peers & lt_mask);
      uint32_t pos = 0;
      if (valid) pos = s_base[wl][d] + rank;
      __syncwarp();
      if (valid && rank == 0) s_base[wl][d] += __popc(peers);
      __syncwarp();
      if (valid) {
        keys_out[pos] = k[u];
        if (HAS_VAL) vals_out[pos] = v[u];
      }
    }
  }
}

int radix_sort_u64(uint64_t* keys_a, uint64_t* keys_b, uint32_t* vals_a, uint32_t* vals_b, size_t n, int num_bits,
                   const RadixSortPlan& plan, uint32_t* hist, uint32_t* scratch, cudaStream_t st,
                   uint64_t** keys_sorted, uint32_t** vals_sorted) {
  uint64_t* kin = keys_a;
  uint64_t* kout = keys_b;
  uint32_t* vin = vals_a;
  uint32_t* vout = vals_b;
  const bool has_val = vals_a != nullptr;
  if (n > 1) {
    const int passes = (num_bits + 7) / 8;
    const unsigned grid = (unsigned)((plan.num_warps + RS_WARPS - 1) / RS_WARPS);
    for (int p = 0; p < passes; ++p) {
      const int shift = p * 8;
      rs_hist_kernel<<<grid, RS_WARPS * 32, 0, st>>>(kin, n, shift, plan.chunk, hist, plan.num_warps);
      IDD_LAUNCH_CHECK();
      IDD_TRY(exclusive_scan_u32(hist, hist, plan.hist_elems, scratch, st));
      if (has_val)
        rs_scatter_kernel<true><<<grid, RS_WARPS * 32, 0, st>>>(kin, vin, kout, vout, n, shift, plan.chunk, hist,
                                                               plan.num_warps);
      else
        rs_scatter_kernel<false><<<grid, RS_WARPS * 32, 0, st>>>(kin, nullptr, kout, nullptr, n, shift, plan.chunk,
                                                                hist, plan.num_warps);
      IDD_LAUNCH_CHECK();
      uint64_t* tk = kin; kin = kout; kout = tk;
      uint32_t* tv = vin; vin = vout; vout = tv;
    }
  }
  *keys_sorted = kin;
  if (vals_sorted) *vals_sorted = vin;
  return IDDGCN_OK;
}

}  // namespace iddgcn
