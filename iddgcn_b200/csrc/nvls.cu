// nvls.cu — exchange kernels over NVSwitch multicast memory (NVLink SHARP / "NVLS") for the row-partitioned SpMM.
//
// The partitioned step needs an all-gather of feature blocks and a reduce-scatter of partial dX blocks.  With ring
// collectives each of them moves (P-1)/P of the table in BOTH link directions; with multicast memory
//   * all-gather     = every rank stores its block ONCE to the multicast address (multimem.st): the switch replicates it
//                      into every peer's buffer  -> send 1 block, receive P-1 blocks;
//   * reduce-scatter = every rank loads ITS block through the multicast address with an in-switch sum
//                      (multimem.ld_reduce.add) -> send P-1 blocks, receive 1 block;
// so the two use opposite directions of the links and can run concurrently.  The multicast mapping itself is set up by
// the host layer (torch.distributed._symmetric_memory: cuMulticast* + handle exchange); these kernels only need the
// multicast virtual address.  Completion/ordering across ranks is the caller's job (symmetric-memory barriers).
#include "common.cuh"

namespace iddgcn {

constexpr int MC_UNROLL = 8;  // independent 16-byte multimem operations in flight per thread

__global__ void __launch_bounds__(512) mc_store_kernel(float4* __restrict__ mc_dst, const float4* __restrict__ src,
                                                       int64_t n4) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (MC_UNROLL - 1) * stride < n4; i += MC_UNROLL * stride) {
    float4 v[MC_UNROLL];
#pragma unroll
    for (int u = 0; u < MC_UNROLL; ++u) v[u] = __ldg(src + i + u * stride);
#pragma unroll
    for (int u = 0; u < MC_UNROLL; ++u)
      asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_dst + i + u * stride),
                   "f"(v[u].x), "f"(v[u].y), "f"(v[u].z), "f"(v[u].w)
                   : "memory");
  }
  for (; i < n4; i += stride) {
    const float4 v = __ldg(src + i);
    asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_dst + i), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
  }
}

__global__ void __launch_bounds__(512) mc_reduce_load_kernel(const float4* __restrict__ mc_src,
                                                             float4* __restrict__ dst, int64_t n4) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (MC_UNROLL - 1) * stride < n4; i += MC_UNROLL * stride) {
    float4 v[MC_UNROLL];
#pragma unroll
    for (int u = 0; u < MC_UNROLL; ++u)
      asm volatile("multimem.ld_reduce.weak.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w)
                   : "l"(mc_src + i + u * stride)
                   : "memory");
#pragma unroll
    for (int u = 0; u < MC_UNROLL; ++u) dst[i + u * stride] = v[u];
  }
  for (; i < n4; i += stride) {
    float4 v;
    asm volatile("multimem.ld_reduce.weak.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(mc_src + i)
                 : "memory");
    dst[i] = v;
  }
}

}  // namespace iddgcn

using namespace iddgcn;

extern "C" {

// Broadcast `n` floats from `src` (local) into every rank's copy of the symmetric buffer: mc_dst is the MULTICAST
// address of the destination range.  n % 4 == 0, 16-byte aligned.  `ctas` bounds the SMs taken from the SpMM kernels.
int iddgcn_mc_store(void* mc_dst, const float* src, int64_t n, int ctas, void* stream) {
  IDD_CHECK(mc_dst && src, IDDGCN_EINVAL, "NULL pointer");
  IDD_CHECK(n % 4 == 0 && ((uintptr_t)mc_dst & 15u) == 0 && ((uintptr_t)src & 15u) == 0, IDDGCN_EINVAL,
            "multicast copies need 16-byte alignment and n %% 4 == 0");
  if (n == 0) return IDDGCN_OK;
  if (ctas <= 0) ctas = 32;
  mc_store_kernel<<<ctas, 512, 0, (cudaStream_t)stream>>>(reinterpret_cast<float4*>(mc_dst),
                                                         reinterpret_cast<const float4*>(src), n / 4);
  IDD_LAUNCH_CHECK();
  return IDDGCN_OK;
}

// dst[i] = Σ_ranks buffer_rank[i] for the range whose MULTICAST address is mc_src (in-switch reduction).
int iddgcn_mc_reduce_load(const void* mc_src, float* dst, int64_t n, int ctas, void* stream) {
  IDD_CHECK(mc_src && dst, IDDGCN_EINVAL, "NULL pointer");
  IDD_CHECK(n % 4 == 0 && ((uintptr_t)mc_src & 15u) == 0 && ((uintptr_t)dst & 15u) == 0, IDDGCN_EINVAL,
            "multicast copies need 16-byte alignment and n %% 4 == 0");
  if (n == 0) return IDDGCN_OK;
  if (ctas <= 0) ctas = 32;
  mc_reduce_load_kernel<<<ctas, 512, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(mc_src),
                                                               reinterpret_cast<float4*>(dst), n / 4);
  IDD_LAUNCH_CHECK();
  return IDDGCN_OK;
}

}  // extern "C"
