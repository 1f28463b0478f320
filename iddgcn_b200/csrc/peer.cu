// peer.cu — device-side plumbing of the fused exchange over NVLink peer memory (grid-partitioned SpMM, dist.py).
//
// The partitioned step has no collective call on its data path.  The backward kernel of rank p writes the partial
// dX rows that belong to rank q straight into q's stage buffer (plain stores to a peer-mapped address: the
// reduce-scatter rides in the kernel's epilogue, see spmm.cu AddSrcs), the feature blocks are pushed to the peers by
// the copy engines (iddgcn_copy_async on a side stream), and ranks order these steps with 32-bit epoch flags that
// live in symmetric memory:
//     iddgcn_peer_signal : after everything enqueued before it on the stream — publish `value` into a slot on each peer
//     iddgcn_peer_wait   : hold the stream until each of this rank's slots has reached `value`
// Both are single tiny kernels, so a step never returns to the host and ranks never meet in a host barrier.  The
// wait gives up after `timeout_ms` (sets *timed_out) instead of hanging the device if a peer died.
#include "common.cuh"

namespace iddgcn {

constexpr int MAX_PEERS = 8;
struct PeerSlots {
  uint32_t* p[MAX_PEERS];
  int n;
};

__global__ void peer_signal_kernel(PeerSlots s, uint32_t value) {
  const int i = threadIdx.x;
  if (i < s.n) {
    __threadfence_system();   // everything this stream wrote before (incl. peer stores of earlier kernels) is visible
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(s.p[i]), "r"(value) : "memory");
  }
}

__global__ void peer_wait_kernel(PeerSlots s, uint32_t value, unsigned long long timeout_ns, int* timed_out) {
  const int i = threadIdx.x;
  if (i < s.n) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (true) {
      uint32_t v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(s.p[i]) : "memory");
      if ((int32_t)(v - value) >= 0) break;
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > timeout_ns) {
        if (timed_out) *timed_out = 1;
        break;
      }
      __nanosleep(200);
    }
  }
}

// plain SM copy local -> peer (16-byte stores over NVLink); alternative to the copy-engine push for A/B runs
__global__ void __launch_bounds__(512) peer_push_kernel(float4* __restrict__ dst, const float4* __restrict__ src,
                                                        int64_t n4) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 7 * stride < n4; i += 8 * stride) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = ld_stream_f4(src + i + u * stride);
#pragma unroll
    for (int u = 0; u < 8; ++u) st_stream_f4(dst + i + u * stride, v[u]);
  }
  for (; i < n4; i += stride) st_stream_f4(dst + i, ld_stream_f4(src + i));
}

}  // namespace iddgcn

using namespace iddgcn;

extern "C" {

int iddgcn_peer_signal(void* const* remote_slots, int n, uint32_t value, void* stream) {
  IDD_CHECK(n >= 0 && n <= MAX_PEERS && (n == 0 || remote_slots), IDDGCN_EINVAL, "n must be in [0, %d]", MAX_PEERS);
  if (n == 0) return IDDGCN_OK;
  PeerSlots s{};
  s.n = n;
  for (int i = 0; i < n; ++i) {
    IDD_CHECK(remote_slots[i] != nullptr, IDDGCN_EINVAL, "slot %d is NULL", i);
    s.p[i] = static_cast<uint32_t*>(remote_slots[i]);
  }
  peer_signal_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(s, value);
  IDD_LAUNCH_CHECK();
  return IDDGCN_OK;
}

int iddgcn_peer_wait(void* const* local_slots, int n, uint32_t value, int timeout_ms, int* timed_out_dev,
                     void* stream) {
  IDD_CHECK(n >= 0 && n <= MAX_PEERS && (n == 0 || local_slots), IDDGCN_EINVAL, "n must be in [0, %d]", MAX_PEERS);
  if (n == 0) return IDDGCN_OK;
  PeerSlots s{};
  s.n = n;
  for (int i = 0; i < n; ++i) {
    IDD_CHECK(local_slots[i] != nullptr, IDDGCN_EINVAL, "slot %d is NULL", i);
    s.p[i] = static_cast<uint32_t*>(local_slots[i]);
  }
  if (timeout_ms <= 0) timeout_ms = 20000;
  peer_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(s, value, (unsigned long long)timeout_ms * 1000000ull,
                                                      timed_out_dev);
  IDD_LAUNCH_CHECK();
  return IDDGCN_OK;
}

int iddgcn_copy_async(void* dst, const void* src, int64_t bytes, void* stream) {
  IDD_CHECK(bytes >= 0 && (bytes == 0 || (dst && src)), IDDGCN_EINVAL, "NULL pointer");
  if (bytes == 0) return IDDGCN_OK;
  IDD_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, (cudaStream_t)stream));
  return IDDGCN_OK;
}

int iddgcn_copy2d_async(void* dst, int64_t dpitch, const void* src, int64_t spitch, int64_t width_bytes,
                        int64_t rows, void* stream) {
  IDD_CHECK(width_bytes >= 0 && rows >= 0 && dpitch >= width_bytes && spitch >= width_bytes, IDDGCN_EINVAL,
            "bad 2-D copy geometry");
  if (width_bytes == 0 || rows == 0) return IDDGCN_OK;
  IDD_CHECK(dst && src, IDDGCN_EINVAL, "NULL pointer");
  IDD_CUDA(cudaMemcpy2DAsync(dst, (size_t)dpitch, src, (size_t)spitch, (size_t)width_bytes, (size_t)rows,
                             cudaMemcpyDefault, (cudaStream_t)stream));
  return IDDGCN_OK;
}

int iddgcn_peer_push(float* dst_peer, const float* src, int64_t n, int ctas, void* stream) {
  IDD_CHECK(dst_peer && src, IDDGCN_EINVAL, "NULL pointer");
  IDD_CHECK(n % 4 == 0 && ((uintptr_t)dst_peer & 15u) == 0 && ((uintptr_t)src & 15u) == 0, IDDGCN_EINVAL,
            "peer copies need 16-byte alignment and n %% 4 == 0");
  if (n == 0) return IDDGCN_OK;
  if (ctas <= 0) ctas = 16;
  peer_push_kernel<<<ctas, 512, 0, (cudaStream_t)stream>>>(reinterpret_cast<float4*>(dst_peer),
                                                          reinterpret_cast<const float4*>(src), n / 4);
  IDD_LAUNCH_CHECK();
  return IDDGCN_OK;
}

}  // extern "C"
