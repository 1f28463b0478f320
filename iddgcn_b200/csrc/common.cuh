// common.cuh — shared helpers for libiddgcn_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <atomic>
#include "../../include/iddgcn.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libiddgcn_b200 is written for sm_100a (B200) only"
#endif

namespace iddgcn {

// ---- error plumbing -------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define IDD_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      ::iddgcn::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),        \
                          __FILE__, __LINE__);                                           \
      return (_e == cudaErrorMemoryAllocation) ? IDDGCN_ENOMEM : IDDGCN_ECUDA;           \
    }                                                                                    \
  } while (0)

#define IDD_CHECK(cond, code, ...)                                                       \
  do {                                                                                   \
    if (!(cond)) {                                                                       \
      ::iddgcn::set_error(__VA_ARGS__);                                                  \
      return (code);                                                                     \
    }                                                                                    \
  } while (0)

#define IDD_TRY(expr)                                                                    \
  do {                                                                                   \
    int _rc = (expr);                                                                    \
    if (_rc != IDDGCN_OK) return _rc;                                                    \
  } while (0)

#define IDD_LAUNCH_CHECK()                                                               \
  do {                                                                                   \
    ::iddgcn::count_launch();                                                            \
    IDD_CUDA(cudaGetLastError());                                                        \
  } while (0)

// ---- pooled device memory (core.cu) -----------------------------------------
cudaError_t pool_alloc(void** out, size_t bytes);
void pool_free(void* p);   // synchronises the device, then caches the block for reuse
void pool_trim();

// ---- device memory RAII (internal workspaces only) -------------------------
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) pool_free(p);
    p = nullptr;
    n = 0;
  }
  // grows only; contents are NOT preserved
  cudaError_t ensure(size_t count) {
    if (count <= n && p) return cudaSuccess;
    release();
    if (count == 0) count = 1;
    cudaError_t e = pool_alloc((void**)&p, count * sizeof(T));
    if (e == cudaSuccess) n = count; else p = nullptr;
    return e;
  }
  T* take() { T* q = p; p = nullptr; n = 0; return q; }
};

// One device allocation carved into 256-byte aligned pieces: cudaMalloc/cudaFree of multi-GB buffers cost more than the
// build kernels themselves, so the whole build uses ONE temporary arena and TWO persistent blocks.
struct Arena {
  char* base = nullptr;
  size_t cap = 0, off = 0;
  static size_t pad(size_t bytes) { return (bytes + 255) & ~(size_t)255; }
  template <typename T>
  T* get(size_t count) {
    T* p = reinterpret_cast<T*>(base + off);
    off += pad(count * sizeof(T));
    return off <= cap ? p : nullptr;
  }
};

inline int sm_count() {
  static int cached = 0;
  if (!cached) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    cached = n;
  }
  return cached;
}

static inline int ceil_log2_u64(uint64_t x) {  // number of bits needed to represent values in [0, x]
  int b = 0;
  while (x) { ++b; x >>= 1; }
  return b;
}

// ---- device helpers ---------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float x) {
  // same form TF/Eigen's logistic evaluates for fp32: 1/(1+exp(-x)); expf (not __expf) keeps <=2 ulp
  return 1.0f / (1.0f + expf(-x));
}

__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {  // read-once data: do not keep in L1
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ int ld_stream_i32(const int* p) {
  int r;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_f4(float4* p, float4 v) {  // write-once output: evict-first
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

}  // namespace iddgcn
