// dlpack.cuh — DLPack (dlpack.h v0.x ABI) argument unpacking shared by the *_dl entry points.
// The ABI structs are restated here so that no third-party header is needed.
#pragma once
#include "common.cuh"

namespace iddgcn {

typedef struct { int32_t device_type; int32_t device_id; } idd_DLDevice;
typedef struct { uint8_t code; uint8_t bits; uint16_t lanes; } idd_DLDataType;
typedef struct {
  void* data; idd_DLDevice device; int32_t ndim; idd_DLDataType dtype; int64_t* shape; int64_t* strides;
  uint64_t byte_offset;
} idd_DLTensor;
typedef struct idd_DLManagedTensor {
  idd_DLTensor dl_tensor; void* manager_ctx; void (*deleter)(struct idd_DLManagedTensor*);
} idd_DLManagedTensor;

// float32, CUDA, on `device`, compact row-major; returns the data pointer, element count and last dimension
inline int dl_f32_cuda(void* capsule_ptr, const char* name, int device, float** data, int64_t* numel,
                       int64_t* last_dim) {
  IDD_CHECK(capsule_ptr != nullptr, IDDGCN_EINVAL, "%s: NULL DLManagedTensor", name);
  const idd_DLTensor& t = reinterpret_cast<idd_DLManagedTensor*>(capsule_ptr)->dl_tensor;
  IDD_CHECK(t.device.device_type == 2 /*kDLCUDA*/, IDDGCN_EINVAL,
            "%s: DLPack tensor is not on a CUDA device (device_type=%d); there is no CPU path", name,
            t.device.device_type);
  IDD_CHECK(t.device.device_id == device, IDDGCN_EINVAL, "%s: tensor on cuda:%d but handle on cuda:%d", name,
            t.device.device_id, device);
  IDD_CHECK(t.dtype.code == 2 && t.dtype.bits == 32 && t.dtype.lanes == 1, IDDGCN_EINVAL, "%s: dtype must be float32",
            name);
  IDD_CHECK(t.ndim >= 1, IDDGCN_EINVAL, "%s: ndim must be >= 1", name);
  int64_t n = 1;
  for (int i = 0; i < t.ndim; ++i) n *= t.shape[i];
  if (t.strides) {
    int64_t expect = 1;
    for (int i = t.ndim - 1; i >= 0; --i) {
      IDD_CHECK(t.shape[i] == 1 || t.strides[i] == expect, IDDGCN_EINVAL, "%s: tensor must be compact row-major",
                name);
      expect *= t.shape[i];
    }
  }
  *data = reinterpret_cast<float*>(reinterpret_cast<char*>(t.data) + t.byte_offset);
  *numel = n;
  *last_dim = t.shape[t.ndim - 1];
  return IDDGCN_OK;
}

// same, and the element count must be `expect` (message names the expected shape)
inline int dl_f32_shaped(void* capsule_ptr, const char* name, int device, int64_t expect, const char* shape_text,
                         float** data) {
  int64_t n = 0, last = 0;
  IDD_TRY(dl_f32_cuda(capsule_ptr, name, device, data, &n, &last));
  IDD_CHECK(n == expect, IDDGCN_EINVAL, "%s must be %s (%lld elements), got %lld", name, shape_text,
            (long long)expect, (long long)n);
  return IDDGCN_OK;
}

}  // namespace iddgcn
