// core.cu — error string, version, launch accounting.
#include <map>
#include <mutex>
#include <unordered_map>
#include "common.cuh"

namespace iddgcn {

// ---- pooled device allocations ------------------------------------------------------------------
// cudaFree (and to a lesser degree cudaMalloc) costs tens of milliseconds in a context that maps several GB, which is
// far more than the kernels of a graph / plan build take.  Handles and temporaries are therefore served from a small
// caching pool: freed blocks are kept (up to POOL_CAP bytes) and handed out again to requests of a similar size.
// pool_free synchronises the device first — the guarantee cudaFree gives — so a recycled block is never still in use.
namespace {
struct PoolBlock { size_t bytes; int device; };
std::mutex g_pool_mu;
std::unordered_map<void*, PoolBlock> g_pool_all;              // every block handed out or cached
std::multimap<size_t, void*> g_pool_free;                     // cached blocks by size
size_t g_pool_cached = 0;
constexpr size_t POOL_CAP = (size_t)24 << 30;

size_t pool_round(size_t bytes) {
  const size_t g = bytes < ((size_t)1 << 20) ? ((size_t)64 << 10) : ((size_t)2 << 20);
  return (bytes + g - 1) / g * g;
}
void pool_trim_locked() {
  for (auto& kv : g_pool_free) {
    cudaFree(kv.second);
    g_pool_all.erase(kv.second);
  }
  g_pool_free.clear();
  g_pool_cached = 0;
}
}  // namespace

cudaError_t pool_alloc(void** out, size_t bytes) {
  *out = nullptr;
  const size_t want = pool_round(bytes ? bytes : 1);
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lk(g_pool_mu);
  for (auto it = g_pool_free.lower_bound(want); it != g_pool_free.end() && it->first <= 2 * want; ++it) {
    auto rec = g_pool_all.find(it->second);
    if (rec != g_pool_all.end() && rec->second.device == dev) {
      *out = it->second;
      g_pool_cached -= it->first;
      g_pool_free.erase(it);
      return cudaSuccess;
    }
  }
  e = cudaMalloc(out, want);
  if (e == cudaErrorMemoryAllocation) {  // give cached blocks back to the driver and retry once
    (void)cudaGetLastError();
    pool_trim_locked();
    e = cudaMalloc(out, want);
  }
  if (e == cudaSuccess) g_pool_all[*out] = PoolBlock{want, dev};
  return e;
}

void pool_free(void* p) {
  if (!p) return;
  int cur = 0, owner = -1;
  cudaGetDevice(&cur);
  {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    auto rec = g_pool_all.find(p);
    if (rec != g_pool_all.end()) owner = rec->second.device;
  }
  if (owner >= 0 && owner != cur) cudaSetDevice(owner);   // wait on the device the block lives on
  cudaDeviceSynchronize();
  if (owner >= 0 && owner != cur) cudaSetDevice(cur);
  std::lock_guard<std::mutex> lk(g_pool_mu);
  auto rec = g_pool_all.find(p);
  if (rec == g_pool_all.end()) {  // not ours (should not happen): plain free
    cudaFree(p);
    return;
  }
  if (g_pool_cached + rec->second.bytes > POOL_CAP) {
    cudaFree(p);
    g_pool_all.erase(rec);
    return;
  }
  g_pool_free.emplace(rec->second.bytes, p);
  g_pool_cached += rec->second.bytes;
}

void pool_trim() {
  cudaDeviceSynchronize();
  std::lock_guard<std::mutex> lk(g_pool_mu);
  pool_trim_locked();
}

uint64_t next_handle_uid() {
  static std::atomic<uint64_t> next{1};
  return next.fetch_add(1, std::memory_order_relaxed);
}

static thread_local char t_error[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_error, sizeof(t_error), fmt, ap);
  va_end(ap);
}

}  // namespace iddgcn

extern "C" {

const char* iddgcn_last_error(void) { return iddgcn::t_error; }

const char* iddgcn_version(void) { return "iddgcn_b200 0.2.0 sm_100a"; }

int64_t iddgcn_launch_count(void) { return iddgcn::g_launches.load(std::memory_order_relaxed); }

int iddgcn_pool_trim(void) {
  iddgcn::pool_trim();
  return IDDGCN_OK;
}

}  // extern "C"
