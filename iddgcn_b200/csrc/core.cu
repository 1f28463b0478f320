// core.cu — error string, version, launch accounting.
#include "common.cuh"

namespace iddgcn {

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

const char* iddgcn_version(void) { return "iddgcn_b200 0.1.0 sm_100a"; }

int64_t iddgcn_launch_count(void) { return iddgcn::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
