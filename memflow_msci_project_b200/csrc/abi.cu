// abi.cu -- library-level pieces of the C ABI (include/memflow_b200.h): error string, counters.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace mf {

static thread_local char t_error[512] = "";
std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_error, sizeof t_error, fmt, ap);
  va_end(ap);
}

bool tma_enabled() {
  static const bool on = [] {
    const char* e = getenv("MF_NO_TMA");
    return !(e && e[0] == '1');
  }();
  return on;
}

}  // namespace mf

extern "C" {

int mf_abi_version(void) { return MF_ABI_VERSION; }
const char* mf_last_error(void) { return mf::t_error; }
unsigned long long mf_launch_count(void) { return mf::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
