// core.cu — error state, launch accounting, device queries.
#include "common.cuh"

namespace setok {

static thread_local char t_error[512] = "";
std::atomic<uint64_t> g_launches{0};
int g_pdl = 1;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_error, sizeof t_error, fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int n = [] {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return 148;
    return v;
  }();
  return n;
}

}  // namespace setok

extern "C" const char* setok_last_error(void) { return setok::t_error; }
extern "C" int setok_abi_version(void) { return 1; }
extern "C" uint64_t setok_launch_count(void) { return setok::g_launches.load(std::memory_order_relaxed); }
extern "C" void setok_debug_set_pdl(int on) { setok::g_pdl = on; }
