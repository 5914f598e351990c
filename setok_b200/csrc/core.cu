// core.cu — error state, launch accounting, device queries.
#include "common.cuh"

#include <mutex>
#include <set>
#include <utility>

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

// Per-device caches: the Python layer drives several GPUs from one process (torch.cuda.device(dev)), and both the SM
// count and cudaFuncSetAttribute(MaxDynamicSharedMemorySize) are properties of the CURRENT device.
static constexpr int kMaxDevices = 64;

int num_sms() {
  static std::atomic<int> cache[kMaxDevices];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 148;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n > 0) return n;
  int v = 0;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
  cache[dev].store(v, std::memory_order_relaxed);
  return v;
}

int ensure_dynamic_smem(const void* kernel, int bytes) {
  static std::mutex mu;
  static std::set<std::pair<int, const void*>> done;
  int dev = 0;
  SETOK_CUDA_OK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(mu);
  const auto key = std::make_pair(dev, kernel);
  if (done.count(key)) return SETOK_OK;
  SETOK_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  done.insert(key);
  return SETOK_OK;
}

}  // namespace setok

extern "C" const char* setok_last_error(void) { return setok::t_error; }
extern "C" int setok_abi_version(void) { return 3; }
extern "C" uint64_t setok_launch_count(void) { return setok::g_launches.load(std::memory_order_relaxed); }
extern "C" void setok_debug_set_pdl(int on) { setok::g_pdl = on; }
