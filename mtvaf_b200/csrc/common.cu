// Library-level plumbing of the C ABI: last-error string, device properties cache.
#include "common.cuh"
#include "../../include/mtvaf_b200.h"
#include <cstdarg>
#include <atomic>
#include <cstring>

namespace mtvaf {

static thread_local char g_last_error[1024] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

static const unsigned long long* g_step_source = nullptr;   // see mtvaf_set_step_source
const unsigned long long* step_source() { return g_step_source; }
void set_step_source(const unsigned long long* p) { g_step_source = p; }

static std::atomic<unsigned long long> g_launches{0};   // statistics only: kernels launched by this library
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static std::atomic<int> g_sm_reserve{0};   // see mtvaf_set_sm_reserve

static int sm_count_device() {
  static int n = 0;   // read-only device-properties cache
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

// SMs the persistent kernels size their grids for: all of them minus the reserve left to a concurrent collective
int sm_count() {
  const int n = sm_count_device() - g_sm_reserve.load(std::memory_order_relaxed);
  return n < 2 ? 2 : n;
}

}  // namespace mtvaf

extern "C" int mtvaf_abi_version(void) { return MTVAF_ABI_VERSION; }
extern "C" int mtvaf_set_step_source(const uint64_t* dev_step) {
  mtvaf::set_step_source(reinterpret_cast<const unsigned long long*>(dev_step));
  return 0;
}
extern "C" int mtvaf_set_sm_reserve(int n_sms) {
  if (n_sms < 0 || n_sms >= mtvaf::sm_count_device() - 1) {
    mtvaf::set_last_error("mtvaf_set_sm_reserve: %d out of range", n_sms);
    return -1;
  }
  mtvaf::g_sm_reserve.store(n_sms & ~1, std::memory_order_relaxed);   // whole TPCs (CTA pairs)
  return 0;
}
extern "C" uint64_t mtvaf_launch_count(void) { return mtvaf::g_launches.load(std::memory_order_relaxed); }
extern "C" const char* mtvaf_last_error(void) { return mtvaf::g_last_error; }
extern "C" int mtvaf_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  MTVAF_CHECK_CUDA(cudaGetDevice(&dev));
  int n = 0, ma = 0, mi = 0;
  MTVAF_CHECK_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  MTVAF_CHECK_CUDA(cudaDeviceGetAttribute(&ma, cudaDevAttrComputeCapabilityMajor, dev));
  MTVAF_CHECK_CUDA(cudaDeviceGetAttribute(&mi, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm_count) *sm_count = n;
  if (cc_major) *cc_major = ma;
  if (cc_minor) *cc_minor = mi;
  return 0;
}
