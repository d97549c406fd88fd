// Library-wide state of libfavit_b200.so: version, thread-local error text, launch counter, device query.
#include <stdarg.h>

#include <atomic>

#include "favit_common.cuh"

namespace favit {

namespace {
thread_local char g_err[512] = "";
thread_local char g_kernel[192] = "";
std::atomic<uint64_t> g_launches{0};
}  // namespace

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void note_kernel(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_kernel, sizeof(g_kernel), fmt, ap);
  va_end(ap);
}

// Driver-API calls (cuTensorMapEncodeTiled) need a context current on the CALLING thread.  A thread that has only run
// torch ops served from the caching allocator (autograd's device threads at the start of a backward pass) may not have
// one yet: cudaSetDevice binds the primary context of the thread's current device, and is legal under stream capture.
void bind_context() {
  thread_local bool bound = false;
  if (bound) return;
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess) bound = true;
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace favit

extern "C" int favit_version(void) { return 100; }  // 0.1.0

extern "C" const char* favit_last_error(void) { return favit::g_err; }

extern "C" const char* favit_last_kernel(void) { return favit::g_kernel; }

extern "C" uint64_t favit_launch_count(void) { return favit::g_launches.load(std::memory_order_relaxed); }

// Compute capability of the current device as major*10+minor (100 on B200); <0 on error.  The Python
// side refuses to run anywhere else: the kernels are sm_100a only.
extern "C" int favit_device_cc(void) {
  int dev = 0, major = 0, minor = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) return -1;
  return major * 10 + minor;
}
