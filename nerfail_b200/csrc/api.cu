// Error plumbing, launch accounting and device queries behind the C ABI.
#include "common.cuh"

namespace nfb {

std::atomic<uint64_t> g_launches{0};

char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

int sm_count() {
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

}  // namespace nfb

extern "C" {

int nfb_abi_version(void) { return NFB_ABI_VERSION; }

const char* nfb_last_error(void) { return nfb::err_buf(); }

uint64_t nfb_launch_count(void) { return nfb::g_launches.load(std::memory_order_relaxed); }

int nfb_device_cc(void) {
  int dev = 0, major = 0, minor = 0;
  NFB_CUDA(cudaGetDevice(&dev));
  NFB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  NFB_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  return major * 10 + minor;
}

}  // extern "C"
