// ABI bookkeeping: version, error strings, device properties.
#include <string.h>

#include "common.cuh"

namespace scan {
static thread_local char g_err[512] = "";
static int g_sm_count[64] = {0};

void set_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
}

int sm_count() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  int& slot = g_sm_count[dev & 63];
  if (slot == 0) slot = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) ? n : 148;
  return slot;
}
}  // namespace scan

extern "C" {

int scan_abi_version(void) { return SCAN_ABI_VERSION; }

const char* scan_strerror(int code) {
  switch (code) {
    case SCAN_OK: return "ok";
    case SCAN_EINVAL: return "invalid argument";
    case SCAN_ECUDA: return "CUDA call failed";
    case SCAN_ENOTSUP: return "configuration not supported";
    case SCAN_ECAPACITY: return "workspace or output capacity too small";
    default: return "unknown error";
  }
}

const char* scan_last_cuda_error(void) { return scan::g_err; }

int scan_init(int device) {
  SCAN_CUDA_CHECK(cudaSetDevice(device));
  int major = 0, minor = 0;
  SCAN_CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  SCAN_CUDA_CHECK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
  if (major != 10) {
    snprintf(scan::g_err, sizeof(scan::g_err), "scan_b200 is built for sm_100a only, device is sm_%d%d", major, minor);
    return SCAN_ENOTSUP;
  }
  scan::sm_count();
  return SCAN_OK;
}
}
