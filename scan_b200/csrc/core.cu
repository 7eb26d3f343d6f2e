// ABI bookkeeping: version, error strings, device properties.
#include <string.h>

#include <algorithm>

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
// dst[i] = src[i] for a small buffer: `src` may be PINNED HOST memory (device-accessible under unified addressing), so tiny
// per-call inputs (the padded GT boxes) reach the GPU through a kernel read instead of the copy engine, where they would queue
// behind a training loop's multi-hundred-MB input prefetch
__global__ void __launch_bounds__(256) upload_small_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, long long n_words) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += (long long)gridDim.x * blockDim.x) dst[i] = src[i];
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

int scan_upload_small(const void* pinned_host_or_device_src, void* dst, int64_t bytes, void* stream) {
  if (bytes == 0) return SCAN_OK;
  if (!pinned_host_or_device_src || !dst || bytes < 0 || (bytes & 3) || bytes > (1 << 22)) return SCAN_EINVAL;
  if (((uintptr_t)pinned_host_or_device_src & 3) || ((uintptr_t)dst & 3)) return SCAN_EINVAL;
  const long long n = bytes / 4;
  const int blocks = (int)std::min<long long>((n + 255) / 256, 64);
  scan::upload_small_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const uint32_t*)pinned_host_or_device_src, (uint32_t*)dst, n);
  SCAN_LAUNCH_CHECK("upload_small_kernel");
  return SCAN_OK;
}

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
