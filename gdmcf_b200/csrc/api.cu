// Error plumbing and device queries of the C ABI (include/gdmcf_sm100.h).
#include <cstdlib>
#include "api_internal.h"

namespace gd {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t err, const char* what) {
  set_error("%s: %s (%s)", what, cudaGetErrorName(err), cudaGetErrorString(err));
  return GDMCF_ECUDA;
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GDMCF_PDL");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

static unsigned long long g_launches = 0;

int cuda_check_launch(const char* kernel) {
  ++g_launches;  // every kernel launch of the library is followed by exactly one cuda_check_launch
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return cuda_fail(err, kernel);
  return GDMCF_OK;
}

}  // namespace gd

extern "C" const char* gdmcf_last_error(void) { return gd::g_err; }

extern "C" unsigned long long gdmcf_launch_count(void) { return gd::g_launches; }

extern "C" int gdmcf_abi_version(void) { return GDMCF_ABI_VERSION; }

static int g_cc_major = -1, g_sms = -1;

static int query_device() {
  if (g_cc_major >= 0) return GDMCF_OK;
  int dev = 0;
  cudaError_t err = cudaGetDevice(&dev);
  if (err != cudaSuccess) return gd::cuda_fail(err, "cudaGetDevice");
  int major = 0, sms = 0;
  if ((err = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev)) != cudaSuccess)
    return gd::cuda_fail(err, "cudaDeviceGetAttribute(cc major)");
  if ((err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess)
    return gd::cuda_fail(err, "cudaDeviceGetAttribute(sm count)");
  g_cc_major = major;
  g_sms = sms;
  return GDMCF_OK;
}

extern "C" int gdmcf_device_check(void) {
  int rc = query_device();
  if (rc) return rc;
  if (g_cc_major != 10) {
    gd::set_error("device compute capability %d.x is not sm_100 (this library has no other code path)", g_cc_major);
    return GDMCF_EARCH;
  }
  return GDMCF_OK;
}

extern "C" int gdmcf_num_sms(void) {
  if (query_device()) return -1;
  return g_sms;
}
