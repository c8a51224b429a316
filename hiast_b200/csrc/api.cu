// Library-level entry points, device queries and host-side test hooks.
#include <mutex>

#include "common.cuh"
#include "scan_math.h"

namespace hiast {

thread_local int g_last_cuda_error = 0;

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

}  // namespace hiast

extern "C" int hiast_version(void) { return 1000 * 0 + 1; }

extern "C" const char* hiast_status_string(int status) {
  switch (status) {
    case HIAST_OK: return "ok";
    case HIAST_ERR_INVALID_ARG: return "invalid argument";
    case HIAST_ERR_UNSUPPORTED: return "unsupported configuration";
    case HIAST_ERR_CUDA: return "CUDA error (see hiast_last_cuda_error)";
    case HIAST_ERR_WORKSPACE: return "workspace too small";
    default: return "unknown status";
  }
}

extern "C" int hiast_last_cuda_error(void) { return hiast::g_last_cuda_error; }

extern "C" int hiast_device_sm_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return 0;
  }
  return hiast::sm_count();
}

extern "C" double hiast_testhook_powi(double x, int n) { return hiast::powi_dd(x, n); }

extern "C" double hiast_testhook_threshold_step(const uint32_t* prefix_row_host, int key_lo, double thr, double alpha,
                                                double beta, double gamma, float* temp_out, int* error_out) {
  const int nb = HIAST_KEY_ONE - key_lo + 1;
  float temp = 0.f;
  int err = 0;
  const double r = hiast::ias_threshold_step(prefix_row_host, nb, key_lo, thr, alpha, beta, gamma, &temp, &err);
  if (temp_out) *temp_out = temp;
  if (error_out) *error_out = err;
  return r;
}
