// Shared helpers for the hiast_b200 kernels (sm_100a only).
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hiast_b200.h"
#include "../../include/hiast_b200_dev.h"

namespace hiast {

extern thread_local int g_last_cuda_error;

inline int cuda_fail(cudaError_t e) {
  g_last_cuda_error = static_cast<int>(e);
  return HIAST_ERR_CUDA;
}

#define HIAST_CUDA_TRY(expr)                                   \
  do {                                                         \
    cudaError_t _e = (expr);                                   \
    if (_e != cudaSuccess) return ::hiast::cuda_fail(_e);      \
  } while (0)

#define HIAST_CHECK_LAUNCH() HIAST_CUDA_TRY(cudaGetLastError())

inline cudaStream_t as_stream(void* s) { return static_cast<cudaStream_t>(s); }

// SM count / max resident CTAs of the current device, cached per device.
int sm_count();

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute of a kernel: remember, per (device, kernel), the largest
// size already granted and call cudaFuncSetAttribute only when a launch needs more.  Returns a HIAST status.
int ensure_dyn_smem_impl(const void* kernel, size_t bytes);
template <typename K>
inline int ensure_dyn_smem(K kernel, size_t bytes) { return ensure_dyn_smem_impl(reinterpret_cast<const void*>(kernel), bytes); }
#define HIAST_TRY(expr)                \
  do {                                 \
    const int _rc = (expr);            \
    if (_rc != HIAST_OK) return _rc;   \
  } while (0)

template <typename K>
inline int resident_grid(K kernel, int threads, size_t dyn_smem) {
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, dyn_smem) != cudaSuccess || occ < 1)
    occ = 1;
  return sm_count() * occ;
}

constexpr int kWarp = 32;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// fp16 round-to-nearest-even bit pattern of a float (the reference's .astype(np.float16)).
__device__ __forceinline__ unsigned fp16_key(float x) {
  return static_cast<unsigned>(__half_as_ushort(__float2half_rn(x)));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ long long warp_sum(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace hiast
