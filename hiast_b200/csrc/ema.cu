// EMA teacher update as ONE multi-tensor launch.
//
// Reference: utils/utils.py:115-123 (update_ema_model): for every parameter pair
//   param_k = param_k * gamma + param_q * (1 - gamma)      (three float32 roundings, gamma a Python float)
// and every buffer pair  buffer_k = buffer_q.  In eager PyTorch that is five kernels per parameter tensor (two clones,
// two multiplies, one add) -- ~1500 launches per training iteration for DeepLabv2-ResNet101 -- moving 28 B per element.
// Here a table of (teacher pointer, student pointer, element count) segments is cut into fixed-size chunks; one CTA
// per chunk streams 12 B per element (read k, read q, write k) with 128-bit accesses where the segment allows it.
#include <algorithm>

#include "common.cuh"

namespace hiast {

struct EmaSeg {
  void* k;
  const void* q;
  long long n;   // elements (EMA) or bytes (copy)
};

constexpr int kThreadsE = 256;

__global__ void __launch_bounds__(kThreadsE) k_ema_update(const EmaSeg* __restrict__ segs, const int* __restrict__ chunk_seg,
                                                          const long long* __restrict__ chunk_off, int chunk_elems, float g,
                                                          float omg) {
  const EmaSeg s = segs[chunk_seg[blockIdx.x]];
  const long long off = chunk_off[blockIdx.x];
  const long long n = min(static_cast<long long>(chunk_elems), s.n - off);
  float* k = static_cast<float*>(s.k) + off;
  const float* q = static_cast<const float*>(s.q) + off;
  if (((reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(q)) & 15) == 0) {
    const long long n4 = n >> 2;
    float4* k4 = reinterpret_cast<float4*>(k);
    const float4* q4 = reinterpret_cast<const float4*>(q);
    constexpr int kU = 4;   // 8 independent 128-bit loads in flight per thread
    long long i = threadIdx.x;
    for (; i + (kU - 1) * kThreadsE < n4; i += kU * kThreadsE) {
      float4 a[kU], b[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        a[u] = k4[i + u * kThreadsE];
        b[u] = __ldcs(q4 + i + u * kThreadsE);
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        a[u].x = __fadd_rn(__fmul_rn(a[u].x, g), __fmul_rn(b[u].x, omg));
        a[u].y = __fadd_rn(__fmul_rn(a[u].y, g), __fmul_rn(b[u].y, omg));
        a[u].z = __fadd_rn(__fmul_rn(a[u].z, g), __fmul_rn(b[u].z, omg));
        a[u].w = __fadd_rn(__fmul_rn(a[u].w, g), __fmul_rn(b[u].w, omg));
        k4[i + u * kThreadsE] = a[u];
      }
    }
    for (; i < n4; i += kThreadsE) {
      float4 a = k4[i];
      const float4 b = __ldcs(q4 + i);
      a.x = __fadd_rn(__fmul_rn(a.x, g), __fmul_rn(b.x, omg));
      a.y = __fadd_rn(__fmul_rn(a.y, g), __fmul_rn(b.y, omg));
      a.z = __fadd_rn(__fmul_rn(a.z, g), __fmul_rn(b.z, omg));
      a.w = __fadd_rn(__fmul_rn(a.w, g), __fmul_rn(b.w, omg));
      k4[i] = a;
    }
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += kThreadsE)
      k[i] = __fadd_rn(__fmul_rn(k[i], g), __fmul_rn(q[i], omg));
  } else {
    for (long long i = threadIdx.x; i < n; i += kThreadsE) k[i] = __fadd_rn(__fmul_rn(k[i], g), __fmul_rn(q[i], omg));
  }
}

__global__ void __launch_bounds__(kThreadsE) k_multi_copy(const EmaSeg* __restrict__ segs, const int* __restrict__ chunk_seg,
                                                          const long long* __restrict__ chunk_off, int chunk_bytes) {
  const EmaSeg s = segs[chunk_seg[blockIdx.x]];
  const long long off = chunk_off[blockIdx.x];
  const long long n = min(static_cast<long long>(chunk_bytes), s.n - off);
  unsigned char* d = static_cast<unsigned char*>(s.k) + off;
  const unsigned char* src = static_cast<const unsigned char*>(s.q) + off;
  if (((reinterpret_cast<uintptr_t>(d) | reinterpret_cast<uintptr_t>(src)) & 15) == 0) {
    const long long n16 = n >> 4;
    for (long long i = threadIdx.x; i < n16; i += kThreadsE) reinterpret_cast<uint4*>(d)[i] = reinterpret_cast<const uint4*>(src)[i];
    for (long long i = (n16 << 4) + threadIdx.x; i < n; i += kThreadsE) d[i] = src[i];
  } else {
    for (long long i = threadIdx.x; i < n; i += kThreadsE) d[i] = src[i];
  }
}

}  // namespace hiast

using namespace hiast;

extern "C" int hiast_ema_update(const void* segs_dev, const int32_t* chunk_seg_dev, const int64_t* chunk_off_dev, int n_chunks,
                                int chunk_elems, float gamma, float one_minus_gamma, void* stream) {
  if (n_chunks < 0 || chunk_elems < 1 || chunk_elems % 4 != 0) return HIAST_ERR_INVALID_ARG;
  if (n_chunks == 0) return HIAST_OK;
  if (!segs_dev || !chunk_seg_dev || !chunk_off_dev) return HIAST_ERR_INVALID_ARG;
  k_ema_update<<<n_chunks, kThreadsE, 0, as_stream(stream)>>>(static_cast<const EmaSeg*>(segs_dev), chunk_seg_dev,
                                                               reinterpret_cast<const long long*>(chunk_off_dev), chunk_elems,
                                                               gamma, one_minus_gamma);
  HIAST_CHECK_LAUNCH();
  return HIAST_OK;
}

extern "C" int hiast_multi_copy(const void* segs_dev, const int32_t* chunk_seg_dev, const int64_t* chunk_off_dev, int n_chunks,
                                int chunk_bytes, void* stream) {
  if (n_chunks < 0 || chunk_bytes < 16 || chunk_bytes % 16 != 0) return HIAST_ERR_INVALID_ARG;
  if (n_chunks == 0) return HIAST_OK;
  if (!segs_dev || !chunk_seg_dev || !chunk_off_dev) return HIAST_ERR_INVALID_ARG;
  k_multi_copy<<<n_chunks, kThreadsE, 0, as_stream(stream)>>>(static_cast<const EmaSeg*>(segs_dev), chunk_seg_dev,
                                                               reinterpret_cast<const long long*>(chunk_off_dev), chunk_bytes);
  HIAST_CHECK_LAUNCH();
  return HIAST_OK;
}
