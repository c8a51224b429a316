// (4) Evaluation confusion matrix / intersection-union as a shared-memory privatised bincount.
//
// Reference: utils/metrics.py:6-19 (intersectionAndUnionGPU: pred[target==255]=255, then three
// torch.histc over [0,K-1]); callers workflows/trainer/base_trainer.py:173-177 and
// workflows/validator.py:96-100.  A (K+1)x(K+1) int64 confusion matrix (index K = value outside
// [0,K)) over the pixels with target != ignore reproduces the three histograms as
// diag / column sums / row sums (tests/test_oracle_golden.py pins that identity).
//
// HBM stream of 2 x elem_bytes per pixel (16 B/px as the reference calls it, int64 + int64;
// 4C+8 B/px for the fused argmax-from-logits form).  Contention (most pixels land on a handful
// of diagonal cells) is removed with one shared matrix per warp and a warp-uniform fast path (match.any
// aggregation into a per-CTA matrix for K > 38); int64 global atomics only once per CTA per non-zero cell.
#include <algorithm>

#include "common.cuh"

namespace hiast {

constexpr int kThreadsM = 256;
constexpr int kMaxSharedCells = 12288;  // 48 KB of uint32

struct CmSink {
  unsigned* s;          // per-CTA shared cells or nullptr
  long long* g;         // global matrix
  unsigned* w;          // this warp's PRIVATE shared matrix or nullptr
  __device__ __forceinline__ void add(bool valid, int cell) const {
    const unsigned active = __ballot_sync(0xffffffffu, valid);
    if (w) {
      // Warp-private matrix: when every valid lane hits the same cell (coherent label maps: almost always) one lane
      // adds the count with one shared atomic; otherwise every lane issues its own shared atomic (random
      // labels: 32 different cells, no serialisation, none of match.any's cost).
      if (active == 0) return;
      const int leader = __ffs(active) - 1;
      const int c0 = __shfl_sync(0xffffffffu, cell, leader);
      if (__all_sync(0xffffffffu, !valid || cell == c0)) {
        if (lane_id() == leader) atomicAdd(w + c0, static_cast<unsigned>(__popc(active)));   // fire and forget
      } else if (valid) {
        atomicAdd(w + cell, 1u);
      }
      return;
    }
    if (!valid) return;
    const unsigned peers = __match_any_sync(active, cell);
    if (lane_id() == __ffs(peers) - 1) {
      const unsigned n = __popc(peers);
      if (s) atomicAdd(s + cell, n);
      else atomicAdd(reinterpret_cast<unsigned long long*>(g) + cell, static_cast<unsigned long long>(n));
    }
  }
};

// use_shared: 0 global atomics, 1 one matrix per CTA, 2 one matrix per warp (cells * warps <= kMaxSharedCells)
__device__ __forceinline__ void cm_init(unsigned* s_cm, int cells, int use_shared) {
  const int n = use_shared == 2 ? cells * (kThreadsM / 32) : (use_shared ? cells : 0);
  for (int i = threadIdx.x; i < n; i += kThreadsM) s_cm[i] = 0;
  __syncthreads();
}

__device__ __forceinline__ void cm_flush(const unsigned* s_cm, int cells, int use_shared, long long* cm) {
  if (!use_shared) return;
  __syncthreads();
  for (int i = threadIdx.x; i < cells; i += kThreadsM) {
    unsigned v = s_cm[i];
    if (use_shared == 2)
      for (int wi = 1; wi < kThreadsM / 32; ++wi) v += s_cm[wi * cells + i];
    if (v) atomicAdd(reinterpret_cast<unsigned long long*>(cm) + i, static_cast<unsigned long long>(v));
  }
}

__device__ __forceinline__ int clamp_class(long long v, int K) { return (v >= 0 && v < K) ? static_cast<int>(v) : K; }

template <typename T>
__global__ void __launch_bounds__(kThreadsM) k_confusion(const T* pred, const T* __restrict__ target,
                                                         long long n, int K, int ignore_index, T* pred_out,
                                                         long long* __restrict__ cm, int use_shared) {
  extern __shared__ unsigned s_cm[];
  const int cells = (K + 1) * (K + 1);
  cm_init(s_cm, cells, use_shared);
  CmSink sink = {use_shared == 1 ? s_cm : nullptr, cm, use_shared == 2 ? s_cm + (threadIdx.x >> 5) * cells : nullptr};
  // per-CTA contiguous chunk, rounded so that every warp iteration is full except the last one
  const long long per = ((n + gridDim.x - 1) / gridDim.x + kThreadsM * 4 - 1) / (kThreadsM * 4) * (kThreadsM * 4);
  const long long i0 = per * blockIdx.x;
  const long long i1 = min(n, i0 + per);
  constexpr int kUnroll = 4;   // independent load pairs in flight per thread
  for (long long base = i0; base < i1; base += kThreadsM * kUnroll) {
    long long p[kUnroll], t[kUnroll];
    bool in[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long i = base + u * kThreadsM + threadIdx.x;
      in[u] = i < i1;
      p[u] = 0;
      t[u] = ignore_index;
      if (in[u]) {
        p[u] = static_cast<long long>(__ldcs(pred + i));
        t[u] = static_cast<long long>(__ldcs(target + i));
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long i = base + u * kThreadsM + threadIdx.x;
      const bool keep = in[u] && (t[u] != ignore_index);
      if (in[u] && pred_out) pred_out[i] = keep ? static_cast<T>(p[u]) : static_cast<T>(ignore_index);
      sink.add(keep, clamp_class(t[u], K) * (K + 1) + clamp_class(p[u], K));
    }
  }
  cm_flush(s_cm, cells, use_shared, cm);
}

// pred = first-index argmax over C of logits [B,C,HW] (torch.argmax(dim=1), base_trainer.py:173)
template <typename T>
__global__ void __launch_bounds__(kThreadsM) k_confusion_logits(const float* __restrict__ logits, const T* __restrict__ target,
                                                                int B, int C, int64_t HW, int K, int ignore_index,
                                                                long long* __restrict__ cm, int use_shared) {
  extern __shared__ unsigned s_cm[];
  const int cells = (K + 1) * (K + 1);
  cm_init(s_cm, cells, use_shared);
  CmSink sink = {use_shared == 1 ? s_cm : nullptr, cm, use_shared == 2 ? s_cm + (threadIdx.x >> 5) * cells : nullptr};
  const long long n = static_cast<long long>(B) * HW;
  const long long per = ((n + gridDim.x - 1) / gridDim.x + kThreadsM - 1) / kThreadsM * kThreadsM;
  const long long i0 = per * blockIdx.x;
  const long long i1 = min(n, i0 + per);
  for (long long base = i0; base < i1; base += kThreadsM) {
    const long long i = base + threadIdx.x;
    const bool in = i < i1;
    long long t = ignore_index;
    int p = 0;
    if (in) {
      t = static_cast<long long>(target[i]);
      const int b = static_cast<int>(i / HW);
      const float* zp = logits + static_cast<size_t>(b) * C * HW + (i - static_cast<long long>(b) * HW);
      float best = __ldcs(zp);
      for (int c = 1; c < C; ++c) {
        const float v = __ldcs(zp + static_cast<size_t>(c) * HW);
        if (v > best || (v != v && best == best)) {  // torch.argmax treats NaN as the maximum
          best = v;
          p = c;
        }
      }
    }
    const bool keep = in && (t != ignore_index);
    sink.add(keep, clamp_class(t, K) * (K + 1) + clamp_class(p, K));
  }
  cm_flush(s_cm, cells, use_shared, cm);
}

__global__ void k_iou_from_cm(const long long* __restrict__ cm, int K, float* __restrict__ inter, float* __restrict__ uni) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  long long row = 0, col = 0;
  for (int j = 0; j <= K; ++j) {
    row += cm[k * (K + 1) + j];   // target == k
    col += cm[j * (K + 1) + k];   // pred == k (target not ignored)
  }
  const long long d = cm[k * (K + 1) + k];
  // the reference's values are float32 histc counts: float(I), float(O) + float(T) - float(I)
  const float fi = static_cast<float>(d);
  inter[k] = fi;
  uni[k] = __fsub_rn(__fadd_rn(static_cast<float>(col), static_cast<float>(row)), fi);
}

int cm_grid(long long n) {
  const long long want = (n + kThreadsM * 8 - 1) / (kThreadsM * 8);
  return static_cast<int>(std::max<long long>(1, std::min<long long>(want, static_cast<long long>(sm_count()) * 8)));
}

}  // namespace hiast

using namespace hiast;

extern "C" int hiast_confusion_matrix(const void* pred, const void* target, int elem_bytes, int64_t n, int K,
                                      int ignore_index, void* pred_masked_out, int64_t* cm, void* stream) {
  if (!pred || !target || !cm) return HIAST_ERR_INVALID_ARG;
  if (elem_bytes != 1 && elem_bytes != 8) return HIAST_ERR_INVALID_ARG;
  if (n < 0 || K < 1 || K > HIAST_MAX_CLASSES) return HIAST_ERR_INVALID_ARG;
  if (n == 0) return HIAST_OK;
  const int cells = (K + 1) * (K + 1);
  const int use_shared = cells * (kThreadsM / 32) <= kMaxSharedCells ? 2 : (cells <= kMaxSharedCells ? 1 : 0);
  const size_t smem = (use_shared == 2 ? cells * (kThreadsM / 32) : (use_shared ? cells : 0)) * sizeof(unsigned);
  const int grid = cm_grid(n);
  cudaStream_t st = as_stream(stream);
  if (elem_bytes == 8)
    k_confusion<long long><<<grid, kThreadsM, smem, st>>>(static_cast<const long long*>(pred),
                                                          static_cast<const long long*>(target), n, K, ignore_index,
                                                          static_cast<long long*>(pred_masked_out),
                                                          reinterpret_cast<long long*>(cm), use_shared);
  else
    k_confusion<uint8_t><<<grid, kThreadsM, smem, st>>>(static_cast<const uint8_t*>(pred),
                                                        static_cast<const uint8_t*>(target), n, K, ignore_index,
                                                        static_cast<uint8_t*>(pred_masked_out),
                                                        reinterpret_cast<long long*>(cm), use_shared);
  HIAST_CHECK_LAUNCH();
  return HIAST_OK;
}

extern "C" int hiast_confusion_from_logits(const float* logits, const void* target, int target_bytes, int B, int C,
                                           int64_t HW, int K, int ignore_index, int64_t* cm, void* stream) {
  if (!logits || !target || !cm) return HIAST_ERR_INVALID_ARG;
  if (target_bytes != 1 && target_bytes != 8) return HIAST_ERR_INVALID_ARG;
  if (B < 0 || C < 1 || HW < 1 || K < 1 || K > HIAST_MAX_CLASSES) return HIAST_ERR_INVALID_ARG;
  if (B == 0) return HIAST_OK;
  const int cells = (K + 1) * (K + 1);
  const int use_shared = cells * (kThreadsM / 32) <= kMaxSharedCells ? 2 : (cells <= kMaxSharedCells ? 1 : 0);
  const size_t smem = (use_shared == 2 ? cells * (kThreadsM / 32) : (use_shared ? cells : 0)) * sizeof(unsigned);
  const int grid = cm_grid(static_cast<long long>(B) * HW);
  cudaStream_t st = as_stream(stream);
  if (target_bytes == 8)
    k_confusion_logits<long long><<<grid, kThreadsM, smem, st>>>(logits, static_cast<const long long*>(target), B, C, HW,
                                                                 K, ignore_index, reinterpret_cast<long long*>(cm), use_shared);
  else
    k_confusion_logits<uint8_t><<<grid, kThreadsM, smem, st>>>(logits, static_cast<const uint8_t*>(target), B, C, HW, K,
                                                               ignore_index, reinterpret_cast<long long*>(cm), use_shared);
  HIAST_CHECK_LAUNCH();
  return HIAST_OK;
}

extern "C" int hiast_iou_from_confusion(const int64_t* cm, int K, float* intersection, float* area_union, void* stream) {
  if (!cm || !intersection || !area_union || K < 1 || K > HIAST_MAX_CLASSES) return HIAST_ERR_INVALID_ARG;
  k_iou_from_cm<<<(K + 63) / 64, 64, 0, as_stream(stream)>>>(reinterpret_cast<const long long*>(cm), K, intersection,
                                                             area_union);
  HIAST_CHECK_LAUNCH();
  return HIAST_OK;
}
