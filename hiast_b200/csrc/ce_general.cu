// Cross-entropy with class weights and / or a refer_labels region (the remaining `CE` switches).
//
// Reference: sseg/models/modules/losses.py:32-36 (`ce`) through compute_loss (:68-72) and compute_loss_by_selected_pixel
// (:75-89).  Two forms:
//   refer_labels is None: nn.CrossEntropyLoss(ignore_index, weight=w): sum_{y != ignore} w[y] * nll / sum_{y != ignore} w[y]
//   refer_labels given:   L = nn.CrossEntropyLoss(weight=w, reduction='none')(logits, labels)   [B,H,W]
//                         mask = region(refer_labels).unsqueeze(1)                              [B,1,H,W]
//                         (L * mask) broadcasts to [B,B,H,W]: out[b',b] = L[b] * mask[b']   (the reference's quirk, kept)
//                         loss = out.sum() / (out != 0).sum()
// No HIAST config reaches either form (the segmentors call seg_loss_fun(logits, labels) only), so these kernels are plain:
// one thread per pixel position walks the batch, log-softmax in ATen's float32 order, float64 accumulation, per-CTA
// partials reduced in a fixed order by a second kernel (deterministic).  The plain CE of the hot path stays in loss.cu.
#include "common.cuh"

namespace hiast {
namespace {

constexpr int kCeThreads = 256;

__device__ __forceinline__ long long load_label(const void* p, int bytes, int64_t i) {
  return bytes == 1 ? static_cast<long long>(static_cast<const uint8_t*>(p)[i]) : static_cast<const long long*>(p)[i];
}

__device__ __forceinline__ bool region_mask(int region, long long refer, int ignore_index) {
  if (region == HIAST_REGION_IGNORED) return refer == ignore_index;
  if (region == HIAST_REGION_CONFIDENT) return refer != ignore_index;
  return true;
}

// log-sum-exp pieces of one pixel (ATen log_softmax: x - max - log(sum exp(x - max)))
__device__ __forceinline__ void pixel_lse(const float* __restrict__ z, int C, int64_t HW, float* mx, float* lg) {
  float m = z[0];
  for (int c = 1; c < C; ++c) m = fmaxf(m, z[c * HW]);
  float s = 0.f;
  for (int c = 0; c < C; ++c) s += expf(z[c * HW] - m);
  *mx = m;
  *lg = logf(s);
}

struct CeArgs {
  const float* z;
  const void* labels;
  const float* w;
  const void* refer;
  int label_bytes, refer_bytes, region, ignore_index, B, C;
  int64_t HW;
};

__device__ __forceinline__ double block_sum(double v, double* s_tmp) {
  v = warp_sum(v);
  __syncthreads();
  if (lane_id() == 0) s_tmp[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0;
  for (int k = 0; k < kCeThreads / 32; ++k) t += s_tmp[k];
  return t;
}

__global__ void __launch_bounds__(kCeThreads) k_ce_general_fwd(CeArgs a, double* __restrict__ partial) {
  __shared__ double s_tmp[kCeThreads / 32];
  double lsum = 0, wsum = 0, cnt = 0;
  for (int64_t p = blockIdx.x * static_cast<int64_t>(kCeThreads) + threadIdx.x; p < a.HW;
       p += static_cast<int64_t>(gridDim.x) * kCeThreads) {
    double pix_sum = 0, pix_w = 0;
    int pix_nnz = 0, msum = 0;
    for (int b = 0; b < a.B; ++b) {
      const int64_t i = static_cast<int64_t>(b) * a.HW + p;
      const long long y = load_label(a.labels, a.label_bytes, i);
      if (a.refer) msum += region_mask(a.region, load_label(a.refer, a.refer_bytes, i), a.ignore_index) ? 1 : 0;
      const bool valid = a.refer ? (y >= 0 && y < a.C) : (y != a.ignore_index && y >= 0 && y < a.C);
      if (!valid) continue;
      const float* zp = a.z + static_cast<int64_t>(b) * a.C * a.HW + p;
      float m, lg;
      pixel_lse(zp, a.C, a.HW, &m, &lg);
      const float nll = -((zp[y * a.HW] - m) - lg);
      const float wy = a.w ? a.w[y] : 1.0f;
      const float l = wy * nll;
      pix_sum += l;
      pix_w += wy;
      pix_nnz += l != 0.f ? 1 : 0;
    }
    if (a.refer) {
      lsum += pix_sum * msum;
      cnt += static_cast<double>(pix_nnz) * msum;
    } else {
      lsum += pix_sum;
      wsum += pix_w;
    }
  }
  const double t0 = block_sum(lsum, s_tmp), t1 = block_sum(wsum, s_tmp), t2 = block_sum(cnt, s_tmp);
  if (threadIdx.x == 0) {
    partial[3 * blockIdx.x + 0] = t0;
    partial[3 * blockIdx.x + 1] = t1;
    partial[3 * blockIdx.x + 2] = t2;
  }
}

__global__ void k_ce_general_finalize(const double* __restrict__ partial, int n, double* sums, long long* count) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double t0 = 0, t1 = 0, t2 = 0;
  for (int i = 0; i < n; ++i) {
    t0 += partial[3 * i];
    t1 += partial[3 * i + 1];
    t2 += partial[3 * i + 2];
  }
  sums[0] = t0;
  sums[1] = t1;
  count[0] = static_cast<long long>(t2);
}

__global__ void __launch_bounds__(kCeThreads) k_ce_general_bwd(CeArgs a, const float* __restrict__ scale,
                                                               float* __restrict__ grad) {
  const float sc = scale[0];
  for (int64_t p = blockIdx.x * static_cast<int64_t>(kCeThreads) + threadIdx.x; p < a.HW;
       p += static_cast<int64_t>(gridDim.x) * kCeThreads) {
    int msum = 1;
    if (a.refer) {
      msum = 0;
      for (int b = 0; b < a.B; ++b)
        msum += region_mask(a.region, load_label(a.refer, a.refer_bytes, static_cast<int64_t>(b) * a.HW + p), a.ignore_index);
    }
    for (int b = 0; b < a.B; ++b) {
      const int64_t i = static_cast<int64_t>(b) * a.HW + p;
      const long long y = load_label(a.labels, a.label_bytes, i);
      const bool valid = a.refer ? (y >= 0 && y < a.C) : (y != a.ignore_index && y >= 0 && y < a.C);
      const float* zp = a.z + static_cast<int64_t>(b) * a.C * a.HW + p;
      float* gp = grad + static_cast<int64_t>(b) * a.C * a.HW + p;
      const float coef = valid ? sc * (a.w ? a.w[y] : 1.0f) * static_cast<float>(msum) : 0.f;
      if (coef == 0.f) {
        for (int c = 0; c < a.C; ++c) gp[c * a.HW] = 0.f;
        continue;
      }
      float m, lg;
      pixel_lse(zp, a.C, a.HW, &m, &lg);
      for (int c = 0; c < a.C; ++c) {
        const float pc = expf((zp[c * a.HW] - m) - lg);
        gp[c * a.HW] = coef * (pc - (c == y ? 1.0f : 0.0f));
      }
    }
  }
}

inline int ce_grid(int64_t HW) {
  const int64_t want = (HW + kCeThreads - 1) / kCeThreads;
  return static_cast<int>(std::min<int64_t>(want, static_cast<int64_t>(sm_count()) * 8));
}

inline bool ce_args_ok(const CeArgs& a) {
  if (!a.z || !a.labels || a.B < 0 || a.C < 1 || a.C > HIAST_MAX_CLASSES || a.HW < 1) return false;
  if (a.label_bytes != 1 && a.label_bytes != 8) return false;
  if (a.refer && a.refer_bytes != 1 && a.refer_bytes != 8) return false;
  if (a.refer && (a.region < 0 || a.region > 2)) return false;
  return true;
}

}  // namespace
}  // namespace hiast

using namespace hiast;

extern "C" size_t hiast_ce_general_workspace_bytes(int64_t HW) {
  if (HW < 1) return 0;
  return static_cast<size_t>(ce_grid(HW)) * 3 * sizeof(double);
}

extern "C" int hiast_ce_general_fwd(const float* logits, const void* labels, int label_bytes, const float* class_weights,
                                    const void* refer_labels, int refer_bytes, int region, int ignore_index, int B, int C,
                                    int64_t HW, double* sums, int64_t* count, void* workspace, size_t workspace_bytes,
                                    void* stream) {
  CeArgs a{logits, labels, class_weights, refer_labels, label_bytes, refer_bytes, region, ignore_index, B, C, HW};
  if (!ce_args_ok(a) || !sums || !count || !workspace) return HIAST_ERR_INVALID_ARG;
  const int grid = ce_grid(HW);
  if (workspace_bytes < static_cast<size_t>(grid) * 3 * sizeof(double)) return HIAST_ERR_WORKSPACE;
  double* partial = static_cast<double*>(workspace);
  k_ce_general_fwd<<<grid, kCeThreads, 0, as_stream(stream)>>>(a, partial);
  HIAST_CHECK_LAUNCH();
  k_ce_general_finalize<<<1, 32, 0, as_stream(stream)>>>(partial, grid, sums, reinterpret_cast<long long*>(count));
  HIAST_CHECK_LAUNCH();
  return HIAST_OK;
}

extern "C" int hiast_ce_general_bwd(const float* logits, const void* labels, int label_bytes, const float* class_weights,
                                    const void* refer_labels, int refer_bytes, int region, int ignore_index, int B, int C,
                                    int64_t HW, const float* scale, float* grad_logits, void* stream) {
  CeArgs a{logits, labels, class_weights, refer_labels, label_bytes, refer_bytes, region, ignore_index, B, C, HW};
  if (!ce_args_ok(a) || !scale || !grad_logits) return HIAST_ERR_INVALID_ARG;
  if (B == 0) return HIAST_OK;
  k_ce_general_bwd<<<ce_grid(HW), kCeThreads, 0, as_stream(stream)>>>(a, scale, grad_logits);
  HIAST_CHECK_LAUNCH();
  return HIAST_OK;
}
