// Validator: multi-scale / flip softmax sum and the final arg-max.
//
// Reference: workflows/validator.py:34-55 (get_multi_scale_and_flip_logits) and :92-93.  Per scale the reference runs
//   pred = softmax(model(x)); pred += flip(softmax(model(flip(x)))); pred = interpolate(pred, image size, bilinear,
//   align_corners=True); results = sum(pred over scales); lbls_pred = results.argmax(1)
// as 7+ full-tensor ATen kernels per scale, with the C x H x W probability tensors written and re-read each time
// (768x1536 -> 1024x2048, C = 19: ~700 MB of traffic per image and scale).  Two kernels replace the chain:
//   k_softmax_flip_sum        probs = softmax(z) [+ softmax(z_flipped) mirrored in x]      (read 1-2x, write 1x)
//   k_probs_upsample_argmax   label = first-index argmax_c sum_s bilinear(probs_s)(y, x)   (read 1x, write 1 B/px)
// Arithmetic is ATen's, operation for operation: sequential max / sum of expf(x - max) / IEEE division (spatial softmax
// forward); the bilinear form read off the sm_100 SASS of upsample_bilinear2d_out_frame (see ias_upsample.cu); scales
// summed in list order (Python's sum: 0 + p0 + p1 ...); argmax keeps the smallest index among equal maxima.
#include "common.cuh"

namespace hiast {
namespace {

constexpr int kValThreads = 256;
constexpr int kMaxScales = 8;

template <int PX>
__device__ __forceinline__ void load_px(const float* p, bool reversed, float (&v)[PX]);

template <>
__device__ __forceinline__ void load_px<1>(const float* p, bool, float (&v)[1]) { v[0] = __ldg(p); }

template <>
__device__ __forceinline__ void load_px<2>(const float* p, bool reversed, float (&v)[2]) {
  const float2 t = __ldg(reinterpret_cast<const float2*>(p));
  v[0] = reversed ? t.y : t.x;
  v[1] = reversed ? t.x : t.y;
}

// One source: max, sum of exponentials (ATen's order) for PX pixels whose channel column starts at p (stride HW).
template <int PX>
__device__ __forceinline__ void softmax_stats(const float* p, int C, int64_t HW, bool rev, float (&m)[PX], float (&s)[PX]) {
  float v[PX];
  load_px<PX>(p, rev, v);
#pragma unroll
  for (int j = 0; j < PX; ++j) m[j] = v[j];
  for (int c = 1; c < C; ++c) {
    load_px<PX>(p + c * HW, rev, v);
#pragma unroll
    for (int j = 0; j < PX; ++j) m[j] = fmaxf(m[j], v[j]);
  }
#pragma unroll
  for (int j = 0; j < PX; ++j) s[j] = 0.f;
  for (int c = 0; c < C; ++c) {
    load_px<PX>(p + c * HW, rev, v);
#pragma unroll
    for (int j = 0; j < PX; ++j) s[j] += expf(v[j] - m[j]);
  }
}

template <int PX>
__global__ void __launch_bounds__(kValThreads) k_softmax_flip_sum(const float* __restrict__ z0, const float* __restrict__ z1,
                                                                   float* __restrict__ out, int B, int C, int h, int w) {
  const int64_t HW = static_cast<int64_t>(h) * w;
  const int wq = w / PX;
  const int64_t n = static_cast<int64_t>(B) * h * wq;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(kValThreads) + threadIdx.x; idx < n;
       idx += static_cast<int64_t>(gridDim.x) * kValThreads) {
    const int xq = static_cast<int>(idx % wq);
    const int64_t row = idx / wq;
    const int y = static_cast<int>(row % h);
    const int b = static_cast<int>(row / h);
    const int x = xq * PX;
    const int64_t img = static_cast<int64_t>(b) * C * HW + static_cast<int64_t>(y) * w;
    const float* p0 = z0 + img + x;
    float m0[PX], s0[PX], m1[PX], s1[PX];
    softmax_stats<PX>(p0, C, HW, false, m0, s0);
    const float* p1 = nullptr;
    if (z1) {  // torch.flip(dims=[3]) of the second prediction: output x reads source w - 1 - x
      p1 = z1 + img + (w - PX - x);
      softmax_stats<PX>(p1, C, HW, true, m1, s1);
    }
    float* o = out + img + x;
    for (int c = 0; c < C; ++c) {
      float v[PX], r[PX];
      load_px<PX>(p0 + c * HW, false, v);
#pragma unroll
      for (int j = 0; j < PX; ++j) r[j] = __fdiv_rn(expf(v[j] - m0[j]), s0[j]);
      if (z1) {
        load_px<PX>(p1 + c * HW, true, v);
#pragma unroll
        for (int j = 0; j < PX; ++j) r[j] = __fadd_rn(r[j], __fdiv_rn(expf(v[j] - m1[j]), s1[j]));
      }
      if (PX == 2) {
        *reinterpret_cast<float2*>(o + c * HW) = make_float2(r[0], r[PX - 1]);
      } else {
        o[c * HW] = r[0];
      }
    }
  }
}

struct ScaleSet {
  const float* p[kMaxScales];
  int h[kMaxScales], w[kMaxScales];
  float rh[kMaxScales], rw[kMaxScales];
  int n;
};

constexpr int kPxV = 4;

__global__ void __launch_bounds__(kValThreads) k_probs_upsample_argmax(ScaleSet sc, int B, int C, int H, int W,
                                                                       uint8_t* __restrict__ label) {
  const int gpr = (W + kPxV - 1) / kPxV;
  const int64_t n = static_cast<int64_t>(B) * H * gpr;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(kValThreads) + threadIdx.x; idx < n;
       idx += static_cast<int64_t>(gridDim.x) * kValThreads) {
    const int gx = static_cast<int>(idx % gpr);
    const int64_t row = idx / gpr;
    const int y = static_cast<int>(row % H);
    const int b = static_cast<int>(row / H);
    const int x0 = gx * kPxV;
    float best[kPxV];
    int arg[kPxV];
#pragma unroll
    for (int j = 0; j < kPxV; ++j) {
      best[j] = -INFINITY;
      arg[j] = 0;
    }
    for (int c = 0; c < C; ++c) {
      float acc[kPxV];
      for (int s = 0; s < sc.n; ++s) {
        const int hs = sc.h[s], ws = sc.w[s];
        const float* plane = sc.p[s] + (static_cast<int64_t>(b) * C + c) * hs * ws;
        // ATen upsample_bilinear2d (align_corners=True): h1r = rheight * y; h1 = (int)h1r; h1p = h1 < hs - 1
        const float h1r = __fmul_rn(sc.rh[s], static_cast<float>(y));
        const int h1 = static_cast<int>(h1r);
        const int h1p = h1 < hs - 1 ? 1 : 0;
        const float h1l = __fsub_rn(h1r, static_cast<float>(h1));
        const float h0l = __fsub_rn(1.0f, h1l);
        const float* r0 = plane + static_cast<int64_t>(h1) * ws;
        const float* r1 = r0 + static_cast<int64_t>(h1p) * ws;
#pragma unroll
        for (int j = 0; j < kPxV; ++j) {
          const int x = min(x0 + j, W - 1);
          const float w1r = __fmul_rn(sc.rw[s], static_cast<float>(x));
          const int w1 = static_cast<int>(w1r);
          const int w1p = w1 < ws - 1 ? 1 : 0;
          const float w1l = __fsub_rn(w1r, static_cast<float>(w1));
          const float w0l = __fsub_rn(1.0f, w1l);
          const float top = __fmaf_rn(w0l, __ldg(r0 + w1), __fmul_rn(w1l, __ldg(r0 + w1 + w1p)));
          const float bot = __fmaf_rn(w0l, __ldg(r1 + w1), __fmul_rn(w1l, __ldg(r1 + w1 + w1p)));
          const float val = __fmaf_rn(h0l, top, __fmul_rn(h1l, bot));
          acc[j] = s == 0 ? val : __fadd_rn(acc[j], val);
        }
      }
#pragma unroll
      for (int j = 0; j < kPxV; ++j) {
        if (acc[j] > best[j] || (c == 0)) {
          best[j] = acc[j];
          arg[j] = c;
        }
      }
    }
    uint8_t* o = label + (static_cast<int64_t>(b) * H + y) * W + x0;
    if (x0 + kPxV <= W && (reinterpret_cast<uintptr_t>(o) & 3) == 0) {
      *reinterpret_cast<uint32_t*>(o) = static_cast<uint32_t>(arg[0]) | (static_cast<uint32_t>(arg[1]) << 8) |
                                        (static_cast<uint32_t>(arg[2]) << 16) | (static_cast<uint32_t>(arg[3]) << 24);
    } else {
#pragma unroll
      for (int j = 0; j < kPxV; ++j)
        if (x0 + j < W) o[j] = static_cast<uint8_t>(arg[j]);
    }
  }
}

}  // namespace
}  // namespace hiast

using namespace hiast;

extern "C" int hiast_softmax_flip_sum(const float* logits, const float* logits_of_flipped, int B, int C, int h, int w,
                                      float* probs, void* stream) {
  if (!logits || !probs || B < 0 || C < 1 || C > HIAST_MAX_CLASSES || h < 1 || w < 1) return HIAST_ERR_INVALID_ARG;
  if (B == 0) return HIAST_OK;
  auto al8 = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 8 == 0; };
  const bool vec = (w % 2 == 0) && al8(logits) && al8(probs) && (!logits_of_flipped || al8(logits_of_flipped));
  const int64_t n = static_cast<int64_t>(B) * h * (vec ? w / 2 : w);
  const int grid = static_cast<int>(std::min<int64_t>((n + kValThreads - 1) / kValThreads, static_cast<int64_t>(sm_count()) * 16));
  if (vec)
    k_softmax_flip_sum<2><<<grid, kValThreads, 0, as_stream(stream)>>>(logits, logits_of_flipped, probs, B, C, h, w);
  else
    k_softmax_flip_sum<1><<<grid, kValThreads, 0, as_stream(stream)>>>(logits, logits_of_flipped, probs, B, C, h, w);
  HIAST_CHECK_LAUNCH();
  return HIAST_OK;
}

extern "C" int hiast_probs_upsample_argmax(const float* const* probs_host, const int* h_host, const int* w_host,
                                           int n_scales, int B, int C, int H, int W, uint8_t* label, void* stream) {
  if (!probs_host || !h_host || !w_host || !label || B < 0 || H < 1 || W < 1) return HIAST_ERR_INVALID_ARG;
  if (C < 1 || C > HIAST_MAX_CLASSES) return HIAST_ERR_INVALID_ARG;
  if (n_scales < 1 || n_scales > kMaxScales) return HIAST_ERR_UNSUPPORTED;
  if (B == 0) return HIAST_OK;
  ScaleSet sc;
  sc.n = n_scales;
  for (int s = 0; s < n_scales; ++s) {
    if (!probs_host[s] || h_host[s] < 1 || w_host[s] < 1) return HIAST_ERR_INVALID_ARG;
    sc.p[s] = probs_host[s];
    sc.h[s] = h_host[s];
    sc.w[s] = w_host[s];
    // area_pixel_compute_scale<float>(in, out, align_corners=true): (in - 1) / (out - 1), 0 when out == 1
    sc.rh[s] = H > 1 ? static_cast<float>(h_host[s] - 1) / static_cast<float>(H - 1) : 0.f;
    sc.rw[s] = W > 1 ? static_cast<float>(w_host[s] - 1) / static_cast<float>(W - 1) : 0.f;
  }
  const int64_t n = static_cast<int64_t>(B) * H * ((W + kPxV - 1) / kPxV);
  const int grid = static_cast<int>(std::min<int64_t>((n + kValThreads - 1) / kValThreads, static_cast<int64_t>(sm_count()) * 16));
  k_probs_upsample_argmax<<<grid, kValThreads, 0, as_stream(stream)>>>(sc, B, C, H, W, label);
  HIAST_CHECK_LAUNCH();
  return HIAST_OK;
}
