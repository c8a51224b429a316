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

// Compile-time C: the channel column of PX pixels stays in registers -- one load and one expf per element.
template <int C, int PX>
__device__ __forceinline__ void softmax_column(const float* p, int64_t HW, bool rev, float (&e)[C][PX]) {
#pragma unroll
  for (int c = 0; c < C; ++c) load_px<PX>(p + c * HW, rev, e[c]);
  float m[PX], s[PX];
#pragma unroll
  for (int j = 0; j < PX; ++j) {
    m[j] = e[0][j];
#pragma unroll
    for (int c = 1; c < C; ++c) m[j] = fmaxf(m[j], e[c][j]);
    s[j] = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      e[c][j] = expf(e[c][j] - m[j]);
      s[j] += e[c][j];
    }
#pragma unroll
    for (int c = 0; c < C; ++c) e[c][j] = __fdiv_rn(e[c][j], s[j]);
  }
}

template <int C, int PX>
__global__ void __launch_bounds__(kValThreads, 2) k_softmax_flip_sum_reg(const float* __restrict__ z0,
                                                                      const float* __restrict__ z1,
                                                                      float* __restrict__ out, int B, int h, int w) {
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
    float p0[C][PX];
    softmax_column<C, PX>(z0 + img + x, HW, false, p0);
    if (z1) {
      float p1[C][PX];
      softmax_column<C, PX>(z1 + img + (w - PX - x), HW, true, p1);
#pragma unroll
      for (int c = 0; c < C; ++c)
#pragma unroll
        for (int j = 0; j < PX; ++j) p0[c][j] = __fadd_rn(p0[c][j], p1[c][j]);
    }
    float* o = out + img + x;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      if (PX == 2) {
        __stcs(reinterpret_cast<float2*>(o + c * HW), make_float2(p0[c][0], p0[c][PX - 1]));
      } else {
        o[c * HW] = p0[c][0];
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

constexpr int kRowsV = 4;  // output rows per thread; adjacent lanes own adjacent columns, so every source load of a warp
                           // covers <= 32 consecutive floats (4 px wide per thread: lanes 3 floats apart, 3x the wavefronts)

// Per-thread interpolation set-up of one scale: none of it depends on the channel.  The first version recomputed it inside the
// channel loop and was instruction-bound at 52 instructions per (pixel, channel) (ncu: issue active 74 %, 1.1 TB/s).
struct ScaleTaps {
  int off0[kRowsV];   // h1 * ws + w1
  int offp[kRowsV];   // ws if h1 < hs - 1 else 0
  float h0l[kRowsV], h1l[kRowsV];
  int w1p;
  float w0l, w1l;
};

__device__ __forceinline__ void make_taps(const ScaleSet& sc, int s, int x, int y0, int H, ScaleTaps& t) {
  // ATen upsample_bilinear2d (align_corners=True): w1r = rwidth * x; w1 = (int)w1r; w1p = w1 < ws - 1; l1 = w1r - w1; l0 = 1 - l1
  const int hs = sc.h[s], ws = sc.w[s];
  const float w1r = __fmul_rn(sc.rw[s], static_cast<float>(x));
  const int w1 = static_cast<int>(w1r);
  t.w1p = w1 < ws - 1 ? 1 : 0;
  t.w1l = __fsub_rn(w1r, static_cast<float>(w1));
  t.w0l = __fsub_rn(1.0f, t.w1l);
#pragma unroll
  for (int j = 0; j < kRowsV; ++j) {
    const int y = min(y0 + j, H - 1);
    const float h1r = __fmul_rn(sc.rh[s], static_cast<float>(y));
    const int h1 = static_cast<int>(h1r);
    t.h1l[j] = __fsub_rn(h1r, static_cast<float>(h1));
    t.h0l[j] = __fsub_rn(1.0f, t.h1l[j]);
    t.off0[j] = h1 * ws + w1;
    t.offp[j] = h1 < hs - 1 ? ws : 0;
  }
}

__device__ __forceinline__ void tap_channel(const float* __restrict__ plane, const ScaleTaps& t, bool first,
                                            float (&acc)[kRowsV]) {
  float v[kRowsV][4];
#pragma unroll
  for (int j = 0; j < kRowsV; ++j) {  // all loads first
    const float* r0 = plane + t.off0[j];
    const float* r1 = r0 + t.offp[j];
    v[j][0] = __ldg(r0);
    v[j][1] = __ldg(r0 + t.w1p);
    v[j][2] = __ldg(r1);
    v[j][3] = __ldg(r1 + t.w1p);
  }
#pragma unroll
  for (int j = 0; j < kRowsV; ++j) {
    const float top = __fmaf_rn(t.w0l, v[j][0], __fmul_rn(t.w1l, v[j][1]));
    const float bot = __fmaf_rn(t.w0l, v[j][2], __fmul_rn(t.w1l, v[j][3]));
    const float val = __fmaf_rn(t.h0l[j], top, __fmul_rn(t.h1l[j], bot));
    acc[j] = first ? val : __fadd_rn(acc[j], val);
  }
}

// NS = 1..3: the taps of every scale live in registers; NS = 0: any number of scales, taps rebuilt per (channel, scale).
template <int NS>
__global__ void __launch_bounds__(kValThreads, NS >= 2 ? 2 : 3) k_probs_upsample_argmax(ScaleSet sc, int B, int C, int H, int W,
                                                                                       uint8_t* __restrict__ label) {
  constexpr int kS = NS > 0 ? NS : 1;
  const int rgroups = (H + kRowsV - 1) / kRowsV;
  const int64_t n = static_cast<int64_t>(B) * rgroups * W;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(kValThreads) + threadIdx.x; idx < n;
       idx += static_cast<int64_t>(gridDim.x) * kValThreads) {
    const int x = static_cast<int>(idx % W);
    const int64_t rg = idx / W;
    const int y0 = static_cast<int>(rg % rgroups) * kRowsV;
    const int b = static_cast<int>(rg / rgroups);
    ScaleTaps taps[kS];
    const float* plane[kS];
    int plane_sz[kS];
    if (NS > 0) {
#pragma unroll
      for (int s = 0; s < kS; ++s) {
        make_taps(sc, s, x, y0, H, taps[s]);
        plane_sz[s] = sc.h[s] * sc.w[s];
        plane[s] = sc.p[s] + static_cast<int64_t>(b) * C * plane_sz[s];
      }
    }
    float best[kRowsV];
    int arg[kRowsV];
#pragma unroll
    for (int j = 0; j < kRowsV; ++j) {
      best[j] = -INFINITY;
      arg[j] = 0;
    }
#pragma unroll 1
    for (int c = 0; c < C; ++c) {
      float acc[kRowsV];
      if (NS > 0) {
#pragma unroll
        for (int s = 0; s < kS; ++s) {
          tap_channel(plane[s], taps[s], s == 0, acc);
          plane[s] += plane_sz[s];
        }
      } else {
        for (int s = 0; s < sc.n; ++s) {
          make_taps(sc, s, x, y0, H, taps[0]);
          tap_channel(sc.p[s] + (static_cast<int64_t>(b) * C + c) * sc.h[s] * sc.w[s], taps[0], s == 0, acc);
        }
      }
#pragma unroll
      for (int j = 0; j < kRowsV; ++j) {
        if (acc[j] > best[j] || (c == 0)) {
          best[j] = acc[j];
          arg[j] = c;
        }
      }
    }
    uint8_t* o = label + (static_cast<int64_t>(b) * H + y0) * W + x;
#pragma unroll
    for (int j = 0; j < kRowsV; ++j)
      if (y0 + j < H) o[static_cast<int64_t>(j) * W] = static_cast<uint8_t>(arg[j]);
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
  if (vec && C == 19)
    k_softmax_flip_sum_reg<19, 2><<<grid, kValThreads, 0, as_stream(stream)>>>(logits, logits_of_flipped, probs, B, h, w);
  else if (vec)
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
  const int64_t n = static_cast<int64_t>(B) * ((H + kRowsV - 1) / kRowsV) * W;
  const int grid = static_cast<int>(std::min<int64_t>((n + kValThreads - 1) / kValThreads, static_cast<int64_t>(sm_count()) * 16));
  cudaStream_t st = as_stream(stream);
  switch (n_scales) {
    case 1: k_probs_upsample_argmax<1><<<grid, kValThreads, 0, st>>>(sc, B, C, H, W, label); break;
    case 2: k_probs_upsample_argmax<2><<<grid, kValThreads, 0, st>>>(sc, B, C, H, W, label); break;
    case 3: k_probs_upsample_argmax<3><<<grid, kValThreads, 0, st>>>(sc, B, C, H, W, label); break;
    default: k_probs_upsample_argmax<0><<<grid, kValThreads, 0, st>>>(sc, B, C, H, W, label); break;
  }
  HIAST_CHECK_LAUNCH();
  return HIAST_OK;
}
