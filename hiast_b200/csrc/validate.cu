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

// Per-thread interpolation taps of one scale (ATen upsample_bilinear2d, align_corners=True:
// w1r = rwidth * x; w1 = (int)w1r; w1p = w1 < ws - 1; l1 = w1r - w1; l0 = 1 - l1; rows likewise).  Offsets are relative to
// (row0, col0) of a window with row pitch `pitch` (the whole plane: row0 = col0 = 0, pitch = ws).
struct ScaleTaps {
  int off0[kRowsV];   // (h1 - row0) * pitch + (w1 - col0)
  int offp[kRowsV];   // pitch if h1 < hs - 1 else 0
  float h0l[kRowsV], h1l[kRowsV];
  int w1p;
  float w0l, w1l;
};

__device__ __forceinline__ void make_taps(const ScaleSet& sc, int s, int x, int y0, int H, int row0, int col0, int pitch,
                                          ScaleTaps& t) {
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
    t.off0[j] = (h1 - row0) * pitch + (w1 - col0);
    t.offp[j] = h1 < hs - 1 ? pitch : 0;
  }
}

template <bool SHARED>
__device__ __forceinline__ void tap_channel(const float* __restrict__ plane, const ScaleTaps& t, bool first,
                                            float (&acc)[kRowsV]) {
  float v[kRowsV][4];
#pragma unroll
  for (int j = 0; j < kRowsV; ++j) {  // all loads first
    const float* r0 = plane + t.off0[j];
    const float* r1 = r0 + t.offp[j];
    if (SHARED) {
      v[j][0] = r0[0];
      v[j][1] = r0[t.w1p];
      v[j][2] = r1[0];
      v[j][3] = r1[t.w1p];
    } else {
      v[j][0] = __ldg(r0);
      v[j][1] = __ldg(r0 + t.w1p);
      v[j][2] = __ldg(r1);
      v[j][3] = __ldg(r1 + t.w1p);
    }
  }
#pragma unroll
  for (int j = 0; j < kRowsV; ++j) {
    const float top = __fmaf_rn(t.w0l, v[j][0], __fmul_rn(t.w1l, v[j][1]));
    const float bot = __fmaf_rn(t.w0l, v[j][2], __fmul_rn(t.w1l, v[j][3]));
    const float val = __fmaf_rn(t.h0l[j], top, __fmul_rn(t.h1l[j], bot));
    acc[j] = first ? val : __fadd_rn(acc[j], val);
  }
}

__device__ __forceinline__ void argmax_step(const float (&acc)[kRowsV], int c, float (&best)[kRowsV], int (&arg)[kRowsV]) {
#pragma unroll
  for (int j = 0; j < kRowsV; ++j) {
    if (acc[j] > best[j] || (c == 0)) {
      best[j] = acc[j];
      arg[j] = c;
    }
  }
}

// Direct version: every thread reads its taps from global memory (any number of scales and any ratio).  Latency-bound:
// a warp's 16 loads per (channel, scale) touch ~20 distinct sectors, so few bytes are in flight per register (ncu: 1.1 TB/s,
// long_scoreboard 14.6 once the index arithmetic is hoisted).  Used when the staged version's windows do not fit.
__global__ void __launch_bounds__(kValThreads) k_probs_upsample_argmax(ScaleSet sc, int B, int C, int H, int W,
                                                                       uint8_t* __restrict__ label) {
  const int rgroups = (H + kRowsV - 1) / kRowsV;
  const int64_t n = static_cast<int64_t>(B) * rgroups * W;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(kValThreads) + threadIdx.x; idx < n;
       idx += static_cast<int64_t>(gridDim.x) * kValThreads) {
    const int x = static_cast<int>(idx % W);
    const int64_t rg = idx / W;
    const int y0 = static_cast<int>(rg % rgroups) * kRowsV;
    const int b = static_cast<int>(rg / rgroups);
    float best[kRowsV];
    int arg[kRowsV];
    for (int c = 0; c < C; ++c) {
      float acc[kRowsV];
      for (int s = 0; s < sc.n; ++s) {
        ScaleTaps t;
        make_taps(sc, s, x, y0, H, 0, 0, sc.w[s], t);
        tap_channel<false>(sc.p[s] + (static_cast<int64_t>(b) * C + c) * sc.h[s] * sc.w[s], t, s == 0, acc);
      }
      argmax_step(acc, c, best, arg);
    }
    uint8_t* o = label + (static_cast<int64_t>(b) * H + y0) * W + x;
#pragma unroll
    for (int j = 0; j < kRowsV; ++j)
      if (y0 + j < H) o[static_cast<int64_t>(j) * W] = static_cast<uint8_t>(arg[j]);
  }
}

// Staged version: a CTA owns a tile of kRowsV output rows x 256 output columns.  The source window of the tile (<= nr_max rows
// x nc_max columns per scale and channel) is brought into shared memory with cp.async for a group of `cg` channels at a time,
// so the bytes in flight no longer depend on registers and every source element is fetched once per tile; the taps are then
// shared-memory reads.  NS = number of scales (taps of all scales live in registers).
struct StageGeom {
  int nr_max[kMaxScales], nc_max[kMaxScales];  // window bounds per scale (nc_max multiple of 4)
  int base[kMaxScales];                        // float offset of the scale's region for one channel group
  int per_channel;                             // floats per channel over all scales
  int cg;                                      // channels per group
  int vec16;                                   // 16-byte cp.async allowed (every ws % 4 == 0, planes 16-byte aligned)
};

__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem))), "l"(gmem));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem))), "l"(gmem));
}

template <int NS>
__global__ void __launch_bounds__(kValThreads) k_probs_upsample_argmax_staged(ScaleSet sc, StageGeom sg, int B, int C, int H,
                                                                              int W, uint8_t* __restrict__ label) {
  extern __shared__ __align__(128) float s_win[];
  const int tiles_x = (W + kValThreads - 1) / kValThreads;
  const int rgroups = (H + kRowsV - 1) / kRowsV;
  int tile = blockIdx.x;
  const int tx = tile % tiles_x;
  tile /= tiles_x;
  const int y0 = (tile % rgroups) * kRowsV;
  const int b = tile / rgroups;
  const int x0 = tx * kValThreads;
  const int x = min(x0 + static_cast<int>(threadIdx.x), W - 1);
  const int x_last = min(x0 + kValThreads - 1, W - 1);
  const int y_last = min(y0 + kRowsV - 1, H - 1);

  ScaleTaps taps[NS];
  int row0[NS], col0[NS], nrows[NS], ncols[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    // window of the tile: rows h1(y0) .. h1(y_last) + 1, columns w1(x0) .. w1(x_last) + 1 (clamped), start aligned down to 4
    row0[s] = static_cast<int>(__fmul_rn(sc.rh[s], static_cast<float>(y0)));
    const int row1 = min(static_cast<int>(__fmul_rn(sc.rh[s], static_cast<float>(y_last))) + 1, sc.h[s] - 1);
    col0[s] = static_cast<int>(__fmul_rn(sc.rw[s], static_cast<float>(x0))) & ~3;
    const int col1 = min(static_cast<int>(__fmul_rn(sc.rw[s], static_cast<float>(x_last))) + 1, sc.w[s] - 1);
    nrows[s] = row1 - row0[s] + 1;
    ncols[s] = min((col1 - col0[s] + 4) & ~3, sg.nc_max[s]);
    make_taps(sc, s, x, y0, H, row0[s], col0[s], sg.nc_max[s], taps[s]);
  }
  float best[kRowsV];
  int arg[kRowsV];
  const int group_floats = sg.per_channel * sg.cg;  // one pipeline stage
  auto issue_group = [&](int c0, float* stage) {
    const int nc = min(sg.cg, C - c0);
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const int ws = sc.w[s];
      const int64_t plane_sz = static_cast<int64_t>(sc.h[s]) * ws;
      const float* src = sc.p[s] + (static_cast<int64_t>(b) * C + c0) * plane_sz + static_cast<int64_t>(row0[s]) * ws + col0[s];
      float* dst = stage + static_cast<size_t>(sg.base[s]) * sg.cg;
      const int pitch = sg.nc_max[s], cstride = sg.nr_max[s] * pitch;
      // a warp per (channel, row) of the window, lanes along the row: one division per row instead of two per element
      // (the first version spent as many instructions on this index arithmetic as on the taps)
      const int lane = lane_id(), warp = threadIdx.x >> 5;
      const int n_rows_total = nc * nrows[s];
      const int cols = min(ncols[s], ws - col0[s]);
      for (int rr = warp; rr < n_rows_total; rr += kValThreads / 32) {
        const int c = rr / nrows[s], r = rr - c * nrows[s];
        const float* src_row = src + c * plane_sz + static_cast<int64_t>(r) * ws;
        float* dst_row = dst + c * cstride + r * pitch;
        if (sg.vec16) {
          for (int k = 4 * lane; k < ncols[s]; k += 128) cp_async16(dst_row + k, src_row + k);
        } else {
          for (int k = lane; k < cols; k += 32) cp_async4(dst_row + k, src_row + k);
        }
      }
    }
    asm volatile("cp.async.commit_group;");
  };
  // two-stage pipeline over channel groups: group g + 1 is in flight while group g is consumed
  issue_group(0, s_win);
  int stage = 0;
  for (int c0 = 0; c0 < C; c0 += sg.cg) {
    const int nc = min(sg.cg, C - c0);
    if (c0 + sg.cg < C) {
      issue_group(c0 + sg.cg, s_win + (stage ^ 1) * group_floats);
      asm volatile("cp.async.wait_group 1;");
    } else {
      asm volatile("cp.async.wait_group 0;");
    }
    __syncthreads();
    const float* win = s_win + stage * group_floats;
    for (int c = 0; c < nc; ++c) {
      float acc[kRowsV];
#pragma unroll
      for (int s = 0; s < NS; ++s)
        tap_channel<true>(win + static_cast<size_t>(sg.base[s]) * sg.cg + c * (sg.nr_max[s] * sg.nc_max[s]), taps[s], s == 0, acc);
      argmax_step(acc, c0 + c, best, arg);
    }
    __syncthreads();  // the stage may be refilled by the next iteration's issue
    stage ^= 1;
  }
  if (x0 + static_cast<int>(threadIdx.x) < W) {
    uint8_t* o = label + (static_cast<int64_t>(b) * H + y0) * W + x;
#pragma unroll
    for (int j = 0; j < kRowsV; ++j)
      if (y0 + j < H) o[static_cast<int64_t>(j) * W] = static_cast<uint8_t>(arg[j]);
  }
}

template <int NS>
int launch_staged(const ScaleSet& sc, const StageGeom& sg, size_t smem, int B, int C, int H, int W, uint8_t* label,
                  cudaStream_t st) {
  HIAST_TRY(ensure_dyn_smem(k_probs_upsample_argmax_staged<NS>, smem));
  const long long tiles = static_cast<long long>(B) * ((H + kRowsV - 1) / kRowsV) * ((W + kValThreads - 1) / kValThreads);
  if (tiles > 0x7FFFFFFFll) return HIAST_ERR_UNSUPPORTED;
  k_probs_upsample_argmax_staged<NS><<<static_cast<int>(tiles), kValThreads, smem, st>>>(sc, sg, B, C, H, W, label);
  HIAST_CHECK_LAUNCH();
  return HIAST_OK;
}

}  // namespace
}  // namespace hiast

using namespace hiast;

static int g_val_direct = 0;
extern "C" int hiast_debug_validate_direct(int on) {
  g_val_direct = on;
  return HIAST_OK;
}

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
  // staged version: window bounds per scale, channel groups sized for a two-stage cp.async pipeline
  if (n_scales <= 3 && !g_val_direct) {
    StageGeom sg;
    sg.per_channel = 0;
    sg.vec16 = 1;
    for (int s = 0; s < n_scales; ++s) {
      sg.nr_max[s] = std::min(sc.h[s], static_cast<int>(static_cast<double>(sc.rh[s]) * (kRowsV - 1)) + 3);
      sg.nc_max[s] = (std::min(sc.w[s], static_cast<int>(static_cast<double>(sc.rw[s]) * (kValThreads - 1)) + 3) + 3 + 3) & ~3;
      sg.base[s] = sg.per_channel;
      sg.per_channel += sg.nr_max[s] * sg.nc_max[s];
      if (sc.w[s] % 4 != 0 || reinterpret_cast<uintptr_t>(sc.p[s]) % 16 != 0) sg.vec16 = 0;
    }
    // All channels in one shot when they fit in 100 KB (2 CTAs per SM cover each other's load phase: measured faster than
    // a two-stage pipeline over smaller channel groups, the tap phase being shared-memory bound); otherwise two stages.
    const size_t per_channel_bytes = sizeof(float) * sg.per_channel;
    size_t smem = 0;
    if (per_channel_bytes * C <= 100 * 1024) {
      sg.cg = C;
      smem = per_channel_bytes * C;
    } else {
      sg.cg = static_cast<int>(std::min<size_t>(C, 50 * 1024 / per_channel_bytes));
      if (sg.cg < 1 && per_channel_bytes * 2 <= 200 * 1024) sg.cg = 1;   // wide windows: one channel per stage
      smem = per_channel_bytes * sg.cg * 2;
    }
    if (sg.cg >= 1) {
      switch (n_scales) {
        case 1: return launch_staged<1>(sc, sg, smem, B, C, H, W, label, st);
        case 2: return launch_staged<2>(sc, sg, smem, B, C, H, W, label, st);
        default: return launch_staged<3>(sc, sg, smem, B, C, H, W, label, st);
      }
    }
  }
  k_probs_upsample_argmax<<<grid, kValThreads, 0, st>>>(sc, B, C, H, W, label);
  HIAST_CHECK_LAUNCH();
  return HIAST_OK;
}
