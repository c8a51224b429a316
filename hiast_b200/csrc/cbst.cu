// CBST policy support: class-balanced thresholds from an every-k-th sample.
//
// Reference: workflows/pseudo_label_generator.py:142-165 (CBSTPseudoGenerator.get_constant_threshold).  Per batch and
// class c the reference gathers conf[label == c] of the whole [B,H,W] batch in raster order, rounds to fp16, keeps the
// elements 0, k, 2k, ... (k = cbst.sample_interval) and appends them to one list per class; after the whole data set
// class_threshold[c] = np.quantile(list_c, 1 - cbst.p).
//
// On the GPU the "every k-th in raster order" rule needs each pixel's rank among the pixels of its class inside its
// batch: (1) per-tile per-class counts, (2) exclusive prefix over the tiles of a batch, (3) a second pass that
// rebuilds the in-tile rank and adds the pixels with rank % k == 0 to one global fp16-key histogram per class.  The
// final quantile is read off the prefix-summed histogram with numpy's 'linear' rule (no extra sample here).
#include <math.h>

#include <algorithm>

#include "common.cuh"
#include "scan_math.h"

namespace hiast {

constexpr int kThreadsB = 256;
constexpr int kPxB = 32;                          // consecutive pixels per thread
constexpr int kTileB = kThreadsB * kPxB;          // 8192 pixels per tile (tiles never straddle images)

__host__ __device__ inline int cbst_row_stride(int nb) { return (nb + 3) & ~3; }

// pass 1 (rank_pass = 0): tile_counts[tile][c] = #pixels of class c in the tile
// pass 3 (rank_pass = 1): rank of every pixel = tile_prefix[tile][c] + #earlier pixels of class c in the tile;
//                         rank % interval == 0 -> hist[c][key]++
__global__ void __launch_bounds__(kThreadsB) k_cbst_pass(const float* __restrict__ conf, const uint8_t* __restrict__ label,
                                                         int n_images, int64_t HW, int C, int tiles_per_image,
                                                         int rank_pass, uint32_t* __restrict__ tile_counts,
                                                         const uint32_t* __restrict__ tile_prefix, int interval,
                                                         int key_lo, int nb, uint32_t* __restrict__ hist) {
  extern __shared__ uint32_t s_cnt[];  // [C][kThreadsB]: private column per thread
  const int n_tiles = tiles_per_image * n_images;
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int img = t / tiles_per_image;
    const int64_t p0 = static_cast<int64_t>(t - img * tiles_per_image) * kTileB + static_cast<int64_t>(threadIdx.x) * kPxB;
    const int n_px = static_cast<int>(max(static_cast<int64_t>(0), min(static_cast<int64_t>(kPxB), HW - p0)));
    const size_t base = static_cast<size_t>(img) * HW + p0;
    for (int c = 0; c < C; ++c) s_cnt[c * kThreadsB + threadIdx.x] = 0;
    for (int j = 0; j < n_px; ++j) {
      const int l = label[base + j];
      if (l < C) s_cnt[l * kThreadsB + threadIdx.x] += 1;
    }
    __syncthreads();
    // per class: exclusive scan over the 256 thread columns (warp w takes classes w, w+8, ...)
    for (int c = threadIdx.x >> 5; c < C; c += kThreadsB / 32) {
      uint32_t* col = s_cnt + c * kThreadsB;
      uint32_t v[kThreadsB / 32];
      uint32_t run = 0;
#pragma unroll
      for (int k = 0; k < kThreadsB / 32; ++k) {      // lane owns 8 consecutive columns
        v[k] = col[lane_id() * (kThreadsB / 32) + k];
        run += v[k];
      }
      uint32_t incl = run;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane_id() >= o) incl += y;
      }
      uint32_t excl = incl - run;
      if (!rank_pass) {
        if (lane_id() == 31) tile_counts[static_cast<size_t>(t) * C + c] = incl;
      } else {
        excl += tile_prefix[static_cast<size_t>(t) * C + c];
#pragma unroll
        for (int k = 0; k < kThreadsB / 32; ++k) {
          col[lane_id() * (kThreadsB / 32) + k] = excl;
          excl += v[k];
        }
      }
    }
    __syncthreads();
    if (rank_pass) {
      for (int j = 0; j < n_px; ++j) {
        const int l = label[base + j];
        if (l < C) {
          const uint32_t rank = s_cnt[l * kThreadsB + threadIdx.x]++;
          if (rank % static_cast<uint32_t>(interval) == 0) {
            int bin = static_cast<int>(fp16_key(conf[base + j])) - key_lo;
            bin = min(max(bin, 0), nb - 1);
            atomicAdd(hist + static_cast<size_t>(l) * cbst_row_stride(nb) + bin, 1u);
          }
        }
      }
    }
    __syncthreads();
  }
}

// pass 2: exclusive prefix of the tile counts over the tiles of each batch (group of images), per class.
__global__ void k_cbst_tile_prefix(const uint32_t* __restrict__ tile_counts, uint32_t* __restrict__ tile_prefix,
                                   int n_images, int group_size, int tiles_per_image, int C) {
  const int g = blockIdx.x;
  const int c = threadIdx.x;
  if (c >= C) return;
  const int t0 = g * group_size * tiles_per_image;
  const int t1 = min(n_images, (g + 1) * group_size) * tiles_per_image;
  uint32_t run = 0;
  for (int t = t0; t < t1; ++t) {
    tile_prefix[static_cast<size_t>(t) * C + c] = run;
    run += tile_counts[static_cast<size_t>(t) * C + c];
  }
}

// class_threshold[c] = np.quantile(samples_c, q) (method 'linear') from the key histogram; float64 result.
__global__ void k_cbst_quantile(const uint32_t* __restrict__ hist, int C, int key_lo, int nb, double q,
                                double* __restrict__ thr, int* __restrict__ error_flag) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const uint32_t* row = hist + static_cast<size_t>(c) * cbst_row_stride(nb);
  long long n = 0;
  for (int b = 0; b < nb; ++b) n += row[b];
  if (n == 0) {                    // np.quantile of an empty list raises IndexError
    thr[c] = nan("");
    if (error_flag) atomicOr(error_flag, 4);
    return;
  }
  if (!(q >= 0.0 && q <= 1.0) && error_flag) atomicOr(error_flag, 1);
  const double vi = __dmul_rn(static_cast<double>(n - 1), q);
  const double fl = floor(vi);
  const double g = __dsub_rn(vi, fl);
  long long lo_i, hi_i;
  if (vi >= static_cast<double>(n - 1)) lo_i = hi_i = n - 1;
  else if (vi < 0.0) lo_i = hi_i = 0;
  else {
    lo_i = static_cast<long long>(fl);
    hi_i = lo_i + 1;
  }
  long long run = 0;
  double a = 0.0, b = 0.0;
  bool have_a = false;
  for (int k = 0; k < nb; ++k) {
    run += row[k];
    if (!have_a && run > lo_i) {
      a = half_bits_to_double(static_cast<unsigned>(key_lo + k));
      have_a = true;
    }
    if (run > hi_i) {
      b = half_bits_to_double(static_cast<unsigned>(key_lo + k));
      break;
    }
  }
  thr[c] = lerp_np(a, b, g);
}

}  // namespace hiast

using namespace hiast;

extern "C" size_t hiast_cbst_workspace_bytes(int n_images, int64_t HW, int C) {
  if (n_images < 0 || HW < 1 || C < 1) return 0;
  const int64_t tiles_per_image = (HW + kTileB - 1) / kTileB;
  return static_cast<size_t>(2) * tiles_per_image * n_images * C * sizeof(uint32_t);
}

extern "C" int hiast_cbst_sample_hist(const float* conf, const uint8_t* label, int n_images, int64_t HW, int C,
                                      int group_size, int sample_interval, int key_lo, uint32_t* hist,
                                      void* workspace, size_t workspace_bytes, void* stream) {
  if (!conf || !label || !hist || !workspace) return HIAST_ERR_INVALID_ARG;
  if (n_images < 0 || HW < 1 || C < 1 || C > 48 || group_size < 1 || sample_interval < 1) return HIAST_ERR_INVALID_ARG;
  if (key_lo < 0 || key_lo > HIAST_KEY_ONE) return HIAST_ERR_INVALID_ARG;
  if (workspace_bytes < hiast_cbst_workspace_bytes(n_images, HW, C)) return HIAST_ERR_WORKSPACE;
  if (n_images == 0) return HIAST_OK;
  cudaStream_t st = as_stream(stream);
  const int tiles_per_image = static_cast<int>((HW + kTileB - 1) / kTileB);
  const long long n_tiles = static_cast<long long>(tiles_per_image) * n_images;
  if (n_tiles >= (1ll << 30)) return HIAST_ERR_UNSUPPORTED;
  uint32_t* counts = static_cast<uint32_t*>(workspace);
  uint32_t* prefix = counts + n_tiles * C;
  const int nb = HIAST_KEY_ONE - key_lo + 1;
  const size_t smem = static_cast<size_t>(C) * kThreadsB * sizeof(uint32_t);
  const int grid = static_cast<int>(std::min<long long>(n_tiles, static_cast<long long>(sm_count()) * 4));
  k_cbst_pass<<<grid, kThreadsB, smem, st>>>(conf, label, n_images, HW, C, tiles_per_image, 0, counts, nullptr, 1, key_lo, nb,
                                             hist);
  HIAST_CHECK_LAUNCH();
  const int n_groups = (n_images + group_size - 1) / group_size;
  k_cbst_tile_prefix<<<n_groups, 64, 0, st>>>(counts, prefix, n_images, group_size, tiles_per_image, C);
  HIAST_CHECK_LAUNCH();
  k_cbst_pass<<<grid, kThreadsB, smem, st>>>(conf, label, n_images, HW, C, tiles_per_image, 1, counts, prefix, sample_interval,
                                             key_lo, nb, hist);
  HIAST_CHECK_LAUNCH();
  return HIAST_OK;
}

extern "C" int hiast_cbst_quantile(const uint32_t* hist, int C, int key_lo, double q, double* thr, int* error_flag,
                                   void* stream) {
  if (!hist || !thr || C < 1 || C > HIAST_MAX_CLASSES || key_lo < 0 || key_lo > HIAST_KEY_ONE) return HIAST_ERR_INVALID_ARG;
  k_cbst_quantile<<<(C + 31) / 32, 32, 0, as_stream(stream)>>>(hist, C, key_lo, HIAST_KEY_ONE - key_lo + 1, q, thr, error_flag);
  HIAST_CHECK_LAUNCH();
  return HIAST_OK;
}
