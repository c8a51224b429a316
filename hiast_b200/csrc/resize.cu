// Nearest-neighbour resize of uint8 label maps on the device (consumer side).
//
// The reference stores pseudo-labels at the inference size (768x1536 in the sl configs) and BaseDataset.load_data
// brings them to the image size with `cv2.resize(lbl, img.shape[:-1][::-1], interpolation=cv2.INTER_NEAREST)`
// (sseg/datasets/loader/base_dataset.py:176; CopyPaste.resize, preprocessor.py:46-52, does the same for donors).
// OpenCV's rule: sx = min(floor(x * (1 / (dst_w / src_w))), src_w - 1) in double precision, same for rows.  The two
// scale factors are computed by the caller's host code exactly that way and passed in as doubles.
#include "common.cuh"

namespace hiast {
namespace {

constexpr int kResizeThreads = 256;
constexpr int kResizePx = 16;

__global__ void __launch_bounds__(kResizeThreads) k_resize_nearest_u8(const uint8_t* __restrict__ src, int Hs, int Ws,
                                                                      uint8_t* __restrict__ dst, int Hd, int Wd,
                                                                      int n_images, double ifx, double ify, int vec) {
  const int groups_per_row = (Wd + kResizePx - 1) / kResizePx;
  const long long n_groups = static_cast<long long>(n_images) * Hd * groups_per_row;
  for (long long gidx = blockIdx.x * static_cast<long long>(kResizeThreads) + threadIdx.x; gidx < n_groups;
       gidx += static_cast<long long>(gridDim.x) * kResizeThreads) {
    const int gx = static_cast<int>(gidx % groups_per_row);
    const long long rowid = gidx / groups_per_row;
    const int y = static_cast<int>(rowid % Hd);
    const int img = static_cast<int>(rowid / Hd);
    const int sy = min(static_cast<int>(floor(y * ify)), Hs - 1);
    const uint8_t* srow = src + (static_cast<size_t>(img) * Hs + sy) * Ws;
    uint8_t* drow = dst + (static_cast<size_t>(img) * Hd + y) * Wd;
    const int x0 = gx * kResizePx;
    if (vec && x0 + kResizePx <= Wd) {
      uint32_t w[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint32_t word = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int x = x0 + 4 * k + q;
          const int sx = min(static_cast<int>(floor(x * ifx)), Ws - 1);
          word |= static_cast<uint32_t>(__ldg(srow + sx)) << (8 * q);
        }
        w[k] = word;
      }
      *reinterpret_cast<uint4*>(drow + x0) = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
      for (int x = x0; x < min(Wd, x0 + kResizePx); ++x)
        drow[x] = __ldg(srow + min(static_cast<int>(floor(x * ifx)), Ws - 1));
    }
  }
}

}  // namespace
}  // namespace hiast

using namespace hiast;

extern "C" int hiast_resize_nearest_u8(const uint8_t* src, int n_images, int Hs, int Ws, uint8_t* dst, int Hd, int Wd,
                                       double inv_scale_x, double inv_scale_y, void* stream) {
  if (!src || !dst || n_images < 0 || Hs < 1 || Ws < 1 || Hd < 1 || Wd < 1) return HIAST_ERR_INVALID_ARG;
  if (!(inv_scale_x > 0.0) || !(inv_scale_y > 0.0)) return HIAST_ERR_INVALID_ARG;
  if (n_images == 0) return HIAST_OK;
  const int groups_per_row = (Wd + kResizePx - 1) / kResizePx;
  const long long n_groups = static_cast<long long>(n_images) * Hd * groups_per_row;
  const int vec = (Wd % kResizePx == 0) && (reinterpret_cast<uintptr_t>(dst) % 16 == 0);
  const long long want = (n_groups + kResizeThreads - 1) / kResizeThreads;
  const int grid = static_cast<int>(std::min<long long>(want, static_cast<long long>(sm_count()) * 32));
  k_resize_nearest_u8<<<grid, kResizeThreads, 0, as_stream(stream)>>>(src, Hs, Ws, dst, Hd, Wd, n_images, inv_scale_x,
                                                                      inv_scale_y, vec);
  HIAST_CHECK_LAUNCH();
  return HIAST_OK;
}
