// (2) Hard-aware pseudo-label augmentation: inter-image masked gather (copy-paste) for sm_100a.
//
// Reference: sseg/datasets/preprocessor.py:102-112 (run_original loop body): for every hard class c,
// selected_mask[lbl_ == c] = True and copy_paste_mask[lbl_ == c] = c (14 full-image compares + 28
// masked stores on the CPU), then two fancy-index copies img[M] = img_[M], lbl[M] = lbl_[M].
// Here: one pass, 13 B/px (read dst img 3 + lbl 1 + mask 1*, donor img 3 + lbl 1; write img 3 +
// lbl 1 + mask 1), 16 pixels per thread with 128-bit accesses, the hard-class set as a 256-bit LUT
// in kernel parameters.  Donor choice (host RNG) stays on the host (preprocessor.py:70-77,93-97).
#include <algorithm>

#include "common.cuh"

namespace hiast {

constexpr int kThreadsP2 = 256;
constexpr int kPxP = 16;

struct HardLut {
  uint32_t w[8];
};

__device__ __forceinline__ bool is_hard(const HardLut& lut, unsigned v) {
  const unsigned idx = v >> 5;
  uint32_t w = lut.w[0];  // static indices only: a dynamically indexed kernel parameter would be spilled to local memory
#pragma unroll
  for (int k = 1; k < 8; ++k) w = (idx == k) ? lut.w[k] : w;
  return (w >> (v & 31)) & 1u;
}

// One 16-pixel group: everything after the donor labels is loaded only if some pixel of the group is pasted
// (hard classes are rare classes: on real label maps most groups stop after the 16-byte label load).
struct PasteGroup {
  uint4 dl4;
  unsigned sel;
  uint4 l4, m4, a4[3], b4[3];
};

__device__ __forceinline__ unsigned paste_select(const HardLut& lut, const uint4& dl4) {
  const unsigned dl[4] = {dl4.x, dl4.y, dl4.z, dl4.w};
  unsigned sel = 0;  // bit j = pixel j selected
#pragma unroll
  for (int j = 0; j < kPxP; ++j) sel |= static_cast<unsigned>(is_hard(lut, (dl[j >> 2] >> (8 * (j & 3))) & 0xff)) << j;
  return sel;
}

__device__ __forceinline__ void paste_store(const PasteGroup& g, uint8_t* lbl, uint8_t* cp_mask, uint8_t* img, size_t dst) {
  const unsigned dl[4] = {g.dl4.x, g.dl4.y, g.dl4.z, g.dl4.w};
  const unsigned lw[4] = {g.l4.x, g.l4.y, g.l4.z, g.l4.w};
  const unsigned mw[4] = {g.m4.x, g.m4.y, g.m4.z, g.m4.w};
  unsigned lo[4], mo[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    unsigned bm = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) bm |= ((g.sel >> (4 * k + j)) & 1u) ? (0xffu << (8 * j)) : 0u;
    lo[k] = (lw[k] & ~bm) | (dl[k] & bm);
    mo[k] = (mw[k] & ~bm) | (dl[k] & bm);
  }
  *reinterpret_cast<uint4*>(lbl + dst) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  *reinterpret_cast<uint4*>(cp_mask + dst) = make_uint4(mo[0], mo[1], mo[2], mo[3]);
  // 16 RGB pixels = 48 bytes = 12 words; byte b belongs to pixel b / 3
  uint4* oi = reinterpret_cast<uint4*>(img + dst * 3);
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const unsigned aw[4] = {g.a4[q].x, g.a4[q].y, g.a4[q].z, g.a4[q].w};
    const unsigned bw[4] = {g.b4[q].x, g.b4[q].y, g.b4[q].z, g.b4[q].w};
    unsigned ow[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      unsigned bm = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int byte = 16 * q + 4 * k + j;
        bm |= ((g.sel >> (byte / 3)) & 1u) ? (0xffu << (8 * j)) : 0u;
      }
      ow[k] = (aw[k] & ~bm) | (bw[k] & bm);
    }
    oi[q] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  }
}

constexpr int kGroupsP = 2;   // 16-pixel groups per thread per tile: their loads are all in flight together

// No __restrict__ on the image buffers: donors may live in the same allocation as the destinations
// (different images); a destination image must not be used as a donor in the same call.
__global__ void __launch_bounds__(kThreadsP2) k_copy_paste(uint8_t* img, uint8_t* lbl, uint8_t* cp_mask,
                                                           const uint8_t* d_img, const uint8_t* d_lbl,
                                                           const int32_t* __restrict__ donor_index, int n_images, int64_t HW,
                                                           int tiles_per_image, long long n_tiles, HardLut lut, int vec) {
  for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int i = static_cast<int>(t / tiles_per_image);
    const int tile = static_cast<int>(t - static_cast<long long>(i) * tiles_per_image);
    const int d = donor_index ? donor_index[i] : i;
    if (vec) {
      PasteGroup g[kGroupsP];
      size_t dst[kGroupsP], src[kGroupsP];
      bool in[kGroupsP];
      // stage 1: the donor labels of every group
#pragma unroll
      for (int k = 0; k < kGroupsP; ++k) {
        const int64_t p0 = ((static_cast<int64_t>(tile) * kGroupsP + k) * kThreadsP2 + threadIdx.x) * kPxP;
        in[k] = p0 < HW;
        dst[k] = static_cast<size_t>(i) * HW + p0;
        src[k] = static_cast<size_t>(d) * HW + p0;
        if (in[k]) g[k].dl4 = __ldcs(reinterpret_cast<const uint4*>(d_lbl + src[k]));
      }
      // stage 2: destination + donor pixels of the groups that paste anything, all loads before any use
#pragma unroll
      for (int k = 0; k < kGroupsP; ++k) {
        g[k].sel = in[k] ? paste_select(lut, g[k].dl4) : 0u;
        if (g[k].sel) {
          g[k].l4 = *reinterpret_cast<const uint4*>(lbl + dst[k]);
          g[k].m4 = *reinterpret_cast<const uint4*>(cp_mask + dst[k]);
          const uint4* di = reinterpret_cast<const uint4*>(d_img + src[k] * 3);
          const uint4* oi = reinterpret_cast<const uint4*>(img + dst[k] * 3);
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            g[k].a4[q] = oi[q];
            g[k].b4[q] = __ldcs(di + q);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < kGroupsP; ++k)
        if (g[k].sel) paste_store(g[k], lbl, cp_mask, img, dst[k]);
    } else {
      for (int k = 0; k < kGroupsP; ++k) {
        const int64_t p0 = ((static_cast<int64_t>(tile) * kGroupsP + k) * kThreadsP2 + threadIdx.x) * kPxP;
        if (p0 >= HW) continue;
        const size_t dst = static_cast<size_t>(i) * HW + p0;
        const size_t src = static_cast<size_t>(d) * HW + p0;
        const int n = static_cast<int>(min(static_cast<int64_t>(kPxP), HW - p0));
        for (int j = 0; j < n; ++j) {
          const unsigned v = d_lbl[src + j];
          if (is_hard(lut, v)) {
            lbl[dst + j] = static_cast<uint8_t>(v);
            cp_mask[dst + j] = static_cast<uint8_t>(v);
            for (int ch = 0; ch < 3; ++ch) img[(dst + j) * 3 + ch] = d_img[(src + j) * 3 + ch];
          }
        }
      }
    }
  }
}

}  // namespace hiast

using namespace hiast;

extern "C" int hiast_copy_paste(uint8_t* img, uint8_t* lbl, uint8_t* cp_mask, const uint8_t* donor_img,
                                const uint8_t* donor_lbl, const int32_t* donor_index, int n_images, int64_t HW,
                                const uint32_t* hard_lut_host, void* stream) {
  if (!img || !lbl || !cp_mask || !donor_img || !donor_lbl || !hard_lut_host) return HIAST_ERR_INVALID_ARG;
  if (n_images < 0 || HW < 1) return HIAST_ERR_INVALID_ARG;
  if (n_images == 0) return HIAST_OK;
  HardLut lut;
  for (int k = 0; k < 8; ++k) lut.w[k] = hard_lut_host[k];
  const int px_per_tile = kThreadsP2 * kPxP * kGroupsP;
  const int tiles_per_image = static_cast<int>((HW + px_per_tile - 1) / px_per_tile);
  const long long n_tiles = static_cast<long long>(tiles_per_image) * n_images;
  auto al16 = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
  const int vec = (HW % kPxP == 0) && al16(img) && al16(lbl) && al16(cp_mask) && al16(donor_img) && al16(donor_lbl);
  const int grid = static_cast<int>(std::min<long long>(n_tiles, static_cast<long long>(sm_count()) * 16));
  k_copy_paste<<<grid, kThreadsP2, 0, as_stream(stream)>>>(img, lbl, cp_mask, donor_img, donor_lbl, donor_index, n_images,
                                                           HW, tiles_per_image, n_tiles, lut, vec);
  HIAST_CHECK_LAUNCH();
  return HIAST_OK;
}
