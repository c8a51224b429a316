// (1b) IAS phase A straight from the network's stride-8 logits: bilinear up-sampling (align_corners=True) fused in.
// Reference: sseg/models/segmentors/self_training_segmentor.py:27 followed by workflows/pseudo_label_generator.py:192-201.
#include "ias_common.cuh"

namespace hiast {

// ---- fused bilinear up-sampling ---------------------------------------------
// The reference up-samples the stride-8 network output to full resolution with
// F.interpolate(mode='bilinear', align_corners=True) (self_training_segmentor.py:27) and only then runs the
// softmax: a 159 MB tensor per image is written and read back although it is a pure function of a 2.5 MB one.
// Here phase A reads the LOW-RESOLUTION logits and interpolates on the fly.  Arithmetic is ATen's, operation for
// operation (read off the sm_100 SASS of upsample_bilinear2d_out_frame<float,float>):
//   src = scale * dst (scale = float(in-1)/float(out-1), computed on the host);  i1 = trunc(src);  l1 = src - i1;  l0 = 1 - l1
//   row(r) = fma(w0, v[r][x1], w1 * v[r][x1 + x1p]);   val = fma(h0, row(y1), h1 * row(y1 + y1p))
// so conf / label are bit-identical to softmax(interpolate(x)).max(1) on CUDA.
// A CTA handles 1024 consecutive pixels of one output row: the two source rows x C channels x the needed source
// columns are staged in shared memory once (a few KB, L2-resident input), then every thread interpolates its
// 4 pixels x C channels from shared memory and continues exactly like the full-resolution kernel.
struct UpArgs {
  PhaseAArgs a;
  int h_in, w_in, H, W;
  float rheight, rwidth;
  int max_cols;   // staged source columns per tile
  int max_rows;   // staged source rows per tile
};

constexpr int kRowsU = 4;   // output rows per tile: the staged source window is reused by all of them

template <int C, int MODE>
__global__ void __launch_bounds__(kThreadsA, 2) k_upsample_softmax_hist(UpArgs u) {
  const PhaseAArgs& a = u.a;
  constexpr bool kShared = (MODE == 6);
  extern __shared__ __align__(128) float s_src[];   // [C][max_rows][max_cols]
  __shared__ uint32_t s_top[C];
  __shared__ int s_sched[2];
  HistSink<MODE> sink;
  sink.nb = a.nb;
  sink.nbs = row_stride(a.nb);
  sink.s = s_top;
  sink.g = a.hist;
  sink.top0 = a.nb - 1;
  sink.run_lbl = 0;
  sink.run_cnt = 0;
  if (kShared) {
    for (int i = threadIdx.x; i < C; i += kThreadsA) s_top[i] = 0;
    __syncthreads();
  }
  auto flush_top = [&]() {
    sink.run_flush();
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += kThreadsA) {
      const uint32_t v = s_top[i];
      if (v) {
        atomicAdd(sink.g + static_cast<size_t>(i) * sink.nbs + sink.top0, v);
        s_top[i] = 0;
      }
    }
    __syncthreads();
  };
  const int tiles_per_row = (u.W + kThreadsA * 4 - 1) / (kThreadsA * 4);
  const int row_blocks = (u.H + kRowsU - 1) / kRowsU;
  const int tiles_per_image = tiles_per_row * row_blocks;
  ChunkSched sched;
  sched.init(a.sched, a.n_tiles);
  int cur_group = -1;
  const size_t plane_in = static_cast<size_t>(u.h_in) * u.w_in;
  for (; sched.cur < sched.n_chunks; sched.advance(s_sched)) {
    sched.fetch(s_sched);
    const int t0 = sched.cur * kChunkTiles;
    const int t1 = static_cast<int>(min(static_cast<long long>(t0) + kChunkTiles, a.n_tiles));
    for (int t = t0; t < t1; ++t) {
      const int img = t / tiles_per_image;
      const int rem = t - img * tiles_per_image;
      const int yb = (rem / tiles_per_row) * kRowsU;
      const int y_end = min(yb + kRowsU, u.H);
      const int x0 = (rem % tiles_per_row) * (kThreadsA * 4);
      const int group = img / a.group_size;
      if (group != cur_group) {
        if (kShared && cur_group >= 0) flush_top();
        cur_group = group;
        sink.g = a.hist + static_cast<size_t>(group) * C * sink.nbs;
      }
      // staged source window: rows [rb, rb + nrows), columns [cb, cb + ncols)
      const int rb = static_cast<int>(__fmul_rn(static_cast<float>(yb), u.rheight));
      const int re = min(static_cast<int>(__fmul_rn(static_cast<float>(y_end - 1), u.rheight)) + 1, u.h_in - 1);
      const int nrows = re - rb + 1;
      const int cb = static_cast<int>(__fmul_rn(static_cast<float>(x0), u.rwidth));
      const int x_last = min(x0 + kThreadsA * 4, u.W) - 1;
      const int ce = min(static_cast<int>(__fmul_rn(static_cast<float>(x_last), u.rwidth)) + 1, u.w_in - 1);
      const int ncols = ce - cb + 1;
      __syncthreads();   // previous tile's readers are done
      const float* src = a.logits + static_cast<size_t>(img) * C * plane_in + static_cast<size_t>(rb) * u.w_in + cb;
      for (int cr = threadIdx.x >> 5; cr < C * nrows; cr += kThreadsA / 32) {
        const int c = cr / nrows;
        const int r = cr - c * nrows;
        const float* srow = src + c * plane_in + static_cast<size_t>(r) * u.w_in;
        float* drow = s_src + (c * u.max_rows + r) * u.max_cols;
        for (int col = lane_id(); col < ncols; col += 32) drow[col] = srow[col];
      }
      __syncthreads();
      const int x = x0 + threadIdx.x * 4;
      const bool valid = x < u.W;   // W % 4 == 0: a thread's 4 pixels are all inside or all outside
      // horizontal source positions of the thread's 4 pixels (same for every row of the block)
      float w0l[4], w1l[4];
      int xo[4], x1p[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float w1r = __fmul_rn(static_cast<float>(x + j), u.rwidth);
        const int x1 = static_cast<int>(w1r);
        x1p[j] = (x1 < u.w_in - 1) ? 1 : 0;
        w1l[j] = __fsub_rn(w1r, static_cast<float>(x1));
        w0l[j] = __fsub_rn(1.0f, w1l[j]);
        xo[j] = x1 - cb;
      }
      for (int y = yb; y < y_end; ++y) {
        const float h1r = __fmul_rn(static_cast<float>(y), u.rheight);
        const int y1 = static_cast<int>(h1r);
        const int y1p = (y1 < u.h_in - 1) ? 1 : 0;
        const float h1l = __fsub_rn(h1r, static_cast<float>(y1));
        const float h0l = __fsub_rn(1.0f, h1l);
        const int r0 = (y1 - rb) * u.max_cols;
        const int r1 = (y1 + y1p - rb) * u.max_cols;
        float v[4][C];
        float cf[4];
        int lb[4];
        if (valid) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float* p0 = s_src + xo[j];
#pragma unroll
            for (int c = 0; c < C; ++c) {
              const float* pc = p0 + c * u.max_rows * u.max_cols;
              const float top = __fmaf_rn(w0l[j], pc[r0], __fmul_rn(w1l[j], pc[r0 + x1p[j]]));
              const float bot = __fmaf_rn(w0l[j], pc[r1], __fmul_rn(w1l[j], pc[r1 + x1p[j]]));
              v[j][c] = __fmaf_rn(h0l, top, __fmul_rn(h1l, bot));
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) softmax_argmax<C>(v[j], cf[j], lb[j]);
          const size_t o4 = (static_cast<size_t>(img) * u.H * u.W + static_cast<size_t>(y) * u.W + x) >> 2;
          reinterpret_cast<float4*>(a.conf)[o4] = make_float4(cf[0], cf[1], cf[2], cf[3]);
          reinterpret_cast<uchar4*>(a.label)[o4] = make_uchar4(lb[0], lb[1], lb[2], lb[3]);
        }
        int bins[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          bins[j] = 0;
          if (valid) bins[j] = min(max(static_cast<int>(fp16_key(cf[j])) - a.key_lo, 0), a.nb - 1);
          else lb[j] = 0;
        }
        sink.template add_px<4>(valid, lb, bins);
      }
    }
  }
  if (kShared && cur_group >= 0) flush_top();
}

// ---- fused up-sampling, second version ---------------------------------------------------------------------
// The first kernel is bound by shared-memory loads: 4 scalar LDS per (pixel, channel).  Here a thread owns a COLUMN
// of kRowsU = 4 vertically adjacent output pixels: they share the two source columns and the horizontal weights, so
// the horizontal interpolation is done once per staged source row (<= 3 rows: 6 LDS and 3 fma per channel for four
// pixels) and each pixel only adds the vertical blend, whose row selection is uniform across the CTA.  The four
// pixels then go through the packed softmax / arg-max as two pairs, the histogram lives in the shared-memory table
// of the group-resident kernel (here it covers practically every key: the staging buffers are small), the source
// window of the next tile is fetched with cp.async while the current one is computed, and the tiles are split
// statically over one 512-thread CTA per SM.  Arithmetic identical to the first kernel (= ATen's).
constexpr int kThreadsU2 = 512;
constexpr int kColsPerThreadU2 = 4;
constexpr int kTileColsU2 = kThreadsU2 * kColsPerThreadU2;

struct UpArgs2 {
  UpArgs u;
  int hi0, words;
};

// One output column of kRowsU rows: horizontal blend of the staged source rows (two of them if SPLIT == 4 or 0), then
// the vertical blend with compile-time row selection.  Same operations as ATen's upsample_bilinear2d kernel.
template <int C, int SPLIT>
__device__ __forceinline__ void interp_column(const float* p0, int max_cols, int x1p, float w0l, float w1l,
                                              const float (&h0l)[kRowsU], const float (&h1l)[kRowsU], float (&v)[kRowsU][C]) {
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float* pc = p0 + c * 3 * max_cols;
    float hr0 = 0.f, hr1, hr2 = 0.f;
    if (SPLIT > 0) hr0 = __fmaf_rn(w0l, pc[0], __fmul_rn(w1l, pc[x1p]));
    hr1 = __fmaf_rn(w0l, pc[max_cols], __fmul_rn(w1l, pc[max_cols + x1p]));
    if (SPLIT < kRowsU) hr2 = __fmaf_rn(w0l, pc[2 * max_cols], __fmul_rn(w1l, pc[2 * max_cols + x1p]));
#pragma unroll
    for (int j = 0; j < kRowsU; ++j) {
      const float tp = (j < SPLIT) ? hr0 : hr1;
      const float bt = (j < SPLIT) ? hr1 : hr2;
      v[j][c] = __fmaf_rn(h0l[j], tp, __fmul_rn(h1l[j], bt));
    }
  }
}

template <int C>
__global__ void __launch_bounds__(kThreadsU2, 1) k_upsample_softmax_hist_v2(UpArgs2 ua) {
  const UpArgs& u = ua.u;
  const PhaseAArgs& a = u.a;
  extern __shared__ __align__(128) unsigned char s_raw[];
  const int stage_floats = C * 3 * u.max_cols;                     // one staging buffer: [C][3][max_cols]
  float* s_src = reinterpret_cast<float*>(s_raw);                  // two of them
  uint32_t* s_tab = reinterpret_cast<uint32_t*>(s_raw + sizeof(float) * 2 * stage_floats);   // [C][words]
  __shared__ uint32_t s_top[C];
  const int nbs = row_stride(a.nb);
  const int top = a.nb - 1;
  const int hi0 = ua.hi0, words = ua.words;
  for (int i = threadIdx.x; i < C * words; i += kThreadsU2) s_tab[i] = 0;
  if (threadIdx.x < C) s_top[threadIdx.x] = 0;
  const int tiles_per_row = (u.W + kTileColsU2 - 1) / kTileColsU2;
  const int row_blocks = (u.H + kRowsU - 1) / kRowsU;
  const int tiles_per_image = tiles_per_row * row_blocks;
  const size_t plane_in = static_cast<size_t>(u.h_in) * u.w_in;
  const long long lo = a.n_tiles * blockIdx.x / gridDim.x, hi = a.n_tiles * (blockIdx.x + 1) / gridDim.x;
  struct Win { int img, yb, y_end, x0, rb, nrows, cb, ncols; };
  auto window = [&](long long t) {
    Win w;
    w.img = static_cast<int>(t / tiles_per_image);
    const int rem = static_cast<int>(t - static_cast<long long>(w.img) * tiles_per_image);
    w.yb = (rem / tiles_per_row) * kRowsU;
    w.y_end = min(w.yb + kRowsU, u.H);
    w.x0 = (rem % tiles_per_row) * kTileColsU2;
    w.rb = static_cast<int>(__fmul_rn(static_cast<float>(w.yb), u.rheight));
    const int re = min(static_cast<int>(__fmul_rn(static_cast<float>(w.y_end - 1), u.rheight)) + 1, u.h_in - 1);
    w.nrows = re - w.rb + 1;
    w.cb = static_cast<int>(__fmul_rn(static_cast<float>(w.x0), u.rwidth));
    const int x_last = min(w.x0 + kTileColsU2, u.W) - 1;
    const int ce = min(static_cast<int>(__fmul_rn(static_cast<float>(x_last), u.rwidth)) + 1, u.w_in - 1);
    w.ncols = ce - w.cb + 1;
    return w;
  };
  auto stage = [&](const Win& w, int buf) {
    const float* src = a.logits + static_cast<size_t>(w.img) * C * plane_in + static_cast<size_t>(w.rb) * u.w_in + w.cb;
    float* dst = s_src + buf * stage_floats;
    for (int cr = threadIdx.x >> 5; cr < C * w.nrows; cr += kThreadsU2 / 32) {
      const int c = cr / w.nrows;
      const int r = cr - c * w.nrows;
      const float* srow = src + c * plane_in + static_cast<size_t>(r) * u.w_in;
      const unsigned drow = static_cast<unsigned>(__cvta_generic_to_shared(dst + (c * 3 + r) * u.max_cols));
      for (int col = lane_id(); col < w.ncols; col += 32)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(drow + col * 4), "l"(srow + col) : "memory");
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };
  int cur_group = -1;
  uint32_t* g_hist = a.hist;
  int run_lbl = 0;
  unsigned run_cnt = 0;
  auto flush = [&]() {
    if (run_cnt) atomicAdd(s_top + run_lbl, run_cnt);
    run_cnt = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < C * words; i += kThreadsU2) {
      const uint32_t wv = s_tab[i];
      if (wv) {
        const int c = i / words, k = i - c * words;
        uint32_t* row = g_hist + static_cast<size_t>(c) * nbs + hi0 + 2 * k;
        if (wv & 0xffffu) atomicAdd(row, wv & 0xffffu);
        if (wv >> 16) atomicAdd(row + 1, wv >> 16);
        s_tab[i] = 0;
      }
    }
    if (threadIdx.x < C) {
      const uint32_t wv = s_top[threadIdx.x];
      if (wv) {
        atomicAdd(g_hist + static_cast<size_t>(threadIdx.x) * nbs + top, wv);
        s_top[threadIdx.x] = 0;
      }
    }
    __syncthreads();
  };
  if (lo < hi) stage(window(lo), 0);
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  __syncthreads();
  for (long long t = lo; t < hi; ++t) {
    const Win w = window(t);
    const int buf = static_cast<int>((t - lo) & 1);
    if (t + 1 < hi) stage(window(t + 1), buf ^ 1);   // that buffer's readers finished before the last barrier
    const int group = w.img / a.group_size;
    if (group != cur_group) {
      if (cur_group >= 0) flush();
      cur_group = group;
      g_hist = a.hist + static_cast<size_t>(group) * C * nbs;
    }
    // vertical positions of the block's rows (uniform over the CTA)
    float h0l[kRowsU], h1l[kRowsU];
    bool top1[kRowsU], bot1[kRowsU], bot2[kRowsU];
#pragma unroll
    for (int j = 0; j < kRowsU; ++j) {
      const int y = min(w.yb + j, u.H - 1);
      const float h1r = __fmul_rn(static_cast<float>(y), u.rheight);
      const int y1 = static_cast<int>(h1r);
      const int y1p = (y1 < u.h_in - 1) ? 1 : 0;
      h1l[j] = __fsub_rn(h1r, static_cast<float>(y1));
      h0l[j] = __fsub_rn(1.0f, h1l[j]);
      const int ti = y1 - w.rb, bi = ti + y1p;
      top1[j] = ti == 1;
      bot1[j] = bi == 1;
      bot2[j] = bi == 2;
    }
    int split = 0;   // rows [0, split): (0, 1); rows [split, 4): (1, 2); -1 if the pattern is anything else
#pragma unroll
    for (int j = 0; j < kRowsU; ++j) {
      const bool first = !top1[j] && bot1[j], second = top1[j] && bot2[j];
      if (first && split == j) split = j + 1;
      else if (!(second && split >= 0 && split <= j)) split = -1;
    }
    const float* sb = s_src + buf * stage_floats;
#pragma unroll 1
    for (int k = 0; k < kColsPerThreadU2; ++k) {
      const int x = w.x0 + k * kThreadsU2 + threadIdx.x;
      if (x < u.W) {
        const float w1r = __fmul_rn(static_cast<float>(x), u.rwidth);
        const int x1 = static_cast<int>(w1r);
        const int x1p = (x1 < u.w_in - 1) ? 1 : 0;
        const float w1l = __fsub_rn(w1r, static_cast<float>(x1));
        const float w0l = __fsub_rn(1.0f, w1l);
        const float* p0 = sb + (x1 - w.cb);
        float v[kRowsU][C];
        // the rows' source-row pattern is uniform over the CTA: rows [0, split) blend staged rows (0, 1), the rest rows
        // (1, 2) -- the only patterns an up-sampling by >= 4 produces away from the bottom border; anything else takes
        // the generic selects
        if (split >= 0) {
          switch (split) {
            case 4: interp_column<C, 4>(p0, u.max_cols, x1p, w0l, w1l, h0l, h1l, v); break;
            case 3: interp_column<C, 3>(p0, u.max_cols, x1p, w0l, w1l, h0l, h1l, v); break;
            case 2: interp_column<C, 2>(p0, u.max_cols, x1p, w0l, w1l, h0l, h1l, v); break;
            case 1: interp_column<C, 1>(p0, u.max_cols, x1p, w0l, w1l, h0l, h1l, v); break;
            default: interp_column<C, 0>(p0, u.max_cols, x1p, w0l, w1l, h0l, h1l, v); break;
          }
        } else {
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const float* pc = p0 + c * 3 * u.max_cols;
            const float hr0 = __fmaf_rn(w0l, pc[0], __fmul_rn(w1l, pc[x1p]));
            const float hr1 = __fmaf_rn(w0l, pc[u.max_cols], __fmul_rn(w1l, pc[u.max_cols + x1p]));
            const float hr2 = __fmaf_rn(w0l, pc[2 * u.max_cols], __fmul_rn(w1l, pc[2 * u.max_cols + x1p]));
#pragma unroll
            for (int j = 0; j < kRowsU; ++j) {
              const float tp = top1[j] ? hr1 : hr0;
              const float bt = bot2[j] ? hr2 : (bot1[j] ? hr1 : hr0);
              v[j][c] = __fmaf_rn(h0l[j], tp, __fmul_rn(h1l[j], bt));
            }
          }
        }
        float cf[kRowsU];
        int lb[kRowsU];
        bool tie[kRowsU];
        bool any_tie = false;
#pragma unroll
        for (int j = 0; j < kRowsU; j += 2) {
          softmax_argmax_pair<C>(v[j], v[j + 1], cf[j], cf[j + 1], lb[j], lb[j + 1], tie[j], tie[j + 1]);
          any_tie |= tie[j] | tie[j + 1];
        }
        if (any_tie) {
#pragma unroll
          for (int j = 0; j < kRowsU; ++j)
            if (tie[j]) softmax_argmax<C>(v[j], cf[j], lb[j]);
        }
#pragma unroll
        for (int j = 0; j < kRowsU; ++j) {
          const int y = w.yb + j;
          if (y < w.y_end) {
            const size_t o = (static_cast<size_t>(w.img) * u.H + y) * u.W + x;
            a.conf[o] = cf[j];
            a.label[o] = static_cast<uint8_t>(lb[j]);
            const int bin = min(max(static_cast<int>(fp16_key(cf[j])) - a.key_lo, 0), top);
            const int l = lb[j];
            if (bin == top) {
              if (l != run_lbl) {
                if (run_cnt) atomicAdd(s_top + run_lbl, run_cnt);
                run_cnt = 0;
                run_lbl = l;
              }
              run_cnt += 1;
            } else if (bin >= hi0) {
              const int kk = bin - hi0;
              const unsigned sh = (kk & 1) * 16;
              tab16_add(s_tab + l * words + (kk >> 1), sh, g_hist + static_cast<size_t>(l) * nbs + bin);
            } else {
              atomicAdd(g_hist + static_cast<size_t>(l) * nbs + bin, 1u);
            }
          }
        }
      }
    }
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    __syncthreads();
  }
  if (cur_group >= 0) flush();
}

}  // namespace hiast

using namespace hiast;

namespace hiast {
bool g_upsample_v1 = false;   // development switch: first up-sampling kernel (hiast_debug_upsample_v1)
}

extern "C" int hiast_debug_upsample_v1(int on) {
  hiast::g_upsample_v1 = on != 0;
  return HIAST_OK;
}

extern "C" int hiast_ias_upsample_softmax_hist(const float* logits_lr, int n_images, int C, int h_in, int w_in, int H, int W,
                                               int group_size, int key_lo, int accumulate, float* conf, uint8_t* label,
                                               uint32_t* hist, void* stream) {
  if (!logits_lr || !conf || !label || !hist) return HIAST_ERR_INVALID_ARG;
  if (n_images < 0 || h_in < 1 || w_in < 1 || H < 1 || W < 1 || group_size < 1) return HIAST_ERR_INVALID_ARG;
  if (key_lo < 0 || key_lo > HIAST_KEY_ONE) return HIAST_ERR_INVALID_ARG;
  if (C != 19 && C != 16) return HIAST_ERR_UNSUPPORTED;
  if (W % 4 != 0 || reinterpret_cast<uintptr_t>(conf) % 16 != 0 || reinterpret_cast<uintptr_t>(label) % 4 != 0)
    return HIAST_ERR_UNSUPPORTED;
  if (H < h_in || W < w_in) return HIAST_ERR_UNSUPPORTED;   // up-sampling only
  cudaStream_t st = as_stream(stream);
  const int n_groups = (n_images + group_size - 1) / group_size;
  if (!accumulate) HIAST_CUDA_TRY(cudaMemsetAsync(hist, 0, hiast_ias_hist_bytes(n_groups, C, key_lo), st));
  if (n_images == 0) return HIAST_OK;
  UpArgs u;
  u.a.logits = logits_lr; u.a.conf = conf; u.a.label = label; u.a.hist = hist;
  u.a.n_images = n_images; u.a.C = C; u.a.HW = static_cast<int64_t>(H) * W;
  u.a.group_size = group_size; u.a.key_lo = key_lo; u.a.nb = HIAST_KEY_ONE - key_lo + 1;
  u.h_in = h_in; u.w_in = w_in; u.H = H; u.W = W;
  // ATen: area_pixel_compute_scale<float>(in, out, align_corners=true) = float(in - 1) / (out - 1), 0 when out == 1
  u.rheight = H > 1 ? static_cast<float>(h_in - 1) / static_cast<float>(H - 1) : 0.f;
  u.rwidth = W > 1 ? static_cast<float>(w_in - 1) / static_cast<float>(W - 1) : 0.f;
  {
    // second version: one column of 4 rows per thread; needs <= 3 staged source rows per block of 4 output rows
    const int rows_needed = std::min(h_in, static_cast<int>(static_cast<double>(kRowsU) * u.rheight) + 3);
    if (rows_needed <= 3 && !g_upsample_v1) {
      UpArgs2 ua;
      ua.u = u;
      UpArgs& v = ua.u;
      v.max_rows = 3;
      v.max_cols = std::min(w_in, static_cast<int>(static_cast<double>(kTileColsU2) * u.rwidth) + 4);
      const int tpr = (W + kTileColsU2 - 1) / kTileColsU2;
      v.a.tiles_per_image = tpr * ((H + kRowsU - 1) / kRowsU);
      v.a.n_tiles = static_cast<long long>(v.a.tiles_per_image) * n_images;
      const size_t stage = sizeof(float) * 2 * C * 3 * v.max_cols;
      constexpr size_t kBudget = 227 * 1024 - 2048;
      if (v.a.n_tiles < (1ll << 31) && stage + 16 * 1024 < kBudget) {
        const int top = v.a.nb - 1;
        int words = static_cast<int>((kBudget - stage) / (sizeof(uint32_t) * C));
        words = std::min(words, (top + 1) / 2);
        ua.words = words;
        ua.hi0 = std::max(top - 2 * words, 0);
        const size_t smem = stage + sizeof(uint32_t) * C * words;
        const int grid = static_cast<int>(std::min<long long>(sm_count(), v.a.n_tiles));
        if (C == 19) {
          HIAST_TRY(ensure_dyn_smem(k_upsample_softmax_hist_v2<19>, kBudget));
          k_upsample_softmax_hist_v2<19><<<grid, kThreadsU2, smem, st>>>(ua);
        } else {
          HIAST_TRY(ensure_dyn_smem(k_upsample_softmax_hist_v2<16>, kBudget));
          k_upsample_softmax_hist_v2<16><<<grid, kThreadsU2, smem, st>>>(ua);
        }
        HIAST_CHECK_LAUNCH();
        return HIAST_OK;
      }
    }
  }
  const int tile_px = kThreadsA * 4;
  u.max_cols = std::min(w_in, static_cast<int>(static_cast<double>(tile_px) * u.rwidth) + 4);
  u.max_rows = std::min(h_in, static_cast<int>(static_cast<double>(kRowsU) * u.rheight) + 3);
  const int tiles_per_row = (W + tile_px - 1) / tile_px;
  u.a.tiles_per_image = tiles_per_row * ((H + kRowsU - 1) / kRowsU);
  u.a.n_tiles = static_cast<long long>(u.a.tiles_per_image) * n_images;
  if (u.a.n_tiles >= (1ll << 31)) return HIAST_ERR_UNSUPPORTED;
  const size_t smem = sizeof(float) * C * u.max_rows * u.max_cols;
  if (smem > 96 * 1024) return HIAST_ERR_UNSUPPORTED;
  int rc = next_sched_slot(&u.a.sched, st);
  if (rc != HIAST_OK) return rc;
  const long long n_chunks = (u.a.n_tiles + kChunkTiles - 1) / kChunkTiles;
  if (C == 19) {
    HIAST_TRY(ensure_dyn_smem(k_upsample_softmax_hist<19, 6>, smem));
    int grid = resident_grid(k_upsample_softmax_hist<19, 6>, kThreadsA, smem);
    if (grid > n_chunks) grid = static_cast<int>(n_chunks);
    k_upsample_softmax_hist<19, 6><<<grid, kThreadsA, smem, st>>>(u);
  } else {
    HIAST_TRY(ensure_dyn_smem(k_upsample_softmax_hist<16, 6>, smem));
    int grid = resident_grid(k_upsample_softmax_hist<16, 6>, kThreadsA, smem);
    if (grid > n_chunks) grid = static_cast<int>(n_chunks);
    k_upsample_softmax_hist<16, 6><<<grid, kThreadsA, smem, st>>>(u);
  }
  HIAST_CHECK_LAUNCH();
  return HIAST_OK;
}

