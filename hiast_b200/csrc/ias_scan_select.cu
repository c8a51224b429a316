// (1c) IAS phases B and C: threshold chain (workflows/pseudo_label_generator.py:171-179,207-209), threshold-and-mask
// pass (:71-89), mean-confidence EMA (:95-105), and the histogram from caller-provided conf / label.
#include <string.h>

#include "ias_common.cuh"

namespace hiast {

// a2 alone: histogram from caller-provided conf / label.
template <typename L>
__global__ void __launch_bounds__(256) k_conf_hist(const float* __restrict__ conf, const L* __restrict__ label,
                                                   long long total, int64_t HW, int C, int group_size, int key_lo,
                                                   int nb, uint8_t* __restrict__ label_out, uint32_t* __restrict__ hist) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long lraw = static_cast<long long>(label[i]);
    if (label_out) label_out[i] = static_cast<uint8_t>(lraw);
    if (lraw < 0 || lraw >= C) continue;
    const int img = static_cast<int>(i / HW);
    int bin = static_cast<int>(fp16_key(conf[i])) - key_lo;
    bin = min(max(bin, 0), nb - 1);
    atomicAdd(hist + (static_cast<size_t>(img / group_size) * C + lraw) * row_stride(nb) + bin, 1u);
  }
}

// ------------------------------------------------------------------------------------------
// phase B
// ------------------------------------------------------------------------------------------

// In-place inclusive prefix sum of every histogram row (one CTA per row).
constexpr int kThreadsP = 256;
__global__ void __launch_bounds__(kThreadsP) k_hist_prefix(uint32_t* __restrict__ hist, int nb) {
  __shared__ uint32_t s_warp[kThreadsP / 32];
  __shared__ uint32_t s_carry;
  uint32_t* row = hist + static_cast<size_t>(blockIdx.x) * row_stride(nb);
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += kThreadsP * 4) {
    const int i0 = base + threadIdx.x * 4;
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = (i0 + k < nb) ? row[i0 + k] : 0u;
    v[1] += v[0]; v[2] += v[1]; v[3] += v[2];
    uint32_t x = v[3];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane_id() >= o) x += y;
    }
    if (lane_id() == 31) s_warp[threadIdx.x >> 5] = x;
    __syncthreads();
    uint32_t off = s_carry;
    for (int w = 0; w < (threadIdx.x >> 5); ++w) off += s_warp[w];
    off += x - v[3];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (i0 + k < nb) row[i0 + k] = v[k] + off;
    __syncthreads();
    if (threadIdx.x == kThreadsP - 1) s_carry = off + v[3];
    __syncthreads();
  }
}

// One CTA per class; rows of prefix sums are staged in shared memory with 16-byte cp.async, double
// buffered, so the serial chain touches only shared memory.  Warp 0 runs the step (all lanes compute the
// same scalars; the two order-statistic searches are warp-cooperative), the other warps only stage.
constexpr int kThreadsS = 128;
// Token ring over peer memory (multi-GPU).  The only cross-GPU dependency of the path is the threshold state f64[C]
// (pseudo_label_generator.py:207-209 carries it from batch to batch).  Instead of an NCCL receive kernel in front of the scan
// and a send kernel behind it, the scan kernel itself takes the hand-off: CTA c (class c) waits until slot c of THIS GPU's
// mailbox carries sequence number >= in_seq, starts from the value found there, and at the end stores its final threshold
// into slot c of the NEXT GPU's mailbox (a peer pointer: the store travels over NVLink) and releases that slot's sequence
// number.  19 independent flags, no cross-CTA synchronisation, no host involvement per hop.
struct RingToken {
  double value[HIAST_MAX_CLASSES + 1];
  unsigned long long seq[HIAST_MAX_CLASSES + 1];
};

__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
  return t;
}
constexpr unsigned long long kRingTimeoutNs = 10ull * 1000ull * 1000ull * 1000ull;   // then error bit 8, never a hang

__global__ void __launch_bounds__(kThreadsS) k_threshold_scan(const uint32_t* __restrict__ prefix, int n_groups, int C,
                                                              int key_lo, int nb, double alpha, double beta, double gamma,
                                                              double* __restrict__ thr_state, double* __restrict__ thr_groups,
                                                              float* __restrict__ temp_groups, int* __restrict__ error_flag,
                                                              const RingToken* token_in, unsigned long long in_seq,
                                                              RingToken* token_out, unsigned long long out_seq) {
  extern __shared__ __align__(128) uint32_t s_rows[];  // [2][nbs]
  const int c = blockIdx.x;
  const int nbs = row_stride(nb);
  auto stage = [&](int g, int buf) {
    const uint4* src = reinterpret_cast<const uint4*>(prefix + (static_cast<size_t>(g) * C + c) * nbs);
    uint4* dst = reinterpret_cast<uint4*>(s_rows + buf * nbs);
    for (int i = threadIdx.x; i < nbs / 4; i += kThreadsS) {
      const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(dst + i));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src + i));
    }
    asm volatile("cp.async.commit_group;\n" ::);
  };
  double thr = thr_state[c];
  int err = 0;
  if (n_groups > 0) stage(0, 0);               // the first row is on its way while the token is awaited
  if (token_in) {
    __shared__ double s_token;
    if (threadIdx.x == 0) {
      const unsigned long long t0 = global_timer_ns();
      while (ld_acquire_sys_u64(&token_in->seq[c]) < in_seq) {
        __nanosleep(100);
        if (global_timer_ns() - t0 > kRingTimeoutNs) {
          err |= 8;
          break;
        }
      }
      s_token = *reinterpret_cast<const volatile double*>(&token_in->value[c]);
    }
    __syncthreads();
    thr = s_token;
  }
  for (int g = 0; g < n_groups; ++g) {
    if (g + 1 < n_groups) {
      stage(g + 1, (g + 1) & 1);
      asm volatile("cp.async.wait_group 1;\n" ::);
    } else {
      asm volatile("cp.async.wait_group 0;\n" ::);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      float temp;
      const uint32_t* row = s_rows + (g & 1) * nbs;
      const WarpSearch search = {row, nb};
      thr = ias_threshold_step(row, nb, key_lo, thr, alpha, beta, gamma, &temp, &err, search);
      if (threadIdx.x == 0) {
        thr_groups[static_cast<size_t>(g) * C + c] = thr;
        if (temp_groups) temp_groups[static_cast<size_t>(g) * C + c] = temp;
      }
    }
    __syncthreads();  // buffer (g&1) is refilled by the stage() of iteration g+1
  }
  if (threadIdx.x == 0) {
    thr_state[c] = thr;
    if (token_out) {
      *reinterpret_cast<volatile double*>(&token_out->value[c]) = thr;
      __threadfence_system();
      st_release_sys_u64(&token_out->seq[c], out_seq);
    }
    if (err && error_flag) atomicOr(error_flag, err);
  }
}

// ------------------------------------------------------------------------------------------
// phase C
// ------------------------------------------------------------------------------------------
constexpr int kThreadsC = 256;
constexpr int kPxC = 16;  // pixels per 128-bit label load
constexpr int kSubC = 2;  // k_select_private: 16-pixel sub-chunks per thread per tile

struct RunAcc {
  int cur;
  unsigned cnt;
  unsigned long long sum;
};

__device__ __forceinline__ void run_flush(const RunAcc& r, unsigned* s_cnt, unsigned long long* s_sum) {
  if (r.cur != HIAST_IGNORE_LABEL && r.cnt) {
    atomicAdd(s_cnt + r.cur, r.cnt);
    atomicAdd(s_sum + r.cur, r.sum);
  }
}

__device__ __forceinline__ void run_push(RunAcc& r, int pl, float cf, unsigned* s_cnt, unsigned long long* s_sum) {
  if (pl != r.cur) {
    run_flush(r, s_cnt, s_sum);
    r.cur = pl;
    r.cnt = 0;
    r.sum = 0;
  }
  r.cnt += 1;
  r.sum += static_cast<unsigned long long>(cf * 4294967296.0f);  // exact for conf >= 2^-9
}

__global__ void __launch_bounds__(kThreadsC) k_select(const float* __restrict__ conf, const uint8_t* __restrict__ label,
                                                      const double* __restrict__ thr_groups, int n_images, int64_t HW,
                                                      int C, int group_size, int tiles_per_image, long long n_tiles,
                                                      uint8_t* __restrict__ plbl, long long* __restrict__ counts,
                                                      unsigned long long* __restrict__ confsum) {
  __shared__ float s_thr[256];
  __shared__ unsigned s_cnt[256];
  __shared__ unsigned long long s_sum[256];
  const bool vec = (HW % kPxC) == 0;
  const long long t0 = n_tiles * blockIdx.x / gridDim.x;
  const long long t1 = n_tiles * (blockIdx.x + 1) / gridDim.x;
  int cur_img = -1;
  auto flush_image = [&]() {
    __syncthreads();
    if (cur_img >= 0 && threadIdx.x < C) {
      const unsigned n = s_cnt[threadIdx.x];
      if (n) {
        atomicAdd(reinterpret_cast<unsigned long long*>(counts) + static_cast<size_t>(cur_img) * C + threadIdx.x,
                  static_cast<unsigned long long>(n));
        atomicAdd(confsum + static_cast<size_t>(cur_img / group_size) * C + threadIdx.x, s_sum[threadIdx.x]);
      }
    }
    __syncthreads();
  };
  for (long long t = t0; t < t1; ++t) {
    const int img = static_cast<int>(t / tiles_per_image);
    const int tile = static_cast<int>(t - static_cast<long long>(img) * tiles_per_image);
    if (img != cur_img) {
      flush_image();
      cur_img = img;
      // float compare threshold: conf < thr (in double)  <=>  conf < smallest float >= thr
      if (threadIdx.x < 256) {
        s_thr[threadIdx.x] = threadIdx.x < C
                                 ? __double2float_ru(thr_groups[static_cast<size_t>(img / group_size) * C + threadIdx.x])
                                 : INFINITY;
        s_cnt[threadIdx.x] = 0;
        s_sum[threadIdx.x] = 0;
      }
      __syncthreads();
    }
    const int64_t p0 = (static_cast<int64_t>(tile) * kThreadsC + threadIdx.x) * kPxC;
    if (p0 >= HW) continue;
    const size_t base = static_cast<size_t>(img) * HW + p0;
    RunAcc r = {HIAST_IGNORE_LABEL, 0u, 0ull};
    if (vec) {
      const uint4 lraw = *reinterpret_cast<const uint4*>(label + base);
      const unsigned lw[4] = {lraw.x, lraw.y, lraw.z, lraw.w};
      unsigned ow[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 cq = *reinterpret_cast<const float4*>(conf + base + 4 * k);
        const float cf[4] = {cq.x, cq.y, cq.z, cq.w};
        unsigned o = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int l = (lw[k] >> (8 * j)) & 0xff;
          const int pl = (cf[j] < s_thr[l]) ? HIAST_IGNORE_LABEL : l;
          o |= static_cast<unsigned>(pl) << (8 * j);
          run_push(r, pl, cf[j], s_cnt, s_sum);
        }
        ow[k] = o;
      }
      *reinterpret_cast<uint4*>(plbl + base) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    } else {
      const int n = static_cast<int>(min(static_cast<int64_t>(kPxC), HW - p0));
      for (int j = 0; j < n; ++j) {
        const int l = label[base + j];
        const float cf = conf[base + j];
        const int pl = (cf < s_thr[l]) ? HIAST_IGNORE_LABEL : l;
        plbl[base + j] = static_cast<uint8_t>(pl);
        run_push(r, pl, cf, s_cnt, s_sum);
      }
    }
    run_flush(r, s_cnt, s_sum);
  }
  flush_image();
}

// Same pass with per-thread PRIVATE shared-memory accumulators (no atomics on the per-pixel path): thread t owns
// column t of s_acc[C][256], one 64-bit word per class packing the kept-pixel count (bits 48..63) and the sum of
// the kept confidences in units of 2^-31 (bits 0..47; exact for conf >= 2^-8, i.e. any softmax maximum over
// <= 255 classes).  A kept pixel costs one 64-bit shared read-modify-write, an ignored pixel nothing -- the first
// version run-length encoded every pixel (ignored ones included) into separate u32 / u64 columns and spent 47
// instructions per pixel at 2.5-3 TB/s.  Used when C * 256 * 8 bytes fit in shared memory (C <= 32).
constexpr int kFlushTilesC = 1536;   // 32 px per thread per tile: the 16-bit count cannot wrap before a flush
__global__ void __launch_bounds__(kThreadsC, 4) k_select_private(const float* __restrict__ conf, const uint8_t* __restrict__ label,
                                                              const double* __restrict__ thr_groups, int n_images, int64_t HW,
                                                              int C, int group_size, int tiles_per_image, int n_tiles,
                                                              uint8_t* __restrict__ plbl, long long* __restrict__ counts,
                                                              unsigned long long* __restrict__ confsum) {
  extern __shared__ __align__(128) unsigned char s_raw[];
  unsigned long long* s_acc = reinterpret_cast<unsigned long long*>(s_raw);          // [C][256]
  __shared__ float s_thr[256];
  for (int i = threadIdx.x; i < C * kThreadsC; i += kThreadsC) s_acc[i] = 0;
  const int t0 = static_cast<int>(static_cast<long long>(n_tiles) * blockIdx.x / gridDim.x);
  const int t1 = static_cast<int>(static_cast<long long>(n_tiles) * (blockIdx.x + 1) / gridDim.x);
  int img = t0 / tiles_per_image;
  int tile = t0 - img * tiles_per_image;
  int cur_img = -1;
  int since_flush = 0;
  auto flush_image = [&]() {
    __syncthreads();
    if (cur_img >= 0) {
      for (int c = threadIdx.x >> 5; c < C; c += kThreadsC / 32) {
        long long n = 0;
        unsigned long long sm = 0;
#pragma unroll
        for (int k = 0; k < kThreadsC / 32; ++k) {
          const int idx = c * kThreadsC + k * 32 + lane_id();
          const unsigned long long w = s_acc[idx];
          n += static_cast<long long>(w >> 48);
          sm += w & 0xffffffffffffull;
          s_acc[idx] = 0;
        }
        n = warp_sum(n);
        sm = static_cast<unsigned long long>(warp_sum(static_cast<long long>(sm)));
        if (lane_id() == 0 && n) {
          atomicAdd(reinterpret_cast<unsigned long long*>(counts) + static_cast<size_t>(cur_img) * C + c,
                    static_cast<unsigned long long>(n));
          atomicAdd(confsum + static_cast<size_t>(cur_img / group_size) * C + c, sm << 1);   // 2^-32 units
        }
      }
    }
    since_flush = 0;
    __syncthreads();
  };
  unsigned long long* my_acc = s_acc + threadIdx.x;
  for (int t = t0; t < t1; ++t) {
    if (img != cur_img || since_flush >= kFlushTilesC) {
      flush_image();
      if (img != cur_img) {
        cur_img = img;
        s_thr[threadIdx.x] = threadIdx.x < C
                                 ? __double2float_ru(thr_groups[static_cast<size_t>(img / group_size) * C + threadIdx.x])
                                 : INFINITY;
        __syncthreads();
      }
    }
    ++since_flush;
    // A tile is kThreadsC * kPxC * kSubC pixels.  Within it every warp-level access is fully coalesced: the
    // thread's pixels are kQuadsC quads of 4 consecutive pixels, quad q at  tile_px0 + (q * kThreadsC + tid) * 4.
    // All loads are issued before any is consumed.
    constexpr int kQuadsC = kPxC * kSubC / 4;
    const int64_t tile_px0 = static_cast<int64_t>(tile) * (kThreadsC * kPxC * kSubC);
    float4 cq[kQuadsC];
    unsigned lq[kQuadsC];
    bool ok[kQuadsC];
#pragma unroll
    for (int q = 0; q < kQuadsC; ++q) {
      const int64_t px = tile_px0 + (static_cast<int64_t>(q) * kThreadsC + threadIdx.x) * 4;
      ok[q] = px < HW;
      if (ok[q]) {
        const size_t base = static_cast<size_t>(img) * HW + px;
        cq[q] = __ldcs(reinterpret_cast<const float4*>(conf + base));
        lq[q] = __ldcs(reinterpret_cast<const unsigned*>(label + base));
      }
    }
#pragma unroll
    for (int q = 0; q < kQuadsC; ++q) {
      if (ok[q]) {
        const int64_t px = tile_px0 + (static_cast<int64_t>(q) * kThreadsC + threadIdx.x) * 4;
        const float cf[4] = {cq[q].x, cq[q].y, cq[q].z, cq[q].w};
        unsigned o = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int l = (lq[q] >> (8 * j)) & 0xff;
          const bool ign = cf[j] < s_thr[l];
          o |= static_cast<unsigned>(ign ? HIAST_IGNORE_LABEL : l) << (8 * j);
          if (!ign) {
            const unsigned v = __float2uint_rz(cf[j] * 2147483648.0f);
            my_acc[l * kThreadsC] += static_cast<unsigned long long>(v) + (1ull << 48);
          }
        }
        __stcs(reinterpret_cast<unsigned*>(plbl + static_cast<size_t>(img) * HW + px), o);
      }
    }
    if (++tile == tiles_per_image) {
      tile = 0;
      ++img;
    }
  }
  flush_image();
}

// class_mean_probs EMA (pseudo_label_generator.py:95-105).  One warp per class: the 32 lanes fetch and
// reduce the (sum, count) of 32 groups in parallel (the means are independent), then the recurrence runs
// over warp shuffles -- only the handful of dependent f64 operations per group stay serial.
constexpr int kWarpsM = 8;
__global__ void __launch_bounds__(kWarpsM * 32) k_meanprob_scan(const unsigned long long* __restrict__ confsum,
                                                                  const long long* __restrict__ counts, int n_images,
                                                                  int group_size, int n_groups, int C, double cp_gamma,
                                                                  double* __restrict__ mean_state) {
  const int c = blockIdx.x * kWarpsM + (threadIdx.x >> 5);
  if (c >= C) return;
  const int lane = lane_id();
  double cmp = mean_state[c];
  const float omg = static_cast<float>(1.0 - cp_gamma);  // python float weak-cast to f32
  for (int g0 = 0; g0 < n_groups; g0 += 32) {
    const int g = g0 + lane;
    float m = 0.f;
    int have = 0;
    if (g < n_groups) {
      long long n = 0;
      const int i1 = min(n_images, (g + 1) * group_size);
      for (int i = g * group_size; i < i1; ++i) n += counts[static_cast<size_t>(i) * C + c];
      if (n > 0) {  // np.mean of an empty gather is nan -> skipped (:100)
        const double mean64 = static_cast<double>(confsum[static_cast<size_t>(g) * C + c]) * 2.3283064365386963e-10 /
                              static_cast<double>(n);
        m = static_cast<float>(mean64);
        have = 1;
      }
    }
    const unsigned mask = __ballot_sync(0xffffffffu, have);
    const int cnt = min(32, n_groups - g0);
    for (int k = 0; k < cnt; ++k) {
      const float mk = __shfl_sync(0xffffffffu, m, k);
      if ((mask >> k) & 1u) {
        if (cmp == 0.0) cmp = static_cast<double>(mk);
        else cmp = __dadd_rn(__dmul_rn(cmp, cp_gamma), static_cast<double>(__fmul_rn(mk, omg)));
      }
    }
  }
  if (lane == 0) mean_state[c] = cmp;
}

}  // namespace hiast

using namespace hiast;

extern "C" int hiast_ias_conf_hist(const float* conf, const void* label, int label_bytes, int n_images, int64_t HW,
                                   int C, int group_size, int key_lo, int accumulate, uint8_t* label_u8_out,
                                   uint32_t* hist, void* stream) {
  if (!conf || !label || !hist) return HIAST_ERR_INVALID_ARG;
  if (label_bytes != 1 && label_bytes != 8) return HIAST_ERR_INVALID_ARG;
  if (n_images < 0 || HW < 1 || C < 1 || C > HIAST_MAX_CLASSES || group_size < 1) return HIAST_ERR_INVALID_ARG;
  if (key_lo < 0 || key_lo > HIAST_KEY_ONE) return HIAST_ERR_INVALID_ARG;
  cudaStream_t st = as_stream(stream);
  const int n_groups = (n_images + group_size - 1) / group_size;
  if (!accumulate) HIAST_CUDA_TRY(cudaMemsetAsync(hist, 0, hiast_ias_hist_bytes(n_groups, C, key_lo), st));
  if (n_images == 0) return HIAST_OK;
  const long long total = static_cast<long long>(n_images) * HW;
  const int nb = HIAST_KEY_ONE - key_lo + 1;
  const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, static_cast<long long>(sm_count()) * 8));
  if (label_bytes == 1)
    k_conf_hist<uint8_t><<<grid, 256, 0, st>>>(conf, static_cast<const uint8_t*>(label), total, HW, C, group_size,
                                               key_lo, nb, label_u8_out, hist);
  else
    k_conf_hist<long long><<<grid, 256, 0, st>>>(conf, static_cast<const long long*>(label), total, HW, C, group_size,
                                                 key_lo, nb, label_u8_out, hist);
  HIAST_CHECK_LAUNCH();
  return HIAST_OK;
}

extern "C" int hiast_ias_threshold_scan_ring(uint32_t* hist, int n_groups, int C, int key_lo, double alpha, double beta,
                                             double gamma, double* thr_state, double* thr_groups, float* temp_groups,
                                             int* error_flag, const void* token_in, uint64_t in_seq, void* token_out,
                                             uint64_t out_seq, void* stream) {
  if (!hist || !thr_state || !thr_groups) return HIAST_ERR_INVALID_ARG;
  if (n_groups < 0 || C < 1 || C > HIAST_MAX_CLASSES || key_lo < 0 || key_lo > HIAST_KEY_ONE) return HIAST_ERR_INVALID_ARG;
  if (n_groups == 0 && !token_in && !token_out) return HIAST_OK;
  cudaStream_t st = as_stream(stream);
  const int nb = HIAST_KEY_ONE - key_lo + 1;
  if (n_groups > 0) {
    k_hist_prefix<<<n_groups * C, kThreadsP, 0, st>>>(hist, nb);
    HIAST_CHECK_LAUNCH();
  }
  const size_t smem = 2 * static_cast<size_t>(row_stride(nb)) * sizeof(uint32_t);
  HIAST_TRY(ensure_dyn_smem(k_threshold_scan, smem));
  k_threshold_scan<<<C, kThreadsS, smem, st>>>(hist, n_groups, C, key_lo, nb, alpha, beta, gamma, thr_state, thr_groups,
                                               temp_groups, error_flag, static_cast<const RingToken*>(token_in), in_seq,
                                               static_cast<RingToken*>(token_out), out_seq);
  HIAST_CHECK_LAUNCH();
  return HIAST_OK;
}

extern "C" int hiast_ias_threshold_scan(uint32_t* hist, int n_groups, int C, int key_lo, double alpha, double beta,
                                        double gamma, double* thr_state, double* thr_groups, float* temp_groups,
                                        int* error_flag, void* stream) {
  return hiast_ias_threshold_scan_ring(hist, n_groups, C, key_lo, alpha, beta, gamma, thr_state, thr_groups, temp_groups,
                                       error_flag, nullptr, 0, nullptr, 0, stream);
}

// Mailbox of the token ring: 4 KB of device memory that other processes on the node can map (CUDA IPC).
extern "C" size_t hiast_ring_mailbox_bytes(void) { return sizeof(RingToken); }

extern "C" int hiast_ring_create(void** local_box_out, void* ipc_handle_out) {
  if (!local_box_out || !ipc_handle_out) return HIAST_ERR_INVALID_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the handle travels as 64 opaque bytes");
  void* p = nullptr;
  HIAST_CUDA_TRY(cudaMalloc(&p, sizeof(RingToken)));
  cudaError_t e = cudaMemset(p, 0, sizeof(RingToken));
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return cuda_fail(e);
  }
  memcpy(ipc_handle_out, &h, sizeof(h));
  *local_box_out = p;
  return HIAST_OK;
}

extern "C" int hiast_ring_open(const void* ipc_handle, void** peer_box_out) {
  if (!ipc_handle || !peer_box_out) return HIAST_ERR_INVALID_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle, sizeof(h));
  void* p = nullptr;
  HIAST_CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *peer_box_out = p;
  return HIAST_OK;
}

extern "C" int hiast_ring_close(void* peer_box) {
  if (peer_box) HIAST_CUDA_TRY(cudaIpcCloseMemHandle(peer_box));
  return HIAST_OK;
}

extern "C" int hiast_ring_destroy(void* local_box) {
  if (local_box) HIAST_CUDA_TRY(cudaFree(local_box));
  return HIAST_OK;
}

extern "C" int hiast_ias_select(const float* conf, const uint8_t* label, const double* thr_groups, int n_images,
                                int64_t HW, int C, int group_size, uint8_t* plbl, int64_t* counts, uint64_t* confsum,
                                void* stream) {
  if (!conf || !label || !thr_groups || !plbl || !counts || !confsum) return HIAST_ERR_INVALID_ARG;
  if (n_images < 0 || HW < 1 || C < 1 || C > HIAST_MAX_CLASSES || group_size < 1) return HIAST_ERR_INVALID_ARG;
  if (n_images == 0) return HIAST_OK;
  cudaStream_t st = as_stream(stream);
  const int px_per_tile = kThreadsC * kPxC;
  const int tiles_per_image = static_cast<int>((HW + px_per_tile - 1) / px_per_tile);
  const long long n_tiles = static_cast<long long>(tiles_per_image) * n_images;
  const bool aligned = (HW % 4 == 0) && (reinterpret_cast<uintptr_t>(conf) % 16 == 0) &&
                       (reinterpret_cast<uintptr_t>(label) % 4 == 0) && (reinterpret_cast<uintptr_t>(plbl) % 4 == 0);
  if (aligned && C <= 32 && n_tiles < (1ll << 31)) {
    const int tiles_pi = static_cast<int>((HW + px_per_tile * kSubC - 1) / (px_per_tile * kSubC));
    const long long ntl = static_cast<long long>(tiles_pi) * n_images;
    const size_t smem = static_cast<size_t>(C) * kThreadsC * sizeof(unsigned long long);
    HIAST_TRY(ensure_dyn_smem(k_select_private, smem));
    // contiguous tile ranges (image-level flushes stay rare), 4x more CTAs than fit at once so that the hardware
    // scheduler evens out the tail (dynamic chunking was measured slower here: every chunk pays an image flush)
    int grid = resident_grid(k_select_private, kThreadsC, smem) * 4;
    if (grid > ntl) grid = static_cast<int>(ntl);
    k_select_private<<<grid, kThreadsC, smem, st>>>(conf, label, thr_groups, n_images, HW, C, group_size, tiles_pi,
                                                   static_cast<int>(ntl), plbl, reinterpret_cast<long long*>(counts),
                                                   reinterpret_cast<unsigned long long*>(confsum));
    HIAST_CHECK_LAUNCH();
    return HIAST_OK;
  }
  int grid = resident_grid(k_select, kThreadsC, 0);
  if (grid > n_tiles) grid = static_cast<int>(n_tiles);
  k_select<<<grid, kThreadsC, 0, st>>>(conf, label, thr_groups, n_images, HW, C, group_size, tiles_per_image, n_tiles,
                                       plbl, reinterpret_cast<long long*>(counts),
                                       reinterpret_cast<unsigned long long*>(confsum));
  HIAST_CHECK_LAUNCH();
  return HIAST_OK;
}

extern "C" int hiast_ias_meanprob_scan(const uint64_t* confsum, const int64_t* counts, int n_images, int group_size,
                                       int n_groups, int C, double cp_gamma, double* mean_state, void* stream) {
  if (!confsum || !counts || !mean_state) return HIAST_ERR_INVALID_ARG;
  if (n_images < 0 || group_size < 1 || n_groups < 0 || C < 1 || C > HIAST_MAX_CLASSES) return HIAST_ERR_INVALID_ARG;
  if (n_groups == 0) return HIAST_OK;
  k_meanprob_scan<<<(C + kWarpsM - 1) / kWarpsM, kWarpsM * 32, 0, as_stream(stream)>>>(reinterpret_cast<const unsigned long long*>(confsum),
                                                               reinterpret_cast<const long long*>(counts), n_images,
                                                               group_size, n_groups, C, cp_gamma, mean_state);
  HIAST_CHECK_LAUNCH();
  return HIAST_OK;
}
