// (1) Instance-adaptive selector (IAS), phase A: softmax + arg-max + confidence + per-(group, class) key histogram.
// (Phases B / C: ias_scan_select.cu; fused up-sampling: ias_upsample.cu; single-kernel window: ias_fused.cu.)
//
// Reference path: workflows/pseudo_label_generator.py:181-213 (IASPseudoGenerator.run),
// :171-179 (get_ias_threshold), :67-106 (select_and_save_confident_label).
//
//   phase A  k_softmax_hist   logits -> conf f32, label u8, per-(group,class) fp16-key histogram
//   phase B  k_hist_prefix    histogram rows -> inclusive prefix sums (parallel over rows)
//            k_threshold_scan one CTA per class, sequential over groups (the only serial chain)
//   phase C  k_select         conf,label,thr -> plbl u8, per-image counts, per-group conf sums
//            k_meanprob_scan  class_mean_probs EMA over groups
//
// Everything here is HBM-bound streaming work; no tensor cores.  Phase A moves 4*C+5 B/px and is
// the roofline kernel (algorithmic bytes 4*C+1 B/px = 77 B/px for C = 19).
#include "ias_common.cuh"

namespace hiast {

// Self test: packed exponential vs expf() over every non-positive float (pairs (v, v - 1 ulp) so both lanes work).
__global__ void k_selftest_packed_expf(unsigned long long* mismatches) {
  // bit patterns 0x80000000 (-0) .. 0xFF800000 (-inf): 0x7F800001 values, plus +0
  const unsigned long long n = 0x7F800001ull;
  unsigned long long bad = 0;
  for (unsigned long long i = (static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 2; i < n + 1;
       i += static_cast<unsigned long long>(gridDim.x) * blockDim.x * 2) {
    const float a = (i < n) ? __uint_as_float(0x80000000u + static_cast<unsigned>(i)) : 0.0f;
    const float b = (i + 1 < n) ? __uint_as_float(0x80000000u + static_cast<unsigned>(i + 1)) : 0.0f;
    float ea, eb;
    pk::unpack(pk::exp2x(pk::pack(a, b)), ea, eb);
    bad += (__float_as_uint(ea) != __float_as_uint(expf(a))) + (__float_as_uint(eb) != __float_as_uint(expf(b)));
  }
  bad = static_cast<unsigned long long>(warp_sum(static_cast<long long>(bad)));
  if (lane_id() == 0 && bad) atomicAdd(mismatches, bad);
}

// PX = pixels per thread: 4 (128-bit loads, 128 registers, 2 CTAs/SM) or 2 (64-bit loads, 3 CTAs/SM).
template <int C, int MODE, int PX>
__global__ void __launch_bounds__(kThreadsA, PX == 4 ? 2 : 3) k_softmax_hist(PhaseAArgs a) {
  using VF = typename VecOf<PX>::F;
  using VU = typename VecOf<PX>::U;
  constexpr bool kShared = (MODE == 3 || MODE == 4 || MODE == 5 || MODE == 6);
  constexpr int kCells = (MODE == 3 || MODE == 5) ? C * kTopBins : ((MODE == 4 || MODE == 6) ? C : 1);
  constexpr int kPer = (MODE == 4 || MODE == 6) ? 1 : kTopBins;   // shared cells per class
  __shared__ uint32_t s_top[kCells];
  const int HW4 = static_cast<int>(a.HW / PX);   // vectors per plane
  HistSink<MODE> sink;
  sink.nb = a.nb;
  sink.nbs = row_stride(a.nb);
  sink.s = s_top;
  sink.g = a.hist;
  sink.top0 = (MODE == 4 || MODE == 6) ? a.nb - 1 : (a.nb > kTopBins ? a.nb - kTopBins : 0);
  sink.run_lbl = 0;
  sink.run_cnt = 0;
  if (kShared) {
    for (int i = threadIdx.x; i < kCells; i += kThreadsA) s_top[i] = 0;
    __syncthreads();
  }
  auto flush_top = [&]() {
    if (MODE == 6) sink.run_flush();
    __syncthreads();
    for (int i = threadIdx.x; i < kCells; i += kThreadsA) {
      const uint32_t v = s_top[i];
      if (v) {
        atomicAdd(sink.g + static_cast<size_t>(i / kPer) * sink.nbs + sink.top0 + (i % kPer), v);
        s_top[i] = 0;
      }
    }
    __syncthreads();
  };
  __shared__ int s_sched[2];
  ChunkSched sched;
  sched.init(a.sched, a.n_tiles);
  int cur_group = -1;
  for (; sched.cur < sched.n_chunks; sched.advance(s_sched)) {
  sched.fetch(s_sched);
  const int t0 = sched.cur * kChunkTiles;
  const int t1 = static_cast<int>(min(static_cast<long long>(t0) + kChunkTiles, a.n_tiles));
  int img = t0 / a.tiles_per_image;
  int tile = t0 - img * a.tiles_per_image;
  for (int t = t0; t < t1; ++t) {
    const int group = img / a.group_size;
    if (group != cur_group) {
      if (kShared && cur_group >= 0) flush_top();
      cur_group = group;
      sink.g = a.hist + static_cast<size_t>(group) * C * sink.nbs;
    }
    const int p4 = tile * kThreadsA + threadIdx.x;
    const bool valid = p4 < HW4;
    float v[PX][C];
    if (valid) {
      const VF* src = reinterpret_cast<const VF*>(a.logits + static_cast<size_t>(img) * C * a.HW) + p4;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        float q[PX];
        unpack(__ldcs(src + static_cast<size_t>(c) * HW4), q);
#pragma unroll
        for (int j = 0; j < PX; ++j) v[j][c] = q[j];
      }
    }
    float cf[PX];
    int lb[PX];
    if (valid) {
#pragma unroll
      for (int j = 0; j < PX; ++j) softmax_argmax<C>(v[j], cf[j], lb[j]);
      const size_t o4 = static_cast<size_t>(img) * HW4 + p4;
      reinterpret_cast<VF*>(a.conf)[o4] = pack_f(cf);
      reinterpret_cast<VU*>(a.label)[o4] = pack_u(lb);
    }
    int bins[PX];
#pragma unroll
    for (int j = 0; j < PX; ++j) {
      bins[j] = 0;
      if (valid) bins[j] = min(max(static_cast<int>(fp16_key(cf[j])) - a.key_lo, 0), a.nb - 1);
      else lb[j] = 0;
    }
    sink.template add_px<PX>(valid, lb, bins);
    if (++tile == a.tiles_per_image) {
      tile = 0;
      ++img;
    }
  }
  }
  if (kShared && cur_group >= 0) flush_top();
}

// ---- software-pipelined variant (cp.async staging) ------------------------------------------------------
// The LDG kernel is co-limited by issue slots and by load latency (ncu: long_scoreboard is the top stall; 128
// registers allow only 16 warps/SM and each warp alternates load -> wait -> ~1100 instructions of math).
// Prefetching the next tile straight into registers does not work: the in-flight LDGs share the warp's six
// scoreboard slots with the MUFU results of the math, so every expf ends up waiting for DRAM (measured: 2.4x
// slower, long_scoreboard 8.9 warps/issue).  cp.async (LDGSTS) is tracked by async-group counters instead of
// the register scoreboard, so here every thread owns C x 16 bytes of shared memory: at the top of an
// iteration it pulls its 4 pixels x C channels into registers (19 conflict-free LDS.128), immediately
// re-issues 19 16-byte cp.async for ITS OWN next tile into the same slots, does the math, and only then waits
// for the group.  No block-level barrier, no lock-step phases (unlike the TMA variant below), same 2 CTAs x 8
// warps per SM as the LDG kernel, and global latency fully overlapped with the math inside every warp.
template <int C, int MODE, int PX, int MATH = 0, int OCC = (PX == 4 ? 2 : 3)>
__global__ void __launch_bounds__(kThreadsA, OCC) k_softmax_hist_sp(PhaseAArgs a) {
  using VF = typename VecOf<PX>::F;
  using VU = typename VecOf<PX>::U;
  constexpr bool kShared = (MODE == 6);
  static_assert(MODE == 1 || MODE == 6, "software-pipelined variant: sink 1 or 6");
  extern __shared__ __align__(128) unsigned char s_stage_raw[];   // VF [C][kThreadsA]
  VF* s_stage = reinterpret_cast<VF*>(s_stage_raw);
  __shared__ uint32_t s_top[C];
  const int HW4 = static_cast<int>(a.HW / PX);   // vectors per plane
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
  VF* my = s_stage + threadIdx.x;
  const unsigned my_u32 = static_cast<unsigned>(__cvta_generic_to_shared(my));
  auto prefetch = [&](int img_, int p4_) {
    // byte pointer bumped by the plane stride: two integer instructions per channel instead of four
    const char* src = reinterpret_cast<const char*>(a.logits + static_cast<size_t>(img_) * C * a.HW) +
                      static_cast<size_t>(p4_) * sizeof(VF);
    const size_t plane = static_cast<size_t>(a.HW) * sizeof(float);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      if (PX == 4)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(my_u32 + c * kThreadsA * 16), "l"(src) : "memory");
      else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(my_u32 + c * kThreadsA * 8), "l"(src) : "memory");
      src += plane;
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };
  __shared__ int s_sched[2];
  ChunkSched sched;
  sched.init(a.sched, a.n_tiles);
  int t0 = sched.cur * kChunkTiles;
  int img = t0 / a.tiles_per_image;
  int tile = t0 - img * a.tiles_per_image;
  int cur_group = -1;
  int p4 = tile * kThreadsA + threadIdx.x;
  bool valid = (sched.cur < sched.n_chunks) && (p4 < HW4);
  if (valid) prefetch(img, p4);
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  for (; sched.cur < sched.n_chunks; sched.advance(s_sched)) {
  sched.fetch(s_sched);
  t0 = sched.cur * kChunkTiles;
  const int t1 = static_cast<int>(min(static_cast<long long>(t0) + kChunkTiles, a.n_tiles));
  for (int t = t0; t < t1; ++t) {
    const int group = img / a.group_size;
    if (group != cur_group) {
      if (kShared && cur_group >= 0) flush_top();
      cur_group = group;
      sink.g = a.hist + static_cast<size_t>(group) * C * sink.nbs;
    }
    int nimg = img, ntile = tile + 1;
    bool has_next = true;
    if (t + 1 < t1) {
      if (ntile == a.tiles_per_image) {
        ntile = 0;
        ++nimg;
      }
    } else {  // first tile of the CTA's next chunk
      has_next = sched.nxt < sched.n_chunks;
      const int nt0 = sched.nxt * kChunkTiles;
      nimg = nt0 / a.tiles_per_image;
      ntile = nt0 - nimg * a.tiles_per_image;
    }
    const int np4 = ntile * kThreadsA + threadIdx.x;
    const bool nvalid = has_next && (np4 < HW4);
    float v[PX][C];
    float cf[PX];
    int lb[PX];
    if (valid) {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        float q[PX];
        unpack(my[c * kThreadsA], q);
#pragma unroll
        for (int j = 0; j < PX; ++j) v[j][c] = q[j];
      }
    }
    // all LDS above are consumed by the first max before the slots are overwritten: keep a true dependency
    float guard = 0.f;
    if (valid) {
#pragma unroll
      for (int c = 0; c < C; ++c) guard = fmaxf(guard, v[0][c]);
    }
    if (nvalid && guard == guard) prefetch(nimg, np4);
    if (valid) {
      if (MATH == 1) {
        bool tie[PX];
        bool any_tie = false;
#pragma unroll
        for (int j = 0; j < PX; j += 2) {
          softmax_argmax_pair<C>(v[j], v[j + 1], cf[j], cf[j + 1], lb[j], lb[j + 1], tie[j], tie[j + 1]);
          any_tie |= tie[j] | tie[j + 1];
        }
        if (any_tie) {   // rare: exact / near ties take the scalar walk (same conf bits, first-index label)
#pragma unroll
          for (int j = 0; j < PX; ++j)
            if (tie[j]) softmax_argmax<C>(v[j], cf[j], lb[j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < PX; ++j) softmax_argmax<C>(v[j], cf[j], lb[j]);
      }
      const size_t o4 = static_cast<size_t>(img) * HW4 + p4;
      reinterpret_cast<VF*>(a.conf)[o4] = pack_f(cf);
      reinterpret_cast<VU*>(a.label)[o4] = pack_u(lb);
    }
    int bins[PX];
#pragma unroll
    for (int j = 0; j < PX; ++j) {
      bins[j] = 0;
      if (valid) bins[j] = min(max(static_cast<int>(fp16_key(cf[j])) - a.key_lo, 0), a.nb - 1);
      else lb[j] = 0;
    }
    sink.template add_px<PX>(valid, lb, bins);
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    img = nimg;
    tile = ntile;
    p4 = np4;
    valid = nvalid;
  }
  }
  if (kShared && cur_group >= 0) flush_top();
}

// ---- group-resident variant (shared-memory histogram) ----------------------------------------------------
// tools/membench.cu shows where the remaining time of the cp.async kernels goes: with the packed math, conf/label
// stores and NO histogram the pipeline streams 6.3 TB/s (26.8 us per 19x1024x2048 map); adding the one global RED
// per pixel costs 6.4 us per map (5.1 TB/s).  A RED whose 32 lanes hit 32 different sectors occupies the SM's
// load/store path for 32 request slots, one per clock: 14 170 pixels per SM per map = 7.5 us.  Shared-memory
// atomics do not have that cost, but a per-CTA table only pays off when it is flushed rarely: the (class, key) space
// of a group is 84 k bins against 4.2 M pixels, so a CTA must see >> 84 k pixels of ONE group between flushes.
// Here a work unit is a contiguous slice of one group (>= 100 k pixels), owned by one 512-thread CTA (one per SM):
//   * keys in [hi0, 0x3C00) -- the upper ~2000 fp16 keys, conf >= ~0.26, where softmax confidences live -- are
//     counted in a shared table of 16-bit counters (two per word; a counter that wraps reports itself through the
//     value the atomic returns and moves 65536 to the global row);
//   * key 0x3C00 (conf == 1.0 in fp16, the saturated pixels) keeps the per-thread run-length counters of sink 6;
//   * the rare low keys take the global RED as before;
//   * at the end of a unit the table is added to the global rows with coalesced REDs (1.2 k requests).
// No dynamic tile scheduler and no block barrier inside a unit: units are handed out from a global counter.
// Measured and dropped: 768 threads x 2 px (80 registers, 24 warps): 66 % of peak against 86 % (more instructions per
// pixel); a cp.async.bulk.prefetch.L2 of the tile after next (two-tile look-ahead): 56 %; advancing the two pixel
// pairs of a thread in lock-step through one channel loop (hand-interleaved dependency chains): 5 % slower than
// letting ptxas schedule the two softmax_argmax_pair calls.
// HINT: L2 eviction priorities -- logits are streamed evict_first, the conf / label spill is stored evict_last so
// that a phase C that follows closely (small windows) finds it in L2.
template <int C, int HINT>
__global__ void __launch_bounds__(kThreadsG, 1) k_softmax_hist_gr(GroupArgs ga) {
  const PhaseAArgs& a = ga.a;
  extern __shared__ __align__(128) unsigned char s_raw[];
  float4* s_stage = reinterpret_cast<float4*>(s_raw);                                  // [C][kThreadsG]
  uint32_t* s_tab = reinterpret_cast<uint32_t*>(s_raw + sizeof(float4) * C * kThreadsG);   // [C][words]
  __shared__ uint32_t s_top[C];
  __shared__ int s_unit[2];
  const int HW4 = static_cast<int>(a.HW / 4);
  const int nbs = row_stride(a.nb);
  const int top = a.nb - 1;
  const int hi0 = ga.hi0, words = ga.words;
  for (int i = threadIdx.x; i < C * words; i += kThreadsG) s_tab[i] = 0;
  if (threadIdx.x < C) s_top[threadIdx.x] = 0;
  float4* my = s_stage + threadIdx.x;
  const unsigned my_u32 = static_cast<unsigned>(__cvta_generic_to_shared(my));
  uint64_t pol_first = 0, pol_last = 0;
  if (HINT) {
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
  }
  auto prefetch = [&](int img_, int p4_) {
    const char* src = reinterpret_cast<const char*>(a.logits + static_cast<size_t>(img_) * C * a.HW) +
                      static_cast<size_t>(p4_) * sizeof(float4);
    const size_t plane = static_cast<size_t>(a.HW) * sizeof(float);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      if (HINT)
        asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;\n" ::"r"(my_u32 + c * kThreadsG * 16), "l"(src), "l"(pol_first) : "memory");
      else
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(my_u32 + c * kThreadsG * 16), "l"(src) : "memory");
      src += plane;
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };
  // unit -> (first image of its group, tile range inside the group)
  auto unit_range = [&](int u, int& img0, int& t0, int& t1) {
    const int g = u / ga.slices, sl = u - g * ga.slices;
    img0 = g * a.group_size;
    const int n_img = min(a.group_size, a.n_images - img0);
    const long long tiles = static_cast<long long>(n_img) * a.tiles_per_image;
    t0 = static_cast<int>(tiles * sl / ga.slices);
    t1 = static_cast<int>(tiles * (sl + 1) / ga.slices);
  };
  int cur = blockIdx.x, nxt = blockIdx.x + gridDim.x, par = 0;
  int img0 = 0, t0 = 0, t1 = 0;
  if (cur < ga.n_units) unit_range(cur, img0, t0, t1);
  int img = img0 + t0 / a.tiles_per_image;
  int tile = t0 - (img - img0) * a.tiles_per_image;
  int p4 = tile * kThreadsG + threadIdx.x;
  bool valid = (cur < ga.n_units) && (t0 < t1) && (p4 < HW4);
  if (valid) prefetch(img, p4);
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  __syncthreads();
  int run_lbl = 0;
  unsigned run_cnt = 0;
  while (cur < ga.n_units) {
    if (threadIdx.x == 0) s_unit[par] = static_cast<int>(atomicAdd(a.sched, 1u)) + 2 * static_cast<int>(gridDim.x);
    uint32_t* g_hist = a.hist + static_cast<size_t>(cur / ga.slices) * C * nbs;
    int nimg0 = 0, nt0 = 0, nt1 = 0;
    if (nxt < ga.n_units) unit_range(nxt, nimg0, nt0, nt1);
    for (int t = t0; t < t1; ++t) {
      int nimg = img, ntile = tile + 1;
      bool has_next = true;
      if (t + 1 < t1) {
        if (ntile == a.tiles_per_image) {
          ntile = 0;
          ++nimg;
        }
      } else {  // first tile of this CTA's next unit
        has_next = (nxt < ga.n_units) && (nt0 < nt1);
        nimg = nimg0 + nt0 / a.tiles_per_image;
        ntile = nt0 - (nimg - nimg0) * a.tiles_per_image;
      }
      const int np4 = ntile * kThreadsG + threadIdx.x;
      const bool nvalid = has_next && (np4 < HW4);
      float v[4][C];
      float cf[4];
      int lb[4];
      if (valid) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float4 q = my[c * kThreadsG];
          v[0][c] = q.x; v[1][c] = q.y; v[2][c] = q.z; v[3][c] = q.w;
        }
      }
      float guard = 0.f;   // true dependency: every LDS above retires before the slots are overwritten
      if (valid) {
#pragma unroll
        for (int c = 0; c < C; ++c) guard = fmaxf(guard, v[0][c]);
      }
      if (nvalid && guard == guard) prefetch(nimg, np4);
      if (valid) {
        bool tie[4];
        bool any_tie = false;
#pragma unroll
        for (int j = 0; j < 4; j += 2) {
          softmax_argmax_pair<C>(v[j], v[j + 1], cf[j], cf[j + 1], lb[j], lb[j + 1], tie[j], tie[j + 1]);
          any_tie |= tie[j] | tie[j + 1];
        }
        if (any_tie) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (tie[j]) softmax_argmax<C>(v[j], cf[j], lb[j]);
        }
        const size_t o4 = static_cast<size_t>(img) * HW4 + p4;
        if (HINT) {
          asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;\n" ::"l"(reinterpret_cast<float4*>(a.conf) + o4),
                       "f"(cf[0]), "f"(cf[1]), "f"(cf[2]), "f"(cf[3]), "l"(pol_last) : "memory");
          const unsigned lw = static_cast<unsigned>(lb[0]) | (static_cast<unsigned>(lb[1]) << 8) |
                              (static_cast<unsigned>(lb[2]) << 16) | (static_cast<unsigned>(lb[3]) << 24);
          asm volatile("st.global.L2::cache_hint.b32 [%0], %1, %2;\n" ::"l"(reinterpret_cast<unsigned*>(a.label) + o4), "r"(lw),
                       "l"(pol_last) : "memory");
        } else {
          reinterpret_cast<float4*>(a.conf)[o4] = make_float4(cf[0], cf[1], cf[2], cf[3]);
          reinterpret_cast<uchar4*>(a.label)[o4] = make_uchar4(lb[0], lb[1], lb[2], lb[3]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
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
            const int idx = bin - hi0;
            const unsigned sh = (idx & 1) * 16;
            tab16_add(s_tab + l * words + (idx >> 1), sh, g_hist + static_cast<size_t>(l) * nbs + bin);
          } else {
            atomicAdd(g_hist + static_cast<size_t>(l) * nbs + bin, 1u);
          }
        }
      }
      asm volatile("cp.async.wait_group 0;\n" ::: "memory");
      img = nimg;
      tile = ntile;
      p4 = np4;
      valid = nvalid;
    }
    // end of the unit: add the shared table and the top-key counters to the group's global rows
    if (run_cnt) atomicAdd(s_top + run_lbl, run_cnt);
    run_cnt = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < C * words; i += kThreadsG) {
      const uint32_t w = s_tab[i];
      if (w) {
        const int c = i / words, k = i - c * words;
        uint32_t* row = g_hist + static_cast<size_t>(c) * nbs + hi0 + 2 * k;
        if (w & 0xffffu) atomicAdd(row, w & 0xffffu);
        if (w >> 16) atomicAdd(row + 1, w >> 16);
        s_tab[i] = 0;
      }
    }
    if (threadIdx.x < C) {
      const uint32_t w = s_top[threadIdx.x];
      if (w) {
        atomicAdd(g_hist + static_cast<size_t>(threadIdx.x) * nbs + top, w);
        s_top[threadIdx.x] = 0;
      }
    }
    __syncthreads();
    const int nn = s_unit[par];
    par ^= 1;
    cur = nxt;
    nxt = nn;
    img0 = nimg0;
    t0 = nt0;
    t1 = nt1;
  }
}

// Static variant of the same kernel: the window's tiles are split into one contiguous range per CTA (image order) and
// a CTA flushes its table whenever its range crosses a group boundary.  One or two flushes per CTA instead of one per
// unit: the unit hand-over of the dynamic version (table flush with ~20 k REDs, two barriers) costs ~4 % at 8 units
// per CTA; with one CTA per SM the SMs progress evenly enough that dynamic balancing buys nothing.
template <int C, int HINT>
__global__ void __launch_bounds__(kThreadsG, 1) k_softmax_hist_grs(GroupArgs ga) {
  const PhaseAArgs& a = ga.a;
  extern __shared__ __align__(128) unsigned char s_raw[];
  float4* s_stage = reinterpret_cast<float4*>(s_raw);                                  // [C][kThreadsG]
  uint32_t* s_tab = reinterpret_cast<uint32_t*>(s_raw + sizeof(float4) * C * kThreadsG);   // [C][words]
  __shared__ uint32_t s_top[C];
  const int HW4 = static_cast<int>(a.HW / 4);
  const int nbs = row_stride(a.nb);
  const int top = a.nb - 1;
  const int hi0 = ga.hi0, words = ga.words;
  for (int i = threadIdx.x; i < C * words; i += kThreadsG) s_tab[i] = 0;
  if (threadIdx.x < C) s_top[threadIdx.x] = 0;
  float4* my = s_stage + threadIdx.x;
  const unsigned my_u32 = static_cast<unsigned>(__cvta_generic_to_shared(my));
  uint64_t pol_first = 0, pol_last = 0;
  if (HINT) {
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
  }
  auto prefetch = [&](int img_, int p4_) {
    const char* src = reinterpret_cast<const char*>(a.logits + static_cast<size_t>(img_) * C * a.HW) +
                      static_cast<size_t>(p4_) * sizeof(float4);
    const size_t plane = static_cast<size_t>(a.HW) * sizeof(float);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      if (HINT)
        asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;\n" ::"r"(my_u32 + c * kThreadsG * 16), "l"(src), "l"(pol_first) : "memory");
      else
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(my_u32 + c * kThreadsG * 16), "l"(src) : "memory");
      src += plane;
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };
  // static split: CTA b owns the tiles [T b / grid, T (b + 1) / grid) of the window in image order
  const long long lo = a.n_tiles * blockIdx.x / gridDim.x, hi = a.n_tiles * (blockIdx.x + 1) / gridDim.x;
  int img = static_cast<int>(lo / a.tiles_per_image);
  int tile = static_cast<int>(lo - static_cast<long long>(img) * a.tiles_per_image);
  int p4 = tile * kThreadsG + threadIdx.x;
  bool valid = (lo < hi) && (p4 < HW4);
  if (valid) prefetch(img, p4);
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  __syncthreads();
  int run_lbl = 0;
  unsigned run_cnt = 0;
  const long long tiles_per_group = static_cast<long long>(a.tiles_per_image) * a.group_size;
  for (long long piece = lo; piece < hi;) {
    // the part of the CTA's range that lies in one group: no barrier inside, one flush at its end
    const int cur_group = img / a.group_size;
    const long long piece_end = min(hi, (static_cast<long long>(cur_group) + 1) * tiles_per_group);
    uint32_t* g_hist = a.hist + static_cast<size_t>(cur_group) * C * nbs;
    const int n_piece = static_cast<int>(piece_end - piece);
    const bool more = piece_end < hi;
    for (int t = 0; t < n_piece; ++t) {
      int nimg = img, ntile = tile + 1;
      const bool has_next = (t + 1 < n_piece) || more;
      if (ntile == a.tiles_per_image) {
        ntile = 0;
        ++nimg;
      }
      const int np4 = ntile * kThreadsG + threadIdx.x;
      const bool nvalid = has_next && (np4 < HW4);
      float v[4][C];
      float cf[4];
      int lb[4];
      if (valid) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float4 q = my[c * kThreadsG];
          v[0][c] = q.x; v[1][c] = q.y; v[2][c] = q.z; v[3][c] = q.w;
        }
      }
      float guard = 0.f;   // true dependency: every LDS above retires before the slots are overwritten
      if (valid) {
#pragma unroll
        for (int c = 0; c < C; ++c) guard = fmaxf(guard, v[0][c]);
      }
      if (nvalid && guard == guard) prefetch(nimg, np4);
      if (valid) {
        bool tie[4];
        bool any_tie = false;
#pragma unroll
        for (int j = 0; j < 4; j += 2) {
          softmax_argmax_pair<C>(v[j], v[j + 1], cf[j], cf[j + 1], lb[j], lb[j + 1], tie[j], tie[j + 1]);
          any_tie |= tie[j] | tie[j + 1];
        }
        if (any_tie) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (tie[j]) softmax_argmax<C>(v[j], cf[j], lb[j]);
        }
        const size_t o4 = static_cast<size_t>(img) * HW4 + p4;
        if (HINT) {
          asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;\n" ::"l"(reinterpret_cast<float4*>(a.conf) + o4),
                       "f"(cf[0]), "f"(cf[1]), "f"(cf[2]), "f"(cf[3]), "l"(pol_last) : "memory");
          const unsigned lw = static_cast<unsigned>(lb[0]) | (static_cast<unsigned>(lb[1]) << 8) |
                              (static_cast<unsigned>(lb[2]) << 16) | (static_cast<unsigned>(lb[3]) << 24);
          asm volatile("st.global.L2::cache_hint.b32 [%0], %1, %2;\n" ::"l"(reinterpret_cast<unsigned*>(a.label) + o4), "r"(lw),
                       "l"(pol_last) : "memory");
        } else {
          reinterpret_cast<float4*>(a.conf)[o4] = make_float4(cf[0], cf[1], cf[2], cf[3]);
          reinterpret_cast<uchar4*>(a.label)[o4] = make_uchar4(lb[0], lb[1], lb[2], lb[3]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
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
            const int idx = bin - hi0;
            const unsigned sh = (idx & 1) * 16;
            tab16_add(s_tab + l * words + (idx >> 1), sh, g_hist + static_cast<size_t>(l) * nbs + bin);
          } else {
            atomicAdd(g_hist + static_cast<size_t>(l) * nbs + bin, 1u);
          }
        }
      }
      asm volatile("cp.async.wait_group 0;\n" ::: "memory");
      img = nimg;
      tile = ntile;
      p4 = np4;
      valid = nvalid;
    }
    piece = piece_end;
    // the CTA leaves the group: add the shared table and the top-key counters to its global rows
    if (run_cnt) atomicAdd(s_top + run_lbl, run_cnt);
    run_cnt = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < C * words; i += kThreadsG) {
      const uint32_t w = s_tab[i];
      if (w) {
        const int c = i / words, kk = i - c * words;
        uint32_t* row = g_hist + static_cast<size_t>(c) * nbs + hi0 + 2 * kk;
        if (w & 0xffffu) atomicAdd(row, w & 0xffffu);
        if (w >> 16) atomicAdd(row + 1, w >> 16);
        s_tab[i] = 0;
      }
    }
    if (threadIdx.x < C) {
      const uint32_t w = s_top[threadIdx.x];
      if (w) {
        atomicAdd(g_hist + static_cast<size_t>(threadIdx.x) * nbs + top, w);
        s_top[threadIdx.x] = 0;
      }
    }
    __syncthreads();
  }
}

// ---- TMA-staged variant --------------------------------------------------------------------------------
// Same arithmetic, different data movement: [C x kTileT] channel tiles are streamed into shared memory with
// bulk async copies (cp.async.bulk, the 1-D TMA path; SASS UBLKCP) that complete on an mbarrier.  The CTA
// is two self-feeding groups of eight warps, each owning one 77.8 KB stage: a group waits on its stage's
// barrier, pulls its 4 pixels x C channels into registers with 128-bit LDS, syncs (named barrier), one
// elected thread immediately issues the bulk copies of the group's NEXT tile into the now free stage, and
// all 256 threads do the math while that copy is in flight.  Global-memory latency is hidden by up to two
// stages (155 KB per SM) in flight instead of by occupancy; the 16 warps only ever wait on LDS.
constexpr int kTileT = 1024;                       // pixels per stage
constexpr int kGroupsT = 2;                        // consumer groups == stages
constexpr int kGroupThreadsT = kTileT / 4;         // 256: one thread per 4 pixels of a stage
constexpr int kGroupWarpsT = kGroupThreadsT / 32;  // 8
constexpr int kThreadsT = kGroupsT * kGroupThreadsT;   // 512 threads x 128 registers = the whole register file

__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n" ::
          "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ void group_sync(int grp) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(grp + 1), "n"(kGroupThreadsT) : "memory");
}

template <int C, int MODE>
__global__ void __launch_bounds__(kThreadsT, 1) k_softmax_hist_tma(PhaseAArgs a) {
  static_assert(MODE == 1 || MODE == 6, "TMA variant: plain REDs (1) or thread-run top-key aggregation (6)");
  constexpr bool kShared = (MODE == 6);
  constexpr int kCells = C;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* stage_buf = reinterpret_cast<float*>(smem_raw);                                   // [kGroups][C][kTile]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + sizeof(float) * kGroupsT * C * kTileT);
  uint32_t* s_top_all = reinterpret_cast<uint32_t*>(full_bar + kGroupsT);                  // [kGroups][C]
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kGroupsT; ++i) mbar_init(full_bar + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  for (int i = threadIdx.x; i < kGroupsT * kCells; i += kThreadsT) s_top_all[i] = 0;
  __syncthreads();
  const int t0 = static_cast<int>(a.n_tiles * blockIdx.x / gridDim.x);
  const int t1 = static_cast<int>(a.n_tiles * (blockIdx.x + 1) / gridDim.x);
  const int grp = warp / kGroupWarpsT;
  const int gtid = threadIdx.x - grp * kGroupThreadsT;
  const int HW4 = static_cast<int>(a.HW >> 2);
  float* my_stage = stage_buf + static_cast<size_t>(grp) * C * kTileT;
  uint64_t* my_bar = full_bar + grp;
  uint64_t policy = 0;
  if (gtid == 0) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(policy));
  // one elected thread per group issues the C bulk copies of a tile into the group's stage
  auto fill = [&](int img_, int tile_) {
    const int64_t px0 = static_cast<int64_t>(tile_) * kTileT;
    const unsigned bytes = static_cast<unsigned>(min(static_cast<int64_t>(kTileT), a.HW - px0)) * 4u;
    mbar_expect_tx(my_bar, bytes * C);
    const float* src = a.logits + static_cast<size_t>(img_) * C * a.HW + px0;
#pragma unroll 1
    for (int c = 0; c < C; ++c) bulk_g2s(my_stage + c * kTileT, src + static_cast<size_t>(c) * a.HW, bytes, my_bar, policy);
  };
  HistSink<MODE> sink;
  sink.nb = a.nb;
  sink.nbs = row_stride(a.nb);
  sink.s = s_top_all + grp * kCells;
  sink.g = a.hist;
  sink.top0 = a.nb - 1;
  sink.run_lbl = 0;
  sink.run_cnt = 0;
  auto flush_top = [&]() {
    sink.run_flush();
    group_sync(grp);
    for (int i = gtid; i < kCells; i += kGroupThreadsT) {
      const uint32_t v = sink.s[i];
      if (v) {
        atomicAdd(sink.g + static_cast<size_t>(i) * sink.nbs + sink.top0, v);
        sink.s[i] = 0;
      }
    }
    group_sync(grp);
  };
  int img = (t0 + grp) / a.tiles_per_image;
  int tile = (t0 + grp) - img * a.tiles_per_image;
  if (gtid == 0 && t0 + grp < t1) fill(img, tile);
  int cur_group = -1;
  const float4* stage = reinterpret_cast<const float4*>(my_stage) + gtid;
  for (int t = t0 + grp; t < t1; t += kGroupsT) {
    const unsigned ph = ((t - t0) / kGroupsT) & 1;
    const int group = img / a.group_size;
    if (group != cur_group) {
      if (kShared && cur_group >= 0) flush_top();
      cur_group = group;
      sink.g = a.hist + static_cast<size_t>(group) * C * sink.nbs;
    }
    const int p4 = tile * kGroupThreadsT + gtid;
    const bool valid = p4 < HW4;
    int nimg = img, ntile = tile + kGroupsT;           // the group's next tile
    while (ntile >= a.tiles_per_image) {
      ntile -= a.tiles_per_image;
      ++nimg;
    }
    float v[4][C];
    mbar_wait(my_bar, ph);
    if (valid) {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float4 q = stage[c * (kTileT / 4)];
        v[0][c] = q.x; v[1][c] = q.y; v[2][c] = q.z; v[3][c] = q.w;
      }
    }
    group_sync(grp);                                    // the stage is in registers: refill it before the math
    if (gtid == 0 && t + kGroupsT < t1) fill(nimg, ntile);
    float cf[4];
    int lb[4];
    if (valid) {
#pragma unroll
      for (int j = 0; j < 4; ++j) softmax_argmax<C>(v[j], cf[j], lb[j]);
      const size_t o4 = static_cast<size_t>(img) * HW4 + p4;
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
    img = nimg;
    tile = ntile;
  }
  if (kShared && cur_group >= 0) flush_top();
}

template <int C>
constexpr size_t tma_smem_bytes() {
  return sizeof(float) * kGroupsT * C * kTileT + kGroupsT * sizeof(uint64_t) + sizeof(uint32_t) * kGroupsT * C;
}

// Scalar path: any C, any HW.  One pixel per thread; correctness path for odd shapes.
__global__ void __launch_bounds__(kThreadsA) k_softmax_hist_generic(PhaseAArgs a) {
  const long long total = static_cast<long long>(a.n_images) * a.HW;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int img = static_cast<int>(i / a.HW);
    const int64_t p = i - static_cast<long long>(img) * a.HW;
    float cf;
    int lb;
    softmax_argmax_generic(a.logits + static_cast<size_t>(img) * a.C * a.HW + p, a.HW, a.C, cf, lb);
    a.conf[i] = cf;
    a.label[i] = static_cast<uint8_t>(lb);
    int bin = static_cast<int>(fp16_key(cf)) - a.key_lo;
    bin = min(max(bin, 0), a.nb - 1);
    atomicAdd(a.hist + (static_cast<size_t>(img / a.group_size) * a.C + lb) * row_stride(a.nb) + bin, 1u);
  }
}

}  // namespace hiast

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
using namespace hiast;

extern "C" int hiast_ias_key_lo(int C) {
  if (C < 1 || C > HIAST_MAX_CLASSES) return HIAST_ERR_INVALID_ARG;
  const float v = 1.0f / static_cast<float>(C);
  return static_cast<int>(__half_as_ushort(__float2half_rn(v)));
}

extern "C" int hiast_ias_hist_row_stride(int key_lo) {
  if (key_lo < 0 || key_lo > HIAST_KEY_ONE) return HIAST_ERR_INVALID_ARG;
  return row_stride(HIAST_KEY_ONE - key_lo + 1);
}

extern "C" size_t hiast_ias_hist_bytes(int n_groups, int C, int key_lo) {
  if (n_groups < 0 || C < 1 || key_lo < 0 || key_lo > HIAST_KEY_ONE) return 0;
  return static_cast<size_t>(n_groups) * C * row_stride(HIAST_KEY_ONE - key_lo + 1) * sizeof(uint32_t);
}

namespace hiast {
// One zeroed work counter per launch, taken round-robin from a static device array (stream-ordered memset
// before the kernel; a slot is reused only after 1023 later launches).
constexpr int kSchedSlots = 1024;
__device__ unsigned g_sched_slots[kSchedSlots];

int next_sched_slot(unsigned** out, cudaStream_t st) {
  static std::atomic<unsigned*> bases[64];        // a __device__ symbol has one address PER DEVICE
  static std::atomic<unsigned> next{0};
  int dev = 0;
  HIAST_CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return HIAST_ERR_UNSUPPORTED;
  unsigned* base = bases[dev].load(std::memory_order_acquire);
  if (!base) {
    void* p = nullptr;
    HIAST_CUDA_TRY(cudaGetSymbolAddress(&p, g_sched_slots));
    base = static_cast<unsigned*>(p);
    bases[dev].store(base, std::memory_order_release);
  }
  unsigned* slot = base + (next.fetch_add(1) % kSchedSlots);
  HIAST_CUDA_TRY(cudaMemsetAsync(slot, 0, sizeof(unsigned), st));
  *out = slot;
  return HIAST_OK;
}
}  // namespace hiast

namespace {

template <int C, int MODE>
int launch_phase_a_tma(PhaseAArgs a, cudaStream_t st) {
  constexpr size_t smem = tma_smem_bytes<C>();
  HIAST_TRY(ensure_dyn_smem(k_softmax_hist_tma<C, MODE>, smem));
  a.tiles_per_image = static_cast<int>((a.HW + kTileT - 1) / kTileT);
  a.n_tiles = static_cast<long long>(a.tiles_per_image) * a.n_images;
  int grid = sm_count();
  if (grid > a.n_tiles) grid = static_cast<int>(a.n_tiles);
  k_softmax_hist_tma<C, MODE><<<grid, kThreadsT, smem, st>>>(a);
  HIAST_CHECK_LAUNCH();
  return HIAST_OK;
}


template <int C, int MODE, int PX, int MATH = 0, int OCC = (PX == 4 ? 2 : 3)>
int launch_phase_a_sp(PhaseAArgs a, cudaStream_t st) {
  const int64_t vecs = a.HW / PX;
  a.tiles_per_image = static_cast<int>((vecs + kThreadsA - 1) / kThreadsA);
  a.n_tiles = static_cast<long long>(a.tiles_per_image) * a.n_images;
  constexpr size_t smem = sizeof(float) * PX * C * kThreadsA;
  HIAST_TRY(ensure_dyn_smem(k_softmax_hist_sp<C, MODE, PX, MATH, OCC>, smem));
  int grid = resident_grid(k_softmax_hist_sp<C, MODE, PX, MATH, OCC>, kThreadsA, smem);
  const long long n_chunks = (a.n_tiles + kChunkTiles - 1) / kChunkTiles;
  if (grid > n_chunks) grid = static_cast<int>(n_chunks);
  const int rc = next_sched_slot(&a.sched, st);
  if (rc != HIAST_OK) return rc;
  k_softmax_hist_sp<C, MODE, PX, MATH, OCC><<<grid, kThreadsA, smem, st>>>(a);
  HIAST_CHECK_LAUNCH();
  return HIAST_OK;
}

template <int C, int MODE, int PX>
int launch_phase_a_ldg(PhaseAArgs a, cudaStream_t st) {
  const int64_t vecs = a.HW / PX;
  a.tiles_per_image = static_cast<int>((vecs + kThreadsA - 1) / kThreadsA);
  a.n_tiles = static_cast<long long>(a.tiles_per_image) * a.n_images;
  int grid = resident_grid(k_softmax_hist<C, MODE, PX>, kThreadsA, 0);
  const long long n_chunks = (a.n_tiles + kChunkTiles - 1) / kChunkTiles;
  if (grid > n_chunks) grid = static_cast<int>(n_chunks);
  const int rc = next_sched_slot(&a.sched, st);
  if (rc != HIAST_OK) return rc;
  k_softmax_hist<C, MODE, PX><<<grid, kThreadsA, 0, st>>>(a);
  HIAST_CHECK_LAUNCH();
  return HIAST_OK;
}

// Shared-memory budget of the group-resident kernel: 227 KB per CTA minus the cp.async staging buffers.
template <int C, int HINT, int STATIC = 0>
int launch_phase_a_gr(PhaseAArgs a, cudaStream_t st, int reserve_sms = 0) {
  constexpr size_t kStage = sizeof(float4) * C * kThreadsG;
  constexpr size_t kBudget = 227 * 1024 - 1024;   // static shared memory + reserve
  static_assert(kStage + 4096 < kBudget, "staging does not fit");
  GroupArgs ga;
  const int top = a.nb - 1;                        // bins [0, top) can live in the table; bin top has its own counters
  int words = static_cast<int>((kBudget - kStage) / (sizeof(uint32_t) * C));
  words = std::min(words, (top + 1) / 2);
  ga.words = words;
  ga.hi0 = std::max(top - 2 * words, 0);
  // a table pair may straddle bin `top` when hi0 == 0 and top is odd: bin top is never counted in the table and the
  // flush adds zero there, so the extra slot is harmless (rows are padded to a multiple of 4 words)
  const int64_t vecs = a.HW / 4;
  a.tiles_per_image = static_cast<int>((vecs + kThreadsG - 1) / kThreadsG);
  a.n_tiles = static_cast<long long>(a.tiles_per_image) * a.n_images;
  const int n_groups = (a.n_images + a.group_size - 1) / a.group_size;
  const int sms = sm_count();
  // slices per group: the smallest count that keeps every SM busy in the last round (>= 95 % of the best reachable)
  const long long tiles_per_group = static_cast<long long>(a.tiles_per_image) * a.group_size;
  const int max_slices = static_cast<int>(std::max<long long>(1, std::min<long long>(256, tiles_per_group / 32)));
  int best = 1;
  double best_eff = 0.0;
  for (int sl = 1; sl <= max_slices; ++sl) {
    const long long units = static_cast<long long>(n_groups) * sl;
    const long long rounds = (units + sms - 1) / sms;
    const double eff = static_cast<double>(units) / static_cast<double>(rounds * sms);
    if (eff > best_eff + 0.02) {
      best_eff = eff;
      best = sl;
    }
  }
  ga.slices = best;
  ga.n_units = n_groups * best;
  const size_t smem = kStage + sizeof(uint32_t) * C * words;
  ga.a = a;
  if constexpr (STATIC) {
    HIAST_TRY(ensure_dyn_smem(k_softmax_hist_grs<C, HINT>, kBudget));
    // reserve_sms: SMs left WITHOUT a phase-A CTA.  A CTA of this kernel takes a whole SM (512 threads x 128 registers), so nothing
    // else can share an SM with it; the SMs left free are where the threshold chain and phase C of the windows before run
    // CONCURRENTLY with this launch (hiast_b200/sharded.py) instead of between two launches.
    const int grid_s = static_cast<int>(std::min<long long>(std::max(1, sms - std::max(0, reserve_sms)), a.n_tiles));
    k_softmax_hist_grs<C, HINT><<<grid_s, kThreadsG, smem, st>>>(ga);
    HIAST_CHECK_LAUNCH();
    return HIAST_OK;
  } else {
    HIAST_TRY(ensure_dyn_smem(k_softmax_hist_gr<C, HINT>, kBudget));
    const int grid = std::min(sms, ga.n_units);
    const int rc = next_sched_slot(&ga.a.sched, st);
    if (rc != HIAST_OK) return rc;
    k_softmax_hist_gr<C, HINT><<<grid, kThreadsG, smem, st>>>(ga);
    HIAST_CHECK_LAUNCH();
    return HIAST_OK;
  }
}

// hist_mode = 10 * pipeline + sink.  pipeline 0: 128-bit LDG, 4 px/thread; 1: TMA-staged; 2: 64-bit LDG, 2 px/thread;
// 3: cp.async software pipeline, 4 px/thread; 4: cp.async, 2 px/thread; 5: as 3 with the packed (f32x2) math;
// 6 / 7: as 4 with the packed math at 3 / 4 CTAs per SM; 80: group-resident kernel (packed math + shared-memory
// histogram) with dynamic units; 81: 80 with L2 eviction hints; 83: group-resident kernel with a static split (the
// default).  sink: see HistSink.  0 = library default.
// NOTE the staging buffers must start on a 128-byte line (extern __shared__ __align__(128)): with a 16-byte aligned
// base every quarter-warp cp.async straddles two lines and the SM issues twice the shared-memory wavefronts AND twice
// the L2 sector requests (ncu: 32 sectors per LDGSTS instead of 16) -- a silent 15-25 % loss.
constexpr int kDefaultHistMode = 83;

template <int C>
int launch_phase_a(const PhaseAArgs& a, int mode_and_reserve, cudaStream_t st) {
  int mode = mode_and_reserve & 0xff;
  const int reserve_sms = (mode_and_reserve >> 8) & 0xff;
  if (mode == 0) mode = kDefaultHistMode;
  switch (mode) {
    case 1: return launch_phase_a_ldg<C, 1, 4>(a, st);        // plain kernel, one global RED per pixel: the cross-check
    case 83: return launch_phase_a_gr<C, 0, 1>(a, st, reserve_sms);   // the product kernel
#ifdef HIAST_DEV_VARIANTS
    // measured and dropped (profiles/): compiled only into the development build
    case 2: return launch_phase_a_ldg<C, 2, 4>(a, st);
    case 3: return launch_phase_a_ldg<C, 3, 4>(a, st);
    case 4: return launch_phase_a_ldg<C, 4, 4>(a, st);
    case 5: return launch_phase_a_ldg<C, 5, 4>(a, st);
    case 6: return launch_phase_a_ldg<C, 6, 4>(a, st);
    case 11: return launch_phase_a_tma<C, 1>(a, st);
    case 16: return launch_phase_a_tma<C, 6>(a, st);
    case 31: return launch_phase_a_sp<C, 1, 4>(a, st);
    case 36: return launch_phase_a_sp<C, 6, 4>(a, st);
    case 41: return launch_phase_a_sp<C, 1, 2>(a, st);
    case 46: return launch_phase_a_sp<C, 6, 2>(a, st);
    case 51: return launch_phase_a_sp<C, 1, 4, 1>(a, st);
    case 56: return launch_phase_a_sp<C, 6, 4, 1>(a, st);
    case 61: return launch_phase_a_sp<C, 1, 2, 1>(a, st);
    case 66: return launch_phase_a_sp<C, 6, 2, 1>(a, st);
    case 80: return launch_phase_a_gr<C, 0>(a, st);
    case 81: return launch_phase_a_gr<C, 1>(a, st);
    case 71: return launch_phase_a_sp<C, 1, 2, 1, 4>(a, st);
    case 76: return launch_phase_a_sp<C, 6, 2, 1, 4>(a, st);
    case 21: return launch_phase_a_ldg<C, 1, 2>(a, st);
    case 25: return launch_phase_a_ldg<C, 5, 2>(a, st);
    case 26: return launch_phase_a_ldg<C, 6, 2>(a, st);
    default: return HIAST_ERR_INVALID_ARG;
#else
    default: return HIAST_ERR_UNSUPPORTED;                    // a development variant: build with HIAST_DEV_VARIANTS=1
#endif
  }
}

}  // namespace

extern "C" int hiast_selftest_packed_expf(unsigned long long* mismatches_dev, void* stream) {
  if (!mismatches_dev) return HIAST_ERR_INVALID_ARG;
  cudaStream_t st = as_stream(stream);
  HIAST_CUDA_TRY(cudaMemsetAsync(mismatches_dev, 0, sizeof(unsigned long long), st));
  k_selftest_packed_expf<<<sm_count() * 8, 256, 0, st>>>(mismatches_dev);
  HIAST_CHECK_LAUNCH();
  return HIAST_OK;
}

extern "C" int hiast_ias_softmax_hist(const float* logits, int n_images, int C, int H, int W, int group_size,
                                      int key_lo, int accumulate, int hist_mode, float* conf, uint8_t* label,
                                      uint32_t* hist, void* stream) {
  if (!logits || !conf || !label || !hist) return HIAST_ERR_INVALID_ARG;
  if (n_images < 0 || C < 1 || C > HIAST_MAX_CLASSES || H < 1 || W < 1 || group_size < 1) return HIAST_ERR_INVALID_ARG;
  if (key_lo < 0 || key_lo > HIAST_KEY_ONE) return HIAST_ERR_INVALID_ARG;
  if (hist_mode < 0 || (hist_mode & 0xff) > 99 || (hist_mode >> 16) != 0) return HIAST_ERR_INVALID_ARG;
  cudaStream_t st = as_stream(stream);
  const int n_groups = (n_images + group_size - 1) / group_size;
  if (!accumulate) HIAST_CUDA_TRY(cudaMemsetAsync(hist, 0, hiast_ias_hist_bytes(n_groups, C, key_lo), st));
  if (n_images == 0) return HIAST_OK;
  PhaseAArgs a;
  a.logits = logits; a.conf = conf; a.label = label; a.hist = hist;
  a.n_images = n_images; a.C = C; a.HW = static_cast<int64_t>(H) * W;
  a.group_size = group_size; a.key_lo = key_lo; a.nb = HIAST_KEY_ONE - key_lo + 1;
  a.sched = nullptr;
  const bool aligned = (a.HW % 4 == 0) && (reinterpret_cast<uintptr_t>(logits) % 16 == 0) &&
                       (reinterpret_cast<uintptr_t>(conf) % 16 == 0) && (reinterpret_cast<uintptr_t>(label) % 4 == 0);
  if (aligned && (C == 19 || C == 16)) {
    if (C == 19) return launch_phase_a<19>(a, hist_mode, st);
    return launch_phase_a<16>(a, hist_mode, st);
  }
  a.tiles_per_image = 0;
  a.n_tiles = 0;
  const long long total = static_cast<long long>(n_images) * a.HW;
  const int grid = static_cast<int>(std::min<long long>((total + kThreadsA - 1) / kThreadsA,
                                                        static_cast<long long>(sm_count()) * 8));
  k_softmax_hist_generic<<<grid, kThreadsA, 0, st>>>(a);
  HIAST_CHECK_LAUNCH();
  return HIAST_OK;
}

