// (1d) IAS phases A + B + C of a window in one persistent kernel (single GPU; opt-in).
#include "ias_common.cuh"

namespace hiast {

// ------------------------------------------------------------------------------------------
// fused window: phases A, B and C of a window in ONE persistent kernel (single GPU)
// ------------------------------------------------------------------------------------------
// The three-kernel pipeline moves 87 B/px (phase A spills conf f32 + label u8, phase C reads them back) and pays
// the threshold chain and phase C as separate passes.  Here the spill never leaves L2:
//   * A-units are the units of k_softmax_hist_gr (a slice of one group, shared-memory histogram), handed out in
//     order; logits are streamed with L2 evict_first, the conf / label spill is stored evict_last;
//   * the CTA that completes the last A-unit of group g becomes its CLOSER: it waits for group g-1 to be closed,
//     stages the group's histogram rows in shared memory, builds their prefix sums (one warp per class), runs the
//     threshold step (same ias_threshold_step as k_threshold_scan) and publishes thr[g] -- the serial chain costs
//     one CTA ~15 us per group while 147 others keep streaming;
//   * C-units (the same slices) become available when their group is closed; a CTA that finishes a unit takes a
//     C-unit first if there is one: conf / label are still in L2 (two groups = 42 MB are in flight), the mask /
//     count / confidence-sum pass costs L2 reads and 1 B/px of stores, and the lines it has consumed are
//     discarded (discard.global.L2) so that the spill is never written back to HBM.
// Waiting happens only (i) in a closer for the previous group's closer and (ii) at the very end for the last
// thresholds; units are claimed by RUNNING CTAs only, so there is no co-residency requirement and no deadlock; every
// spin has a bail-out that raises error bit 4 instead of hanging.  Multi-GPU runs keep the three kernels: the
// thresholds of a window arrive from another rank long after its phase A.
struct FusedArgs {
  GroupArgs ga;
  double alpha, beta, gamma;
  double* thr_state;               // f64 [C] in / out
  double* thr_groups;              // f64 [G][C]
  float* temp_groups;              // f32 [G][C] (may be null)
  uint8_t* plbl;
  unsigned long long* counts;      // [n_images][C]
  unsigned long long* confsum;     // [G][C]
  int* error_flag;
  unsigned* ws;                    // [0] next A-unit, [1] next C-unit, [2] closed groups, [4 + g] finished A-units of g
  int n_groups;
  int discard;
  unsigned long long* trace;       // development: [cta][kTraceEvents][4] 6 words per unit, see the kernel) or null
};

constexpr int kTraceEvents = 256;
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

enum { kUnitNone = 0, kUnitA = 1, kUnitC = 2, kUnitDone = 3 };
constexpr unsigned kSpinLimit = 1u << 23;   // x ~0.25 us: about two seconds, then error bit 4

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// thread 0 only
__device__ int fused_select_unit(const FusedArgs& f, bool& a_exhausted, bool spin, int& idx) {
  const unsigned n_units = static_cast<unsigned>(f.ga.n_units), slices = static_cast<unsigned>(f.ga.slices);
  unsigned spins = 0;
  for (;;) {
    const unsigned closed = ld_acquire_u32(f.ws + 2);
    const unsigned cn = ld_relaxed_u32(f.ws + 1);
    if (cn < closed * slices) {
      if (atomicCAS(f.ws + 1, cn, cn + 1) == cn) {
        idx = static_cast<int>(cn);
        return kUnitC;
      }
      continue;
    }
    if (!a_exhausted) {
      const unsigned an = atomicAdd(f.ws, 1u);
      if (an < n_units) {
        idx = static_cast<int>(an);
        return kUnitA;
      }
      a_exhausted = true;
    }
    if (cn >= n_units) return kUnitDone;
    if (!spin) return kUnitNone;
    __nanosleep(200);
    if (++spins > kSpinLimit) {
      atomicOr(f.error_flag, 4);
      return kUnitDone;
    }
  }
}

#ifdef HIAST_DEV_VARIANTS
template <int C, int DISCARD>
__global__ void __launch_bounds__(kThreadsG, 1) k_ias_fused(FusedArgs f) {
  const GroupArgs& ga = f.ga;
  const PhaseAArgs& a = ga.a;
  extern __shared__ __align__(128) unsigned char s_raw[];
  float4* s_stage = reinterpret_cast<float4*>(s_raw);                                       // [C][kThreadsG]
  unsigned long long* s_acc = reinterpret_cast<unsigned long long*>(s_raw);                 // C-units: [C][kThreadsG]
  uint32_t* s_tab = reinterpret_cast<uint32_t*>(s_raw + sizeof(float4) * C * kThreadsG);    // [C][words]
  __shared__ uint32_t s_top[C];
  __shared__ float s_thr[256];
  __shared__ int s_sel[2];
  __shared__ int s_closer;
  const int HW4 = static_cast<int>(a.HW / 4);
  const int nbs = row_stride(a.nb);
  const int top = a.nb - 1;
  const int hi0 = ga.hi0, words = ga.words;
  for (int i = threadIdx.x; i < C * words; i += kThreadsG) s_tab[i] = 0;
  if (threadIdx.x < C) s_top[threadIdx.x] = 0;
  float4* my = s_stage + threadIdx.x;
  const unsigned my_u32 = static_cast<unsigned>(__cvta_generic_to_shared(my));
  uint64_t pol_first, pol_last;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
  auto prefetch = [&](int img_, int p4_) {
    const char* src = reinterpret_cast<const char*>(a.logits + static_cast<size_t>(img_) * C * a.HW) +
                      static_cast<size_t>(p4_) * sizeof(float4);
    const size_t plane = static_cast<size_t>(a.HW) * sizeof(float);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;\n" ::"r"(my_u32 + c * kThreadsG * 16), "l"(src),
                   "l"(pol_first) : "memory");
      src += plane;
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };
  auto unit_range = [&](int u, int& g, int& img0, int& t0, int& t1) {
    g = u / ga.slices;
    const int sl = u - g * ga.slices;
    img0 = g * a.group_size;
    const int n_img = min(a.group_size, a.n_images - img0);
    const long long tiles = static_cast<long long>(n_img) * a.tiles_per_image;
    t0 = static_cast<int>(tiles * sl / ga.slices);
    t1 = static_cast<int>(tiles * (sl + 1) / ga.slices);
  };
  bool a_exhausted = false;   // meaningful in thread 0
  int kind, idx = 0;
  if (threadIdx.x == 0) {
    int i2 = 0;
    s_sel[0] = fused_select_unit(f, a_exhausted, true, i2);
    s_sel[1] = i2;
  }
  __syncthreads();
  kind = s_sel[0];
  idx = s_sel[1];
  __syncthreads();
  bool first_in_flight = false;   // the first tile of the coming A-unit has been prefetched
  int n_ev = 0;
  while (kind != kUnitDone) {
    int nkind = kUnitNone, nidx = 0;
    bool have_next = false;
    unsigned long long tr_t0 = 0, tr_wait = 0, tr_b0 = 0, tr_b1 = 0;
    if (f.trace && threadIdx.x == 0) tr_t0 = gtime();
    if (kind == kUnitA) {
      // ------------------------------------------------------------------ A-unit
      int g, img0, t0, t1;
      unit_range(idx, g, img0, t0, t1);
      uint32_t* g_hist = a.hist + static_cast<size_t>(g) * C * nbs;
      int img = img0 + t0 / a.tiles_per_image;
      int tile = t0 - (img - img0) * a.tiles_per_image;
      int p4 = tile * kThreadsG + threadIdx.x;
      bool valid = (t0 < t1) && (p4 < HW4);
      if (!first_in_flight && valid) prefetch(img, p4);
      first_in_flight = false;
      asm volatile("cp.async.wait_group 0;\n" ::: "memory");
      int run_lbl = 0;
      unsigned run_cnt = 0;
      for (int t = t0; t < t1; ++t) {
        int nimg = img, ntile = tile + 1;
        bool has_next = true;
        if (t + 1 < t1) {
          if (ntile == a.tiles_per_image) {
            ntile = 0;
            ++nimg;
          }
        } else {
          // last tile: pick the next unit now, so that the first tile of a following A-unit is in flight during it
          if (threadIdx.x == 0) {
            int i2 = 0;
            s_sel[0] = fused_select_unit(f, a_exhausted, false, i2);
            s_sel[1] = i2;
          }
          __syncthreads();
          nkind = s_sel[0];
          nidx = s_sel[1];
          have_next = true;
          has_next = false;
          if (nkind == kUnitA) {
            int ng, nimg0, nt0, nt1;
            unit_range(nidx, ng, nimg0, nt0, nt1);
            has_next = nt0 < nt1;
            nimg = nimg0 + nt0 / a.tiles_per_image;
            ntile = nt0 - (nimg - nimg0) * a.tiles_per_image;
            first_in_flight = true;   // uniform: every thread with a valid pixel prefetches below
          }
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
        float guard = 0.f;
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
          asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;\n" ::"l"(reinterpret_cast<float4*>(a.conf) + o4),
                       "f"(cf[0]), "f"(cf[1]), "f"(cf[2]), "f"(cf[3]), "l"(pol_last) : "memory");
          const unsigned lw = static_cast<unsigned>(lb[0]) | (static_cast<unsigned>(lb[1]) << 8) |
                              (static_cast<unsigned>(lb[2]) << 16) | (static_cast<unsigned>(lb[3]) << 24);
          asm volatile("st.global.L2::cache_hint.b32 [%0], %1, %2;\n" ::"l"(reinterpret_cast<unsigned*>(a.label) + o4), "r"(lw),
                       "l"(pol_last) : "memory");
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
              const int k = bin - hi0;
              const unsigned sh = (k & 1) * 16;
              tab16_add(s_tab + l * words + (k >> 1), sh, g_hist + static_cast<size_t>(l) * nbs + bin);
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
      if (!have_next) {   // empty unit (t0 == t1): nothing was selected inside the loop
        nkind = kUnitNone;
      }
      // flush the shared table, then report the unit
      if (run_cnt) atomicAdd(s_top + run_lbl, run_cnt);
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
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) {
        const unsigned done = atomicAdd(f.ws + 4 + g, 1u);
        s_closer = (done == static_cast<unsigned>(ga.slices) - 1) ? 1 : 0;
      }
      __syncthreads();
      if (s_closer) {
        // -------------------------------------------------------------- close group g (phase B for one group)
        if (f.trace && threadIdx.x == 0) tr_b0 = gtime();
        if (threadIdx.x == 0) {
          unsigned spins = 0;
          while (ld_acquire_u32(f.ws + 2) != static_cast<unsigned>(g)) {
            __nanosleep(100);
            if (++spins > kSpinLimit) {
              atomicOr(f.error_flag, 4);
              break;
            }
          }
        }
        __threadfence();
        __syncthreads();
        // The group's C histogram rows (final: every A-unit of the group has been flushed and fenced) are pulled
        // into the staging buffer a batch at a time with coalesced cp.async, all in flight at once -- one L2 round
        // trip per batch even while 147 CTAs saturate the memory system -- then one warp per row builds the
        // inclusive prefix in shared memory and runs the threshold step on it.  (Scanning the rows in place in
        // global memory costs ~90 us per group under load: the chain then runs slower than the groups arrive.)
        // The staging buffer may hold the already prefetched first tile of this CTA's next A-unit: it is dropped
        // and fetched again.
        if (f.trace && threadIdx.x == 0) tr_wait = gtime();
        first_in_flight = false;
        const int warp = threadIdx.x >> 5, lane = lane_id();
        uint32_t* s_rows = reinterpret_cast<uint32_t*>(s_raw);
        const int rows_fit = max(1, static_cast<int>(sizeof(float4) * C * kThreadsG / (sizeof(uint32_t) * nbs)));
        const int rows_per_batch = min(rows_fit, kThreadsG / 32);
        const int vec_per_row = nbs / 4;
        for (int c0 = 0; c0 < C; c0 += rows_per_batch) {
          const int nr = min(rows_per_batch, C - c0);
          const uint4* src = reinterpret_cast<const uint4*>(a.hist + (static_cast<size_t>(g) * C + c0) * nbs);
          for (int i = threadIdx.x; i < nr * vec_per_row; i += kThreadsG) {
            const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(reinterpret_cast<uint4*>(s_rows) + i));
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src + i) : "memory");
          }
          asm volatile("cp.async.commit_group;\n" ::: "memory");
          asm volatile("cp.async.wait_group 0;\n" ::: "memory");
          __syncthreads();
          if (warp < nr) {
            const int c = c0 + warp;
            uint32_t* row = s_rows + static_cast<size_t>(warp) * nbs;
            uint32_t carry = 0;
            for (int base = 0; base < a.nb; base += 128) {
              const int i0 = base + lane * 4;
              uint4 q = (i0 < nbs) ? *reinterpret_cast<const uint4*>(row + i0) : make_uint4(0, 0, 0, 0);
              uint32_t vv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (i0 + k >= a.nb) vv[k] = 0;
              vv[1] += vv[0]; vv[2] += vv[1]; vv[3] += vv[2];
              uint32_t x = vv[3];
#pragma unroll
              for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += y;
              }
              const uint32_t off = carry + x - vv[3];
              carry += __shfl_sync(0xffffffffu, x, 31);
              if (i0 < nbs) *reinterpret_cast<uint4*>(row + i0) = make_uint4(vv[0] + off, vv[1] + off, vv[2] + off, vv[3] + off);
            }
            __syncwarp();
            const double thr = __ldcg(f.thr_state + c);
            float temp = 0.f;
            int err = 0;
            const WarpSearch search = {row, a.nb};
            const double nthr = ias_threshold_step(row, a.nb, a.key_lo, thr, f.alpha, f.beta, f.gamma, &temp, &err, search);
            if (lane == 0) {
              f.thr_groups[static_cast<size_t>(g) * C + c] = nthr;
              if (f.temp_groups) f.temp_groups[static_cast<size_t>(g) * C + c] = temp;
              f.thr_state[c] = nthr;
              if (err) atomicOr(f.error_flag, err);
            }
          }
          __syncthreads();
        }
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) st_release_u32(f.ws + 2, static_cast<unsigned>(g) + 1u);
        if (f.trace && threadIdx.x == 0) tr_b1 = gtime();
      }
    } else if (kind == kUnitC) {
      // ------------------------------------------------------------------ C-unit
      int g, img0, t0, t1;
      unit_range(idx, g, img0, t0, t1);
      for (int i = threadIdx.x; i < C * kThreadsG; i += kThreadsG) s_acc[i] = 0;
      if (threadIdx.x < 256)
        s_thr[threadIdx.x] = threadIdx.x < C ? __double2float_ru(__ldcg(f.thr_groups + static_cast<size_t>(g) * C + threadIdx.x))
                                             : INFINITY;
      __syncthreads();
      unsigned long long* my_acc = s_acc + threadIdx.x;
      auto flush_image = [&](int img_) {
        __syncthreads();
        for (int c = threadIdx.x >> 5; c < C; c += kThreadsG / 32) {
          long long n = 0;
          unsigned long long sm = 0;
#pragma unroll
          for (int k = 0; k < kThreadsG / 32; ++k) {
            const int i = c * kThreadsG + k * 32 + lane_id();
            const unsigned long long w = s_acc[i];
            n += static_cast<long long>(w >> 48);
            sm += w & 0xffffffffffffull;
            s_acc[i] = 0;
          }
          n = warp_sum(n);
          sm = static_cast<unsigned long long>(warp_sum(static_cast<long long>(sm)));
          if (lane_id() == 0 && n) {
            atomicAdd(f.counts + static_cast<size_t>(img_) * C + c, static_cast<unsigned long long>(n));
            atomicAdd(f.confsum + static_cast<size_t>(g) * C + c, sm << 1);
          }
        }
        __syncthreads();
      };
      constexpr int kQ = 8;   // tiles (quads per thread) in flight
      int cur_img = img0 + t0 / a.tiles_per_image;
      const bool can_discard = DISCARD && f.discard;
      for (int t = t0; t < t1;) {
        const int img = img0 + t / a.tiles_per_image;
        const int tile = t - (img - img0) * a.tiles_per_image;
        const int nq = min(min(kQ, t1 - t), a.tiles_per_image - tile);
        if (img != cur_img) {
          flush_image(cur_img);
          cur_img = img;
        }
        float4 cq[kQ];
        unsigned lq[kQ];
        bool ok[kQ];
#pragma unroll
        for (int q = 0; q < kQ; ++q) {
          const int p4 = (tile + q) * kThreadsG + threadIdx.x;
          ok[q] = (q < nq) && (p4 < HW4);
          if (ok[q]) {
            const size_t o4 = static_cast<size_t>(img) * HW4 + p4;
            cq[q] = __ldcg(reinterpret_cast<const float4*>(a.conf) + o4);
            lq[q] = __ldcg(reinterpret_cast<const unsigned*>(a.label) + o4);
          }
        }
#pragma unroll
        for (int q = 0; q < kQ; ++q) {
          if (ok[q]) {
            const int p4 = (tile + q) * kThreadsG + threadIdx.x;
            const size_t o4 = static_cast<size_t>(img) * HW4 + p4;
            const float cf[4] = {cq[q].x, cq[q].y, cq[q].z, cq[q].w};
            unsigned o = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int l = (lq[q] >> (8 * j)) & 0xff;
              const bool ign = cf[j] < s_thr[l];
              o |= static_cast<unsigned>(ign ? HIAST_IGNORE_LABEL : l) << (8 * j);
              if (!ign) {
                const unsigned vq = __float2uint_rz(cf[j] * 2147483648.0f);
                my_acc[l * kThreadsG] += static_cast<unsigned long long>(vq) + (1ull << 48);
              }
            }
            __stcs(reinterpret_cast<unsigned*>(f.plbl) + o4, o);
          }
          if (can_discard) {
            // every lane of the warp has consumed its part of the lines: drop them from L2 without a write-back
            const bool full = __all_sync(0xffffffffu, ok[q]);
            if (full) {
              const int p4 = (tile + q) * kThreadsG + threadIdx.x;
              const size_t o4 = static_cast<size_t>(img) * HW4 + p4;
              if ((lane_id() & 7) == 0)
                asm volatile("discard.global.L2 [%0], 128;\n" ::"l"(reinterpret_cast<const float4*>(a.conf) + o4) : "memory");
              if (lane_id() == 0)
                asm volatile("discard.global.L2 [%0], 128;\n" ::"l"(reinterpret_cast<const unsigned*>(a.label) + o4) : "memory");
            }
          }
        }
        t += nq;
      }
      flush_image(cur_img);
    }
    if (f.trace && threadIdx.x == 0 && n_ev < kTraceEvents) {
      unsigned long long* e = f.trace + (static_cast<size_t>(blockIdx.x) * kTraceEvents + n_ev) * 6;
      e[0] = (static_cast<unsigned long long>(kind) << 32) | static_cast<unsigned>(idx);
      e[1] = tr_t0;
      e[2] = gtime();
      e[3] = tr_b0;      // closer: start of the wait for the previous group
      e[4] = tr_wait;    // closer: wait over, threshold step starts
      e[5] = tr_b1;      // closer: group published
      ++n_ev;
    }
    // ---------------------------------------------------------------------- next unit
    if (!have_next || nkind == kUnitNone) {
      if (threadIdx.x == 0) {
        int i2 = 0;
        s_sel[0] = fused_select_unit(f, a_exhausted, true, i2);
        s_sel[1] = i2;
      }
      __syncthreads();
      nkind = s_sel[0];
      nidx = s_sel[1];
      first_in_flight = false;
    }
    __syncthreads();
    kind = nkind;
    idx = nidx;
  }
}

#endif  // HIAST_DEV_VARIANTS
}  // namespace hiast

using namespace hiast;

namespace hiast {
unsigned long long* g_fused_trace = nullptr;
}
extern "C" int hiast_debug_set_fused_trace(void* dev_buffer) {
  hiast::g_fused_trace = static_cast<unsigned long long*>(dev_buffer);
  return HIAST_OK;
}

extern "C" size_t hiast_ias_fused_workspace_bytes(int n_images, int group_size) {
  if (n_images < 0 || group_size < 1) return 0;
  const size_t g = static_cast<size_t>((n_images + group_size - 1) / group_size);
  return sizeof(unsigned) * (4 + g);
}

#ifdef HIAST_DEV_VARIANTS
namespace hiast {
template <int C>
int launch_fused(FusedArgs f, int groups_in_flight, cudaStream_t st) {
  constexpr size_t kStage = sizeof(float4) * C * kThreadsG;
  constexpr size_t kBudget = 227 * 1024 - 2048;
  PhaseAArgs& a = f.ga.a;
  const int top = a.nb - 1;
  int words = static_cast<int>((kBudget - kStage) / (sizeof(uint32_t) * C));
  words = std::min(words, (top + 1) / 2);
  f.ga.words = words;
  f.ga.hi0 = std::max(top - 2 * words, 0);
  a.tiles_per_image = static_cast<int>((a.HW / 4 + kThreadsG - 1) / kThreadsG);
  a.n_tiles = static_cast<long long>(a.tiles_per_image) * a.n_images;
  const int sms = sm_count();
  const long long tiles_per_group = static_cast<long long>(a.tiles_per_image) * a.group_size;
  int slices = std::max(1, sms / std::max(1, groups_in_flight));
  slices = static_cast<int>(std::max<long long>(1, std::min<long long>(slices, tiles_per_group / 4)));
  f.ga.slices = slices;
  f.ga.n_units = f.n_groups * slices;
  HIAST_TRY(ensure_dyn_smem(k_ias_fused<C, 1>, kBudget));
  const size_t smem = kStage + sizeof(uint32_t) * C * words;
  const int grid = std::min(sms, f.ga.n_units);
  k_ias_fused<C, 1><<<grid, kThreadsG, smem, st>>>(f);
  HIAST_CHECK_LAUNCH();
  return HIAST_OK;
}
}  // namespace hiast
#endif  // HIAST_DEV_VARIANTS

extern "C" int hiast_ias_fused_window(const float* logits, int n_images, int C, int H, int W, int group_size, int key_lo,
                                      double alpha, double beta, double gamma, float* conf_scratch, uint8_t* label_scratch,
                                      uint32_t* hist, double* thr_state, double* thr_groups, float* temp_groups,
                                      uint8_t* plbl, int64_t* counts, uint64_t* confsum, int* error_flag, void* workspace,
                                      size_t workspace_bytes, int flags, void* stream) {
  using namespace hiast;
  if (!logits || !conf_scratch || !label_scratch || !hist || !thr_state || !thr_groups || !plbl || !counts || !confsum ||
      !error_flag || !workspace)
    return HIAST_ERR_INVALID_ARG;
  if (n_images < 0 || C < 1 || C > HIAST_MAX_CLASSES || H < 1 || W < 1 || group_size < 1) return HIAST_ERR_INVALID_ARG;
  if (key_lo < 0 || key_lo > HIAST_KEY_ONE) return HIAST_ERR_INVALID_ARG;
  if (workspace_bytes < hiast_ias_fused_workspace_bytes(n_images, group_size)) return HIAST_ERR_WORKSPACE;
  if (n_images == 0) return HIAST_OK;
  const int64_t HW = static_cast<int64_t>(H) * W;
  const bool aligned = (HW % 4 == 0) && (reinterpret_cast<uintptr_t>(logits) % 16 == 0) &&
                       (reinterpret_cast<uintptr_t>(conf_scratch) % 16 == 0) && (reinterpret_cast<uintptr_t>(label_scratch) % 4 == 0) &&
                       (reinterpret_cast<uintptr_t>(plbl) % 4 == 0);
  if (!aligned || (C != 19 && C != 16)) return HIAST_ERR_UNSUPPORTED;   // callers fall back to the three-kernel path
#ifndef HIAST_DEV_VARIANTS
  // measured 40 % slower than the three kernels on B200: compiled only into the development build
  (void)alpha; (void)beta; (void)gamma; (void)temp_groups; (void)flags; (void)stream;
  return HIAST_ERR_UNSUPPORTED;
#else
  cudaStream_t st = as_stream(stream);
  const int n_groups = (n_images + group_size - 1) / group_size;
  HIAST_CUDA_TRY(cudaMemsetAsync(hist, 0, hiast_ias_hist_bytes(n_groups, C, key_lo), st));
  HIAST_CUDA_TRY(cudaMemsetAsync(counts, 0, sizeof(int64_t) * n_images * C, st));
  HIAST_CUDA_TRY(cudaMemsetAsync(confsum, 0, sizeof(uint64_t) * n_groups * C, st));
  HIAST_CUDA_TRY(cudaMemsetAsync(workspace, 0, hiast_ias_fused_workspace_bytes(n_images, group_size), st));
  FusedArgs f;
  PhaseAArgs& a = f.ga.a;
  a.logits = logits; a.conf = conf_scratch; a.label = label_scratch; a.hist = hist;
  a.n_images = n_images; a.C = C; a.HW = HW;
  a.group_size = group_size; a.key_lo = key_lo; a.nb = HIAST_KEY_ONE - key_lo + 1;
  a.sched = nullptr;
  f.alpha = alpha; f.beta = beta; f.gamma = gamma;
  f.thr_state = thr_state; f.thr_groups = thr_groups; f.temp_groups = temp_groups;
  f.plbl = plbl;
  f.counts = reinterpret_cast<unsigned long long*>(counts);
  f.confsum = reinterpret_cast<unsigned long long*>(confsum);
  f.error_flag = error_flag;
  f.ws = static_cast<unsigned*>(workspace);
  f.n_groups = n_groups;
  // lines can be discarded whole only if every image plane starts on a 128-byte line in both spill arrays
  const bool lines = (HW % 128 == 0) && (reinterpret_cast<uintptr_t>(conf_scratch) % 128 == 0) &&
                     (reinterpret_cast<uintptr_t>(label_scratch) % 128 == 0);
  f.discard = (lines && !(flags & 1)) ? 1 : 0;
  f.trace = g_fused_trace;
  int gif = (flags >> 4) & 0xf;
  if (gif == 0) gif = 2;
  if (C == 19) return launch_fused<19>(f, gif, st);
  return launch_fused<16>(f, gif, st);
#endif
}

