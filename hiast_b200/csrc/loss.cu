// (3) Region-adaptive regularisation + consistency losses, fused forward and backward (sm_100a).
//
// Reference: sseg/models/segmentors/self_training_segmentor.py:30-53 (compute_loss),
// :128-137 (build_region_weight), :140-150 (_entropy), :153-163 (_kld);
// sseg/models/modules/losses.py:32-36 (ce), :39-61 (soft_ce / SoftCELoss), :75-89
// (compute_loss_by_selected_pixel).  Closed forms: SURVEY.md Appendix A.6.
//
// The reference launches ~25 full-tensor kernels forward (three [B,C,H,W] weight tensors, three
// log_softmax, a softmax, masked products, bool-index compactions for the divisors) and about as
// many backward.  Here: ONE forward pass over (z, t, plbl) computes the log-softmax of a pixel once
// and feeds all four masked reductions + the three integer divisors; ONE backward pass recomputes
// it and writes grad_z.  Both are pure HBM streams:
//   forward  reads 4C (z) + 4C (t) + 1|8 (plbl) B/px
//   backward reads the same and writes 4C B/px
// The log-softmax uses the same fp32 op sequence as ATen's spatial kernel (sequential max, sum of
// expf(z-max) in channel order, z - max - logf(sum)) so that the data-dependent SoftCE divisor
// (#non-zero products, losses.py:89) is reproduced exactly.
#include <math.h>

#include <algorithm>

#include "common.cuh"
#include "packed_math.cuh"

namespace hiast {

constexpr int kThreadsL = 256;
constexpr int kPxL = 2;  // pixels per thread (one 64-bit load per channel)

struct LossArgs {
  const float* z;
  const float* t;
  const void* plbl;
  int plbl_bytes;
  int B;
  int C;
  int64_t HW;
  int region;
  int terms;   // HIAST_TERM_* | HIAST_CST_* (kind of the consistency term)
};

__device__ __forceinline__ int cst_kind(int terms) { return terms & HIAST_CST_SOFTCE_LOGITS; }

__device__ __forceinline__ int load_label(const void* p, int bytes, size_t i) {
  if (bytes == 1) return static_cast<const uint8_t*>(p)[i];
  const long long v = static_cast<const long long*>(p)[i];
  return (v < 0 || v > 255) ? 256 : static_cast<int>(v);  // 256 = out of range, not ignore
}

// labels of the pixel pair at element position lp (even): issued one iteration ahead so that the load is never waited for
__device__ __forceinline__ void load_label_pair(const void* p, int bytes, size_t lp, int& ya, int& yb) {
  ya = load_label(p, bytes, lp);
  yb = load_label(p, bytes, lp + 1);
}

__device__ __forceinline__ bool in_region(int region, bool ignored) {
  return region == HIAST_REGION_ALL || (region == HIAST_REGION_IGNORED ? ignored : !ignored);
}

struct PixelSums {
  double ce, kld, ent, cst;
  long long n_conf, n_ign, n_nz;
};

// log-softmax pieces of one pixel held in registers
template <int C>
struct PixelLS {
  float m, logs, inv_s, sum;
  __device__ __forceinline__ void init(const float (&z)[C]) {
    m = z[0];
#pragma unroll
    for (int c = 1; c < C; ++c) m = fmaxf(m, z[c]);
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) s += expf(z[c] - m);
    logs = logf(s);
    inv_s = 1.0f / s;
    sum = s;
  }
  __device__ __forceinline__ float prob(float zc) const { return __fdiv_rn(expf(zc - m), sum); }  // ATen softmax
  __device__ __forceinline__ float logp(float zc) const { return (zc - m) - logs; }
  __device__ __forceinline__ float p(float zc) const { return expf(zc - m) * inv_s; }
};

template <int C>
__device__ __forceinline__ void pixel_forward(const float (&z)[C], const float (&t)[C], int y, int region, int terms,
                                              PixelSums& acc) {
  PixelLS<C> ls;
  ls.init(z);
  const bool ignored = (y == HIAST_IGNORE_LABEL);
  if (ignored) acc.n_ign += 1; else acc.n_conf += 1;
  if (!ignored) {
    if (terms & HIAST_TERM_CE) {
      float lp = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) lp = (c == y) ? ls.logp(z[c]) : lp;
      acc.ce += static_cast<double>(-lp);
    }
    if (terms & HIAST_TERM_KLD) {
      float sl = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) sl += ls.logp(z[c]);
      acc.kld += static_cast<double>(-sl * (1.0f / C));
    }
  } else if (terms & HIAST_TERM_ENT) {
    float h = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) h += ls.p(z[c]) * ls.logp(z[c]);
    acc.ent += static_cast<double>(-h);
  }
  if ((terms & HIAST_TERM_CST) && in_region(region, ignored)) {
    float sc = 0.f;
    int nz = 0;
    const int kind = cst_kind(terms);
    if (kind == HIAST_CST_KLDIV) {
      // losses.py:16-23: KLDivLoss(reduction='none')(log_softmax(z), softmax(t)) = xlogy(tp,tp) - tp*logp
      PixelLS<C> lt;
      lt.init(t);
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float tp = lt.prob(t[c]);
        const float elem = (tp == 0.f ? 0.f : __fmul_rn(tp, logf(tp))) - __fmul_rn(tp, ls.logp(z[c]));
        nz += (elem != 0.f);
        sc += elem;
      }
    } else if (kind == HIAST_CST_MSE) {
#pragma unroll
      for (int c = 0; c < C; ++c) {      // losses.py:9-13: (z - t)^2
        const float d = z[c] - t[c];
        const float elem = __fmul_rn(d, d);
        nz += (elem != 0.f);
        sc += elem;
      }
    } else {
      PixelLS<C> lt;
      if (kind == HIAST_CST_SOFTCE_LOGITS) lt.init(t);   // teacher logits: softmax fused here (8f rank 4)
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float tc = (kind == HIAST_CST_SOFTCE_LOGITS) ? lt.prob(t[c]) : t[c];
        const float prod = __fmul_rn(-ls.logp(z[c]), tc);
        nz += (prod != 0.f);
        sc += prod;
      }
    }
    acc.cst += static_cast<double>(sc);
    acc.n_nz += nz;
  }
}

// Backward of one pixel with a small register footprint: `z` is overwritten by its log-softmax (the gradient
// only needs logp and p = exp(logp)), KLDIV targets are turned into probabilities in place, and the per-pixel
// scalars are kept in this struct; grad(c) is then evaluated channel by channel right before the store.
template <int C>
struct PixelBwd {
  float a, kb, ce, ent, h, cst, T, mls;  // mls = max + log-sum (to rebuild z for MSE)
  int y, kind;
  bool ignored, has_cst;

  __device__ __forceinline__ void init(float (&z)[C], float (&t)[C], int y_, int region, int terms, const float (&sc)[4]) {
    PixelLS<C> ls;
    ls.init(z);
#pragma unroll
    for (int c = 0; c < C; ++c) z[c] = ls.logp(z[c]);
    mls = ls.m + ls.logs;
    y = y_;
    ignored = (y == HIAST_IGNORE_LABEL);
    kind = cst_kind(terms);
    has_cst = (terms & HIAST_TERM_CST) && in_region(region, ignored);
    ce = (!ignored && (terms & HIAST_TERM_CE)) ? sc[0] : 0.f;
    const float kld = (!ignored && (terms & HIAST_TERM_KLD)) ? sc[1] : 0.f;
    a = ce + kld;
    kb = kld * (1.0f / C);
    ent = (ignored && (terms & HIAST_TERM_ENT)) ? sc[2] : 0.f;
    h = 0.f;
    if (ent != 0.f || (ignored && (terms & HIAST_TERM_ENT))) {
#pragma unroll
      for (int c = 0; c < C; ++c) h += expf(z[c]) * z[c];  // = -H
    }
    cst = has_cst ? sc[3] : 0.f;
    T = 0.f;
    if (has_cst && kind != HIAST_CST_MSE) {
      if (kind == HIAST_CST_KLDIV || kind == HIAST_CST_SOFTCE_LOGITS) {   // targets arrive as logits
        PixelLS<C> lt;
        lt.init(t);
#pragma unroll
        for (int c = 0; c < C; ++c) t[c] = lt.prob(t[c]);
      }
#pragma unroll
      for (int c = 0; c < C; ++c) T += t[c];
    }
  }

  __device__ __forceinline__ float grad(int c, float lp, float tc) const {
    const float p = expf(lp);
    float g;
    if (!ignored) g = a * p - kb - ((c == y) ? ce : 0.f);
    else g = -ent * p * (lp - h);
    if (has_cst) g += (kind == HIAST_CST_MSE) ? cst * 2.0f * ((lp + mls) - tc) : cst * (p * T - tc);
    return g;
  }
};

struct Partial {
  double s[4];
  long long n[3];
  long long pad;
};

__device__ __forceinline__ void block_reduce_store(const PixelSums& acc, Partial* out) {
  __shared__ Partial s_part[kThreadsL / 32];
  double s0 = warp_sum(acc.ce), s1 = warp_sum(acc.kld), s2 = warp_sum(acc.ent), s3 = warp_sum(acc.cst);
  long long n0 = warp_sum(acc.n_conf), n1 = warp_sum(acc.n_ign), n2 = warp_sum(acc.n_nz);
  if (lane_id() == 0) {
    Partial& p = s_part[threadIdx.x >> 5];
    p.s[0] = s0; p.s[1] = s1; p.s[2] = s2; p.s[3] = s3;
    p.n[0] = n0; p.n[1] = n1; p.n[2] = n2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    Partial r = s_part[0];
    for (int w = 1; w < kThreadsL / 32; ++w) {
      for (int k = 0; k < 4; ++k) r.s[k] += s_part[w].s[k];
      for (int k = 0; k < 3; ++k) r.n[k] += s_part[w].n[k];
    }
    r.pad = 0;
    out[blockIdx.x] = r;
  }
}

// Vector path: HW even, C compile-time.  Grid-stride over pairs of pixels, software-pipelined with cp.async:
// every thread owns 2C x 8 bytes of shared memory (its pixel pair's z and t columns); at the top of an iteration it
// pulls them into registers, immediately re-issues the 8-byte cp.async of ITS next pixel pair into the same
// slots, and does the math while they are in flight (LDGSTS is tracked by async groups, not by the register
// scoreboard the MUFU results of the math use -- same reasoning as k_softmax_hist_sp in ias_phase_a.cu).
template <int C>
struct LossStage {
  float2* my;          // this thread's column: [2C][kThreadsL] float2, z channels then t channels
  unsigned my_u32;
  const LossArgs& a;
  int64_t HW2;
  bool need_t;

  __device__ __forceinline__ LossStage(float2* smem, const LossArgs& args)
      : my(smem + threadIdx.x), my_u32(static_cast<unsigned>(__cvta_generic_to_shared(smem + threadIdx.x))), a(args),
        HW2(args.HW / kPxL), need_t((args.terms & HIAST_TERM_CST) != 0) {}

  __device__ __forceinline__ void prefetch(long long i) const {
    const int b = static_cast<int>(i / HW2);
    const int64_t p2 = i - static_cast<long long>(b) * HW2;
    const float2* zs = reinterpret_cast<const float2*>(a.z + static_cast<size_t>(b) * C * a.HW) + p2;
#pragma unroll
    for (int c = 0; c < C; ++c)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(my_u32 + c * kThreadsL * 8),
                   "l"(zs + static_cast<size_t>(c) * HW2) : "memory");
    if (need_t) {
      const float2* ts = reinterpret_cast<const float2*>(a.t + static_cast<size_t>(b) * C * a.HW) + p2;
#pragma unroll
      for (int c = 0; c < C; ++c)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(my_u32 + (C + c) * kThreadsL * 8),
                     "l"(ts + static_cast<size_t>(c) * HW2) : "memory");
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  }
  __device__ __forceinline__ void wait() const { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

  // Split pipeline of the packed kernels: z and t travel in separate cp.async groups so that t never has to live in
  // registers -- it is read from the staging slots channel by channel after the first (z-only) pass and its slots
  // are refilled after the last pass.  wait_older() = every group but the most recent one has landed.
  __device__ __forceinline__ void prefetch_z(long long i) const {
    const int b = static_cast<int>(i / HW2);
    const int64_t p2 = i - static_cast<long long>(b) * HW2;
    const float2* zs = reinterpret_cast<const float2*>(a.z + static_cast<size_t>(b) * C * a.HW) + p2;
#pragma unroll
    for (int c = 0; c < C; ++c)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(my_u32 + c * kThreadsL * 8),
                   "l"(zs + static_cast<size_t>(c) * HW2) : "memory");
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  }
  __device__ __forceinline__ void prefetch_t(long long i) const {
    if (need_t) {
      const int b = static_cast<int>(i / HW2);
      const int64_t p2 = i - static_cast<long long>(b) * HW2;
      const float2* ts = reinterpret_cast<const float2*>(a.t + static_cast<size_t>(b) * C * a.HW) + p2;
#pragma unroll
      for (int c = 0; c < C; ++c)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(my_u32 + (C + c) * kThreadsL * 8),
                     "l"(ts + static_cast<size_t>(c) * HW2) : "memory");
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");   // an empty group keeps the group count uniform
  }
  // The same with the pair's position (image b, pair index p2 inside the image) carried along by the caller: the packed kernels
  // advance it incrementally (PairPos) instead of dividing a 64-bit index three times per iteration.
  __device__ __forceinline__ void prefetch_z_at(int b, int64_t p2) const {
    const float2* zs = reinterpret_cast<const float2*>(a.z + static_cast<size_t>(b) * C * a.HW) + p2;
#pragma unroll
    for (int c = 0; c < C; ++c)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(my_u32 + c * kThreadsL * 8),
                   "l"(zs + static_cast<size_t>(c) * HW2) : "memory");
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  }
  __device__ __forceinline__ void prefetch_t_at(int b, int64_t p2) const {
    if (need_t) {
      const float2* ts = reinterpret_cast<const float2*>(a.t + static_cast<size_t>(b) * C * a.HW) + p2;
#pragma unroll
      for (int c = 0; c < C; ++c)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(my_u32 + (C + c) * kThreadsL * 8),
                     "l"(ts + static_cast<size_t>(c) * HW2) : "memory");
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");   // an empty group keeps the group count uniform
  }
  __device__ __forceinline__ void wait_older() const { asm volatile("cp.async.wait_group 1;\n" ::: "memory"); }
  // z of the labels' channels, read from the staging column by index (two LDS instead of a 19-deep select chain per lane);
  // labels outside [0, C) -- ignore, out of range -- read channel 0, the callers discard the value.  Folded into the guard.
  __device__ __forceinline__ float load_zy(int ya, int yb, float& zya, float& zyb) const {
    zya = my[(static_cast<unsigned>(ya) < static_cast<unsigned>(C) ? ya : 0) * kThreadsL].x;
    zyb = my[(static_cast<unsigned>(yb) < static_cast<unsigned>(C) ? yb : 0) * kThreadsL].y;
    return fmaxf(zya, zyb);
  }
  __device__ __forceinline__ float load_z(float (&z)[kPxL][C]) const {
    float guard = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float2 q = my[c * kThreadsL];
      z[0][c] = q.x; z[1][c] = q.y;
      guard = fmaxf(guard, q.x);
    }
    return guard;
  }
  __device__ __forceinline__ pk::u64 t2(int c) const {
    if (!need_t) return 0ull;
    const float2 q = my[(C + c) * kThreadsL];
    return pk::pack(q.x, q.y);
  }

  // registers <- shared; returns a value that depends on every load so that the refill can be ordered after it
  __device__ __forceinline__ float load(float (&z)[kPxL][C], float (&t)[kPxL][C]) const {
    float guard = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float2 q = my[c * kThreadsL];
      z[0][c] = q.x; z[1][c] = q.y;
      guard = fmaxf(guard, q.x);
    }
    if (need_t) {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float2 q = my[(C + c) * kThreadsL];
        t[0][c] = q.x; t[1][c] = q.y;
        guard = fmaxf(guard, q.x);
      }
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) t[0][c] = t[1][c] = 0.f;
    }
    return guard;
  }
};

template <int C>
__global__ void __launch_bounds__(kThreadsL, 2) k_loss_fwd(LossArgs a, Partial* __restrict__ partials) {
  extern __shared__ __align__(128) float2 s_loss_stage[];
  const LossStage<C> st(s_loss_stage, a);
  const long long total = static_cast<long long>(a.B) * st.HW2;
  const long long stride = static_cast<long long>(gridDim.x) * kThreadsL;
  PixelSums acc = {0.0, 0.0, 0.0, 0.0, 0, 0, 0};
  long long i = static_cast<long long>(blockIdx.x) * kThreadsL + threadIdx.x;
  if (i < total) st.prefetch(i);
  st.wait();
  for (; i < total; i += stride) {
    float z[kPxL][C], t[kPxL][C];
    const float guard = st.load(z, t);
    if (i + stride < total && guard == guard) st.prefetch(i + stride);
    const int b = static_cast<int>(i / st.HW2);
    const size_t lp = static_cast<size_t>(b) * a.HW + (i - static_cast<long long>(b) * st.HW2) * kPxL;
#pragma unroll
    for (int j = 0; j < kPxL; ++j) pixel_forward<C>(z[j], t[j], load_label(a.plbl, a.plbl_bytes, lp + j), a.region, a.terms, acc);
    st.wait();
  }
  block_reduce_store(acc, partials);
}

template <int C>
__global__ void __launch_bounds__(kThreadsL, 2) k_loss_bwd(LossArgs a, const float* __restrict__ scales,
                                                           float* __restrict__ grad) {
  extern __shared__ __align__(128) float2 s_loss_stage[];
  const LossStage<C> st(s_loss_stage, a);
  const long long total = static_cast<long long>(a.B) * st.HW2;
  const long long stride = static_cast<long long>(gridDim.x) * kThreadsL;
  const float sc[4] = {scales[0], scales[1], scales[2], scales[3]};
  // The reference divides a masked sum by an element count: an empty region makes its scale inf/NaN and
  // autograd then yields NaN for EVERY element (inf * 0).  `poison` is 0 unless some enabled scale is
  // non-finite, in which case it is NaN.
  float poison = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (a.terms & (1 << k)) poison += 0.f * sc[k];
  long long i = static_cast<long long>(blockIdx.x) * kThreadsL + threadIdx.x;
  if (i < total) st.prefetch(i);
  st.wait();
  for (; i < total; i += stride) {
    float z[kPxL][C], t[kPxL][C];
    const float guard = st.load(z, t);
    if (i + stride < total && guard == guard) st.prefetch(i + stride);
    const int b = static_cast<int>(i / st.HW2);
    const int64_t p2 = i - static_cast<long long>(b) * st.HW2;
    const size_t lp = static_cast<size_t>(b) * a.HW + p2 * kPxL;
    PixelBwd<C> px0, px1;
    px0.init(z[0], t[0], load_label(a.plbl, a.plbl_bytes, lp), a.region, a.terms, sc);
    px1.init(z[1], t[1], load_label(a.plbl, a.plbl_bytes, lp + 1), a.region, a.terms, sc);
    float2* gs = reinterpret_cast<float2*>(grad + static_cast<size_t>(b) * C * a.HW) + p2;
#pragma unroll
    for (int c = 0; c < C; ++c)
      __stcs(gs + static_cast<size_t>(c) * st.HW2,
             make_float2(px0.grad(c, z[0][c], t[0][c]) + poison, px1.grad(c, z[1][c], t[1][c]) + poison));
    st.wait();
  }
}

// ---- packed-pair kernels (consistency kind SoftCE, the HIAST configuration) --------------------------------------------
// The scalar kernels above spend 1232 (backward) instructions per pixel -- up to three accurate expf per channel -- and
// are issue-bound at 60 % of the HBM roofline.  Here the two pixels of a thread go through the f32x2 forms (packed_math.cuh):
// ONE exact exponential per channel and pair (the softmax sum); the entropy inner product, the sum of log-probabilities and the
// label's log-probability fall out of that loop, and the gradient pass takes p from ex2.approx (round 2: 108 -> ~60 issued
// instructions per channel and pair).  log-softmax keeps ATen's operation sequence ((z - m) - log(sum), sum in channel order)
// so that the data-dependent SoftCE divisor #(prod != 0) stays exact, and the count of non-zero products is C unless
// min_c |prod_c| == 0 (then it is recounted).

// expf for the gradient pass: ex2.approx of d * log2(e) -- three issue slots per pair instead of the dozen of the exact
// pair exponential.  |relative error| <= 2^-22 + 1.45 |d| 2^-24 (and whatever is lost below e^-87 is below 1e-37 absolute),
// far inside the 1e-5 the gradients are held to.  The SUM of exponentials -- which decides log-softmax and through it the
// exact SoftCE divisor #(prod != 0) -- keeps pk::exp2x.
__device__ __forceinline__ pk::u64 exp_fast2(pk::u64 d2) {
  float fa, fb;
  pk::unpack(pk::mul2(d2, pk::splat(1.4426950408889634f)), fa, fb);
  return pk::pack(pk::ex2_ftz(fa), pk::ex2_ftz(fb));
}

// lanes of a packed pair kept (mask all ones) or replaced by +0.0f (mask 0)
__device__ __forceinline__ pk::u64 lane_mask(bool a, bool b) {
  return (a ? 0x00000000ffffffffull : 0ull) | (b ? 0xffffffff00000000ull : 0ull);
}

// Log-softmax pieces of a pixel pair.  The softmax pass computes e_c = exp(z_c - m) once per channel anyway; the scalars
// the forward sums need fall out of the same loop instead of a second exponential per channel:
//   sum_c p_c logp_c (= -entropy)   = (sum_c e_c d_c) / s - log s          d_c = z_c - m
//   sum_c logp_c     (KL to uniform) = sum_c d_c - C log s
//   logp_y           (cross entropy) = d_y - log s                          (the very bits of (z_y - m) - log s)
// (kernels that do not use one of them pay nothing: the dead chains are eliminated.)
template <int C>
struct PairLS {
  float ma, mb, sa, sb;
  pk::u64 negm, logs2, inv2, h2, sl2, lpy2;
  // (zya, zyb) = z of the labels' channels (LossStage::load_zy); a label outside [0, C) contributes logp_y = 0 like the select
  // chain `c == y ? lp : 0` it replaces
  __device__ __forceinline__ void init(const float (&za)[C], const float (&zb)[C], int ya, int yb, float zya, float zyb) {
    ma = za[0];
    mb = zb[0];
#pragma unroll
    for (int c = 1; c < C; ++c) {
      ma = fmaxf(ma, za[c]);
      mb = fmaxf(mb, zb[c]);
    }
    negm = pk::pack(-ma, -mb);
    pk::u64 s2 = 0, ed2 = 0, sd2 = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const pk::u64 d2 = pk::add2(pk::pack(za[c], zb[c]), negm);
      const pk::u64 e2 = pk::exp2x(d2);
      s2 = (c == 0) ? e2 : pk::add2(s2, e2);
      ed2 = (c == 0) ? pk::mul2(e2, d2) : pk::fma2(e2, d2, ed2);
      sd2 = (c == 0) ? d2 : pk::add2(sd2, d2);
    }
    pk::unpack(s2, sa, sb);
    logs2 = pk::pack(logf(sa), logf(sb));
    inv2 = pk::pack(1.0f / sa, 1.0f / sb);
    h2 = pk::sub2(pk::mul2(ed2, inv2), logs2);
    sl2 = pk::fma2(pk::splat(-static_cast<float>(C)), logs2, sd2);
    lpy2 = pk::sub2(pk::add2(pk::pack(zya, zyb), negm), logs2) &
           lane_mask(static_cast<unsigned>(ya) < static_cast<unsigned>(C), static_cast<unsigned>(yb) < static_cast<unsigned>(C));
  }
  __device__ __forceinline__ pk::u64 d(float a, float b) const { return pk::add2(pk::pack(a, b), negm); }
};

// rare path: some -logp * t of the pixel at (b, p) is exactly zero; count the non-zero ones from global memory
// (the registers and the staging slots have moved on)
template <int C>
__device__ __noinline__ int recount_nonzero(const LossArgs& a, int b, int64_t p, float m, float logs) {
  int nz = 0;
  for (int c = 0; c < C; ++c) {
    const size_t i = (static_cast<size_t>(b) * C + c) * a.HW + p;
    nz += (__fmul_rn(-((a.z[i] - m) - logs), a.t[i]) != 0.f);
  }
  return nz;
}

// SoftCE pass over the channels of a pair (the only forward term that needs one): sum_c logp_c t_c with log-softmax in ATen's
// operation sequence ((z - m) - log(sum)), the smallest |product| per pixel (the count of non-zero products is C unless it is
// 0) and the teacher mass T = sum_c t_c the gradient uses.
template <int C>
struct PairCst {
  pk::u64 sc2, T2;
  float mna, mnb;
  __device__ __forceinline__ void run(const float (&za)[C], const float (&zb)[C], const PairLS<C>& ls, const LossStage<C>& st) {
    sc2 = 0;
    T2 = 0;
    mna = INFINITY;
    mnb = INFINITY;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const pk::u64 lp2 = pk::sub2(ls.d(za[c], zb[c]), ls.logs2);
      const pk::u64 t2 = st.t2(c);
      T2 = (c == 0) ? t2 : pk::add2(T2, t2);
      const pk::u64 pr2 = pk::mul2(lp2, t2);   // = -(-lp * t), same magnitude and zero-ness
      sc2 = (c == 0) ? pr2 : pk::add2(sc2, pr2);
      float pa, pb;
      pk::unpack(pr2, pa, pb);
      mna = fminf(mna, fabsf(pa));
      mnb = fminf(mnb, fabsf(pb));
    }
  }
};

// One pixel's contributions to the forward sums; `add(k, v)` accumulates v into sum k (0 CE, 1 KLD, 2 ENT, 3 CST, 4 n_conf,
// 5 n_ign, 6 n_nz) -- registers in the forward kernel, shared-memory columns in the one-pass kernel.
template <int C, class Add>
__device__ __forceinline__ void pixel_sums(bool ign, int terms, int region, float lpy, float sl, float h, float scv, float mn,
                                           const LossArgs& a, int b, int64_t p, float m, float logs, Add add) {
  if (!ign) {
    if (terms & HIAST_TERM_CE) add(0, static_cast<double>(-lpy));
    if (terms & HIAST_TERM_KLD) add(1, static_cast<double>(-sl * (1.0f / C)));
    add(4, 1.0);
  } else {
    if (terms & HIAST_TERM_ENT) add(2, static_cast<double>(-h));
    add(5, 1.0);
  }
  if ((terms & HIAST_TERM_CST) && in_region(region, ign)) {
    add(3, static_cast<double>(-scv));
    add(6, static_cast<double>((mn > 0.f) ? C : recount_nonzero<C>(a, b, p, m, logs)));   // a zero (or NaN) product: the slow way
  }
}

template <int C, class Add>
__device__ __forceinline__ void pair_sums(const PairLS<C>& ls, const PairCst<C>& cs, bool iga, bool igb, const LossArgs& a,
                                          int b, int64_t p0, Add add) {
  float lpya, lpyb, sla, slb, ha, hb, sca, scb, logsa, logsb;
  pk::unpack(ls.lpy2, lpya, lpyb);
  pk::unpack(ls.sl2, sla, slb);
  pk::unpack(ls.h2, ha, hb);
  pk::unpack(cs.sc2, sca, scb);
  pk::unpack(ls.logs2, logsa, logsb);
  pixel_sums<C>(iga, a.terms, a.region, lpya, sla, ha, sca, cs.mna, a, b, p0, ls.ma, logsa, add);
  pixel_sums<C>(igb, a.terms, a.region, lpyb, slb, hb, scb, cs.mnb, a, b, p0 + 1, ls.mb, logsb, add);
}

// Coefficients of a pair's gradient:  g_c = A p_c + K - [c == y] ce - ENT p_c (logp_c - h) - CST (t_c - p_c T)
struct GradScales {
  float s_ce, s_kld, s_ent, s_cst;
  pk::u64 poison2;   // NaN iff an enabled scale is non-finite (empty region in the reference: every gradient is NaN)
  __device__ __forceinline__ void init(const float (&sc)[4], int terms) {
    float poison = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (terms & (1 << k)) poison += 0.f * sc[k];
    poison2 = pk::splat(poison);
    s_ce = (terms & HIAST_TERM_CE) ? sc[0] : 0.f;
    s_kld = (terms & HIAST_TERM_KLD) ? sc[1] : 0.f;
    s_ent = (terms & HIAST_TERM_ENT) ? sc[2] : 0.f;
    s_cst = (terms & HIAST_TERM_CST) ? sc[3] : 0.f;
  }
};

// Writes the gradient of a pair channel by channel right before the store; returns a value that depends on every read of
// the t slots (they may be refilled once it is known).
template <int C>
__device__ __forceinline__ float pair_gradient_store(const float (&za)[C], const float (&zb)[C], const PairLS<C>& ls,
                                                     const LossStage<C>& st, const GradScales& gs, int ya, int yb, bool iga,
                                                     bool igb, bool csa, bool csb, pk::u64 T2, float2* out, int64_t HW2) {
  const pk::u64 A2 = pk::pack(iga ? 0.f : gs.s_ce + gs.s_kld, igb ? 0.f : gs.s_ce + gs.s_kld);
  const pk::u64 K2 = pk::add2(pk::pack(iga ? 0.f : -(gs.s_kld * (1.0f / C)), igb ? 0.f : -(gs.s_kld * (1.0f / C))), gs.poison2);
  const pk::u64 NE2 = pk::pack(iga ? -gs.s_ent : 0.f, igb ? -gs.s_ent : 0.f);
  const pk::u64 NC2 = pk::pack(csa ? -gs.s_cst : 0.f, csb ? -gs.s_cst : 0.f);
  const float cea = iga ? 0.f : gs.s_ce, ceb = igb ? 0.f : gs.s_ce;
  const pk::u64 ign_mask = lane_mask(iga, igb);
  const pk::u64 NT2 = T2 ^ 0x8000000080000000ull;
  const pk::u64 hm2 = ls.h2 & ign_mask;
  float tguard = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const pk::u64 d2 = ls.d(za[c], zb[c]);
    const pk::u64 p2 = pk::mul2(exp_fast2(d2), ls.inv2);
    const pk::u64 lpm2 = pk::sub2(d2, ls.logs2) & ign_mask;       // confident lanes: p * (0 - 0) * 0
    const pk::u64 t2 = st.t2(c);
    pk::u64 g2 = pk::fma2(p2, A2, K2);
    g2 = pk::fma2(pk::mul2(p2, pk::sub2(lpm2, hm2)), NE2, g2);    // - ent p (lp - h)
    g2 = pk::fma2(pk::fma2(p2, NT2, t2), NC2, g2);                // - cst (t - p T)
    float ga, gb, t0, t1;
    pk::unpack(g2, ga, gb);
    pk::unpack(t2, t0, t1);
    tguard = fmaxf(tguard, t0);
    if (c == ya) ga -= cea;
    if (c == yb) gb -= ceb;
    __stcs(out + static_cast<size_t>(c) * HW2, make_float2(ga, gb));
  }
  return tguard;
}

// Position of a thread's pixel pair: image b and pair index p2 inside the image, advanced by the grid stride without a division.
struct PairPos {
  int b;
  int64_t p2;
  __device__ __forceinline__ void init(long long i, int64_t HW2) {
    b = static_cast<int>(i / HW2);
    p2 = i - static_cast<long long>(b) * HW2;
  }
  __device__ __forceinline__ PairPos next(long long stride, int64_t HW2) const {
    PairPos n = {b, p2 + stride};
    while (n.p2 >= HW2) {
      n.p2 -= HW2;
      ++n.b;
    }
    return n;
  }
  __device__ __forceinline__ size_t label_pos(int64_t HW) const { return static_cast<size_t>(b) * HW + p2 * kPxL; }
};

template <int C>
__global__ void __launch_bounds__(kThreadsL, 2) k_loss_fwd_pk(LossArgs a, Partial* __restrict__ partials) {
  extern __shared__ __align__(128) float2 s_loss_stage[];
  const LossStage<C> st(s_loss_stage, a);
  const long long total = static_cast<long long>(a.B) * st.HW2;
  const long long stride = static_cast<long long>(gridDim.x) * kThreadsL;
  const bool want_cst = (a.terms & HIAST_TERM_CST) != 0;
  double acc[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  long long i = static_cast<long long>(blockIdx.x) * kThreadsL + threadIdx.x;
  PairPos pos = {0, 0};
  int ya = 0, yb = 0;
  if (i < total) {
    pos.init(i, st.HW2);
    st.prefetch_z_at(pos.b, pos.p2);
    st.prefetch_t_at(pos.b, pos.p2);
    load_label_pair(a.plbl, a.plbl_bytes, pos.label_pos(a.HW), ya, yb);
  }
  for (; i < total; i += stride) {
    st.wait_older();                       // z of this pair (its t may still be in flight)
    float z[kPxL][C];
    float zya, zyb;
    const float guard = fmaxf(st.load_z(z), st.load_zy(ya, yb, zya, zyb));
    const bool more = i + stride < total;
    const PairPos nxt = pos.next(stride, st.HW2);
    int nya = 0, nyb = 0;
    if (more && guard == guard) {
      st.prefetch_z_at(nxt.b, nxt.p2);
      load_label_pair(a.plbl, a.plbl_bytes, nxt.label_pos(a.HW), nya, nyb);
    } else {
      asm volatile("cp.async.commit_group;\n" ::: "memory");
    }
    const bool iga = (ya == HIAST_IGNORE_LABEL), igb = (yb == HIAST_IGNORE_LABEL);
    PairLS<C> ls;
    ls.init(z[0], z[1], ya, yb, zya, zyb);
    st.wait_older();                       // this pair's teacher probabilities have landed in the staging slots
    PairCst<C> cs;
    cs.sc2 = cs.T2 = 0;
    cs.mna = cs.mnb = INFINITY;
    if (want_cst) cs.run(z[0], z[1], ls, st);
    pair_sums<C>(ls, cs, iga, igb, a, pos.b, pos.p2 * kPxL, [&](int k, double v) { acc[k] += v; });
    const float tguard = fminf(cs.mna, cs.mnb);   // never NaN; depends on every read of the t slots
    if (more && tguard == tguard) st.prefetch_t_at(nxt.b, nxt.p2);
    else asm volatile("cp.async.commit_group;\n" ::: "memory");
    pos = nxt;
    ya = nya;
    yb = nyb;
  }
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  PixelSums sums;
  sums.ce = acc[0]; sums.kld = acc[1]; sums.ent = acc[2]; sums.cst = acc[3];
  sums.n_conf = static_cast<long long>(acc[4]);      // counts as exact doubles (< 2^53)
  sums.n_ign = static_cast<long long>(acc[5]);
  sums.n_nz = static_cast<long long>(acc[6]);
  block_reduce_store(sums, partials);
}

// Upstream gradients of the four loss terms as autograd delivers them (hiast_st_loss_bwd_checked_terms): device pointers to
// f32 scalars (nullptr: no gradient flows into that term) and the divisors the forward call wrote.  The scales are derived
// here -- float(double(g_k) / divisor_k), the arithmetic of hiast_b200/losses.py -- instead of by four tiny torch kernels.
struct GoutArgs {
  const float* g[4];
  const double* divisors;      // nullptr: take `scales` as given
  const float* hint_weights;   // with upstream_out: *upstream_out = gout_{k0} / hint_weights[k0] (what losses.GradHint observes)
  float* upstream_out;
  int k0;
  const float* term_weights;   // != nullptr: the outputs were the WEIGHTED terms w_k * loss_k, so gout_k = *g[k] * w_k (f32, as
                               // autograd's MulBackward would have computed it)
};

template <int C>
__global__ void __launch_bounds__(kThreadsL, 2) k_loss_bwd_pk(LossArgs a, const float* __restrict__ scales,
                                                              float* __restrict__ grad, const float* __restrict__ scales_used,
                                                              GoutArgs go) {
  extern __shared__ __align__(128) float2 s_loss_stage[];
  const LossStage<C> st(s_loss_stage, a);
  const long long total = static_cast<long long>(a.B) * st.HW2;
  const long long stride = static_cast<long long>(gridDim.x) * kThreadsL;
  float sc[4];
  if (go.divisors) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float g = 0.f;
      if (((a.terms >> k) & 1) && go.g[k]) g = go.term_weights ? __fmul_rn(*go.g[k], go.term_weights[k]) : *go.g[k];
      sc[k] = (((a.terms >> k) & 1) && go.g[k]) ? static_cast<float>(static_cast<double>(g) / go.divisors[k]) : 0.f;
      if (k == go.k0 && blockIdx.x == 0 && threadIdx.x == 0 && go.upstream_out && go.g[k])
        *go.upstream_out = __fdiv_rn(g, go.hint_weights[k]);
    }
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) sc[k] = scales[k];
  }
  if (scales_used) {
    // `grad` already holds the gradient the one-pass kernel wrote for the scales it ASSUMED: nothing to do if the scales
    // autograd actually delivered are the same bits (every thread takes the same branch; no barrier has been reached)
    bool same = true;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (a.terms & (1 << k)) same = same && (__float_as_uint(sc[k]) == __float_as_uint(scales_used[k]));
    if (same) return;
  }
  GradScales gsc;
  gsc.init(sc, a.terms);
  const bool want_cst = (a.terms & HIAST_TERM_CST) != 0;
  long long i = static_cast<long long>(blockIdx.x) * kThreadsL + threadIdx.x;
  PairPos pos = {0, 0};
  int ya = 0, yb = 0;
  if (i < total) {
    pos.init(i, st.HW2);
    st.prefetch_z_at(pos.b, pos.p2);
    st.prefetch_t_at(pos.b, pos.p2);
    load_label_pair(a.plbl, a.plbl_bytes, pos.label_pos(a.HW), ya, yb);
  }
  for (; i < total; i += stride) {
    st.wait_older();                       // z of this pair
    float z[kPxL][C];
    float zya, zyb;
    const float guard = fmaxf(st.load_z(z), st.load_zy(ya, yb, zya, zyb));
    const bool more = i + stride < total;
    const PairPos nxt = pos.next(stride, st.HW2);
    int nya = 0, nyb = 0;
    if (more && guard == guard) {
      st.prefetch_z_at(nxt.b, nxt.p2);
      load_label_pair(a.plbl, a.plbl_bytes, nxt.label_pos(a.HW), nya, nyb);
    } else {
      asm volatile("cp.async.commit_group;\n" ::: "memory");
    }
    const int b = pos.b;
    const int64_t p2i = pos.p2;
    const bool iga = (ya == HIAST_IGNORE_LABEL), igb = (yb == HIAST_IGNORE_LABEL);
    PairLS<C> ls;
    ls.init(z[0], z[1], ya, yb, zya, zyb);
    const bool csa = want_cst && in_region(a.region, iga);
    const bool csb = want_cst && in_region(a.region, igb);
    st.wait_older();                       // this pair's teacher probabilities
    pk::u64 T2 = 0;
    if (csa || csb) {
#pragma unroll
      for (int c = 0; c < C; ++c) T2 = (c == 0) ? st.t2(c) : pk::add2(T2, st.t2(c));
    }
    float2* gs = reinterpret_cast<float2*>(grad + static_cast<size_t>(b) * C * a.HW) + p2i;
    const float tguard = pair_gradient_store<C>(z[0], z[1], ls, st, gsc, ya, yb, iga, igb, csa, csb, T2, gs, st.HW2);
    if (more && tguard == tguard) st.prefetch_t_at(nxt.b, nxt.p2);    // the t slots have been read
    else asm volatile("cp.async.commit_group;\n" ::: "memory");
    pos = nxt;
    ya = nya;
    yb = nyb;
  }
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}

// ---- one-pass forward + backward ------------------------------------------------------------------
// The two-pass design reads (z, t, plbl) twice: 160 + 236 = 396 B/px.  The CE / KLD / ENT divisors depend on the labels only
// (n_conf, n_ign) and the SoftCE divisor is C * n_region unless some product is exactly zero, so after a label-only pre-pass
// (1 or 8 B/px) ONE pass can write the gradient for an ASSUMED upstream gradient per term while it accumulates the forward
// sums: 236 (+ labels) B/px.  autograd later delivers the real upstream gradients; hiast_st_loss_bwd_checked compares the
// scales they imply with the ones used here and rewrites the gradient only when they differ (first step after a change of
// the loss scale, an exactly-zero SoftCE product, weights changed): exact in every case, one pass in the steady state.
struct LabelCounts {
  unsigned long long n_conf, n_ign;
};

// Programmatic dependent launch (sm_90+): a kernel launched with the stream-serialisation attribute may start while its
// predecessor in the stream is still running; it must execute pdl_wait() before touching anything the predecessor writes.
// pdl_launch_dependents() in the predecessor lets the dependent grid be scheduled from that point on (otherwise: at its exit).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }

// Label-only pre-pass: block k writes its (n_conf, n_ign) to out[k] -- no atomics, hence no memset in front of it.  The grid is
// one wave, and every block releases the dependent launch at once: the one-pass kernel starts beside it and only waits
// (pdl_wait) when it first needs the counts.
__global__ void __launch_bounds__(256) k_label_count(const void* __restrict__ plbl, int plbl_bytes, long long n,
                                                     LabelCounts* __restrict__ out) {
  pdl_launch_dependents();
  long long ign = 0, tot = 0;
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long nth = static_cast<long long>(gridDim.x) * blockDim.x;
  if (plbl_bytes == 8 && (reinterpret_cast<uintptr_t>(plbl) % 16 == 0)) {
    const longlong2* p = static_cast<const longlong2*>(plbl);
    for (long long i = tid; i < n / 2; i += nth) {
      const longlong2 v = __ldg(p + i);
      ign += (v.x == HIAST_IGNORE_LABEL) + (v.y == HIAST_IGNORE_LABEL);
      tot += 2;
    }
    if (tid == 0 && (n & 1)) {
      ign += (static_cast<const long long*>(plbl)[n - 1] == HIAST_IGNORE_LABEL);
      tot += 1;
    }
  } else if (plbl_bytes == 1 && (reinterpret_cast<uintptr_t>(plbl) % 16 == 0)) {
    const uint4* p = static_cast<const uint4*>(plbl);
    for (long long i = tid; i < n / 16; i += nth) {
      const uint4 v = __ldg(p + i);
      ign += __popc(__vcmpeq4(v.x, 0xffffffffu) & 0x01010101u) + __popc(__vcmpeq4(v.y, 0xffffffffu) & 0x01010101u) +
             __popc(__vcmpeq4(v.z, 0xffffffffu) & 0x01010101u) + __popc(__vcmpeq4(v.w, 0xffffffffu) & 0x01010101u);
      tot += 16;
    }
    for (long long i = (n / 16) * 16 + tid; i < n; i += nth) {
      ign += (static_cast<const uint8_t*>(plbl)[i] == HIAST_IGNORE_LABEL);
      tot += 1;
    }
  } else {
    for (long long i = tid; i < n; i += nth) {
      ign += (load_label(plbl, plbl_bytes, static_cast<size_t>(i)) == HIAST_IGNORE_LABEL);
      tot += 1;
    }
  }
  __shared__ long long s_ign[8], s_tot[8];
  ign = warp_sum(ign);
  tot = warp_sum(tot);
  if (lane_id() == 0) {
    s_ign[threadIdx.x >> 5] = ign;
    s_tot[threadIdx.x >> 5] = tot;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) {
      ign += s_ign[w];
      tot += s_tot[w];
    }
    out[blockIdx.x].n_ign = static_cast<unsigned long long>(ign);
    out[blockIdx.x].n_conf = static_cast<unsigned long long>(tot - ign);
  }
}

// every thread of the block returns the totals of the pre-pass's per-block counts
__device__ __forceinline__ LabelCounts sum_label_counts(const LabelCounts* __restrict__ parts, int n_parts) {
  __shared__ unsigned long long s_c[kThreadsL / 32], s_i[kThreadsL / 32];
  unsigned long long nc = 0, ni = 0;
  for (int k = threadIdx.x; k < n_parts; k += kThreadsL) {
    nc += parts[k].n_conf;
    ni += parts[k].n_ign;
  }
  nc = static_cast<unsigned long long>(warp_sum(static_cast<long long>(nc)));
  ni = static_cast<unsigned long long>(warp_sum(static_cast<long long>(ni)));
  if (lane_id() == 0) {
    s_c[threadIdx.x >> 5] = nc;
    s_i[threadIdx.x >> 5] = ni;
  }
  __syncthreads();
  LabelCounts r = {0, 0};
#pragma unroll
  for (int w = 0; w < kThreadsL / 32; ++w) {
    r.n_conf += s_c[w];
    r.n_ign += s_i[w];
  }
  return r;
}

// scales of the gradient for upstream gradients gw[4]: float(double(gw) / divisor), the arithmetic of the autograd
// Function (hiast_b200/losses.py) so that equal inputs give equal bits
__device__ __forceinline__ void fused_scales(const LabelCounts& lc, const float* gw, int C, int region, float (&sc)[4]) {
  const double nc = static_cast<double>(lc.n_conf), ni = static_cast<double>(lc.n_ign);
  const double nr = region == HIAST_REGION_ALL ? nc + ni : (region == HIAST_REGION_IGNORED ? ni : nc);
  sc[0] = static_cast<float>(static_cast<double>(gw[0]) / nc);
  sc[1] = static_cast<float>(static_cast<double>(gw[1]) / (C * nc));
  sc[2] = static_cast<float>(static_cast<double>(gw[2]) / (C * ni));
  sc[3] = static_cast<float>(static_cast<double>(gw[3]) / (C * nr));
}

template <int C>
__global__ void __launch_bounds__(kThreadsL, 2) k_loss_fused_pk(LossArgs a, const LabelCounts* __restrict__ lc,
                                                                const float* __restrict__ grad_weights,
                                                                float* __restrict__ scales_used, float* __restrict__ grad,
                                                                Partial* __restrict__ partials, int n_count_parts) {
  extern __shared__ __align__(128) float2 s_loss_stage[];
  const LossStage<C> st(s_loss_stage, a);
  const long long total = static_cast<long long>(a.B) * st.HW2;
  const long long stride = static_cast<long long>(gridDim.x) * kThreadsL;
  const bool want_cst = (a.terms & HIAST_TERM_CST) != 0;
  // the seven forward accumulators live in a per-thread column of shared memory, not in registers: the gradient pass needs
  // every register the two-pass backward kernel uses (128 per thread at two CTAs per SM)
  double* sacc = reinterpret_cast<double*>(s_loss_stage + 2 * C * kThreadsL) + threadIdx.x;
#pragma unroll
  for (int k = 0; k < 7; ++k) sacc[k * kThreadsL] = 0.0;
  long long i = static_cast<long long>(blockIdx.x) * kThreadsL + threadIdx.x;
  PairPos pos = {0, 0};
  int ya = 0, yb = 0;
  if (i < total) {
    pos.init(i, st.HW2);
    st.prefetch_z_at(pos.b, pos.p2);
    st.prefetch_t_at(pos.b, pos.p2);
    load_label_pair(a.plbl, a.plbl_bytes, pos.label_pos(a.HW), ya, yb);
  }
  // The label counts come from the pre-pass this grid was launched beside (programmatic dependent launch): the first pair's
  // logits, teacher probabilities and labels are already on their way when the block waits for it.  Every thread of the block
  // takes part (block-wide sum of the pre-pass's per-block counts), also those without a pair.
  pdl_wait();
  float sc[4];
  {
    const LabelCounts c0 = sum_label_counts(lc, n_count_parts);
    const float gw[4] = {grad_weights[0], grad_weights[1], grad_weights[2], grad_weights[3]};
    fused_scales(c0, gw, C, a.region, sc);
    if (blockIdx.x == 0 && threadIdx.x < 4) scales_used[threadIdx.x] = sc[threadIdx.x];
  }
  GradScales gsc;
  gsc.init(sc, a.terms);
  for (; i < total; i += stride) {
    st.wait_older();                       // z of this pair
    float z[kPxL][C];
    float zya, zyb;
    const float guard = fmaxf(st.load_z(z), st.load_zy(ya, yb, zya, zyb));
    const bool more = i + stride < total;
    const PairPos nxt = pos.next(stride, st.HW2);
    int nya = 0, nyb = 0;
    if (more && guard == guard) {
      st.prefetch_z_at(nxt.b, nxt.p2);
      load_label_pair(a.plbl, a.plbl_bytes, nxt.label_pos(a.HW), nya, nyb);
    } else {
      asm volatile("cp.async.commit_group;\n" ::: "memory");
    }
    const int b = pos.b;
    const int64_t p2i = pos.p2;
    const bool iga = (ya == HIAST_IGNORE_LABEL), igb = (yb == HIAST_IGNORE_LABEL);
    PairLS<C> ls;
    ls.init(z[0], z[1], ya, yb, zya, zyb);           // softmax pass: everything CE / KLD / ENT need falls out of it
    const bool csa = want_cst && in_region(a.region, iga);
    const bool csb = want_cst && in_region(a.region, igb);
    st.wait_older();                       // this pair's teacher probabilities
    PairCst<C> cs;
    cs.sc2 = cs.T2 = 0;
    cs.mna = cs.mnb = INFINITY;
    if (want_cst) cs.run(z[0], z[1], ls, st);   // pass 1 (SoftCE only): the sum of products, their zero-ness, the teacher mass
    pair_sums<C>(ls, cs, iga, igb, a, b, p2i * kPxL, [&](int k, double v) { sacc[k * kThreadsL] += v; });
    // pass 2: the gradient, channel by channel right before the store
    float2* gs = reinterpret_cast<float2*>(grad + static_cast<size_t>(b) * C * a.HW) + p2i;
    const float tguard = pair_gradient_store<C>(z[0], z[1], ls, st, gsc, ya, yb, iga, igb, csa, csb, cs.T2, gs, st.HW2);
    if (more && tguard == tguard) st.prefetch_t_at(nxt.b, nxt.p2);    // the t slots have been read
    else asm volatile("cp.async.commit_group;\n" ::: "memory");
    pos = nxt;
    ya = nya;
    yb = nyb;
  }
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  pdl_launch_dependents();                 // the finalize kernel may be scheduled; it waits for this grid before it reads
  PixelSums acc;
  acc.ce = sacc[0 * kThreadsL]; acc.kld = sacc[1 * kThreadsL]; acc.ent = sacc[2 * kThreadsL]; acc.cst = sacc[3 * kThreadsL];
  acc.n_conf = static_cast<long long>(sacc[4 * kThreadsL]);      // counts as exact doubles (< 2^53)
  acc.n_ign = static_cast<long long>(sacc[5 * kThreadsL]);
  acc.n_nz = static_cast<long long>(sacc[6 * kThreadsL]);
  block_reduce_store(acc, partials);
}

// Generic path: runtime C <= 255, any HW; channel column re-read through L1.
constexpr int kMaxCGeneric = 255;

__device__ __forceinline__ void generic_ls(const float* __restrict__ zp, int64_t cs, int C, float& m, float& logs, float& inv_s) {
  m = zp[0];
  for (int c = 1; c < C; ++c) m = fmaxf(m, zp[c * cs]);
  float s = 0.f;
  for (int c = 0; c < C; ++c) s += expf(zp[c * cs] - m);
  logs = logf(s);
  inv_s = 1.0f / s;
}

__global__ void __launch_bounds__(kThreadsL) k_loss_fwd_generic(LossArgs a, Partial* __restrict__ partials) {
  const long long total = static_cast<long long>(a.B) * a.HW;
  PixelSums acc = {0.0, 0.0, 0.0, 0.0, 0, 0, 0};
  const int C = a.C;
  for (long long i = static_cast<long long>(blockIdx.x) * kThreadsL + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * kThreadsL) {
    const int b = static_cast<int>(i / a.HW);
    const int64_t p = i - static_cast<long long>(b) * a.HW;
    const float* zp = a.z + static_cast<size_t>(b) * C * a.HW + p;
    const float* tp = a.t ? a.t + static_cast<size_t>(b) * C * a.HW + p : nullptr;
    float m, logs, inv_s;
    generic_ls(zp, a.HW, C, m, logs, inv_s);
    const int y = load_label(a.plbl, a.plbl_bytes, static_cast<size_t>(i));
    const bool ignored = (y == HIAST_IGNORE_LABEL);
    if (ignored) acc.n_ign += 1; else acc.n_conf += 1;
    if (!ignored) {
      if ((a.terms & HIAST_TERM_CE) && y < C) acc.ce += static_cast<double>(-((zp[y * a.HW] - m) - logs));
      if (a.terms & HIAST_TERM_KLD) {
        float sl = 0.f;
        for (int c = 0; c < C; ++c) sl += (zp[c * a.HW] - m) - logs;
        acc.kld += static_cast<double>(-sl * (1.0f / C));
      }
    } else if (a.terms & HIAST_TERM_ENT) {
      float h = 0.f;
      for (int c = 0; c < C; ++c) {
        const float zc = zp[c * a.HW];
        h += (expf(zc - m) * inv_s) * ((zc - m) - logs);
      }
      acc.ent += static_cast<double>(-h);
    }
    if ((a.terms & HIAST_TERM_CST) && in_region(a.region, ignored)) {
      float sc = 0.f;
      int nz = 0;
      const int kind = cst_kind(a.terms);
      float tm = 0.f, tlogs = 0.f, tinv = 0.f, tsum = 1.f;
      if (kind == HIAST_CST_KLDIV || kind == HIAST_CST_SOFTCE_LOGITS) {
        generic_ls(tp, a.HW, C, tm, tlogs, tinv);
        tsum = 0.f;
        for (int c = 0; c < C; ++c) tsum += expf(tp[c * a.HW] - tm);
      }
      for (int c = 0; c < C; ++c) {
        const float lp = (zp[c * a.HW] - m) - logs;
        float elem;
        if (kind == HIAST_CST_KLDIV) {
          const float pr = __fdiv_rn(expf(tp[c * a.HW] - tm), tsum);
          elem = (pr == 0.f ? 0.f : __fmul_rn(pr, logf(pr))) - __fmul_rn(pr, lp);
        } else if (kind == HIAST_CST_MSE) {
          const float d = zp[c * a.HW] - tp[c * a.HW];
          elem = __fmul_rn(d, d);
        } else if (kind == HIAST_CST_SOFTCE_LOGITS) {
          elem = __fmul_rn(-lp, __fdiv_rn(expf(tp[c * a.HW] - tm), tsum));
        } else {
          elem = __fmul_rn(-lp, tp[c * a.HW]);
        }
        nz += (elem != 0.f);
        sc += elem;
      }
      acc.cst += static_cast<double>(sc);
      acc.n_nz += nz;
    }
  }
  block_reduce_store(acc, partials);
}

__global__ void __launch_bounds__(kThreadsL) k_loss_bwd_generic(LossArgs a, const float* __restrict__ scales,
                                                                float* __restrict__ grad) {
  const long long total = static_cast<long long>(a.B) * a.HW;
  const int C = a.C;
  const float s_ce = (a.terms & HIAST_TERM_CE) ? scales[0] : 0.f;
  const float s_kld = (a.terms & HIAST_TERM_KLD) ? scales[1] : 0.f;
  const float s_ent = (a.terms & HIAST_TERM_ENT) ? scales[2] : 0.f;
  const float s_cst = (a.terms & HIAST_TERM_CST) ? scales[3] : 0.f;
  const float poison = 0.f * s_ce + 0.f * s_kld + 0.f * s_ent + 0.f * s_cst;  // NaN iff an enabled scale is non-finite
  for (long long i = static_cast<long long>(blockIdx.x) * kThreadsL + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * kThreadsL) {
    const int b = static_cast<int>(i / a.HW);
    const int64_t p = i - static_cast<long long>(b) * a.HW;
    const size_t off = static_cast<size_t>(b) * C * a.HW + p;
    const float* zp = a.z + off;
    const float* tp = a.t ? a.t + off : nullptr;
    float* gp = grad + off;
    float m, logs, inv_s;
    generic_ls(zp, a.HW, C, m, logs, inv_s);
    const int y = load_label(a.plbl, a.plbl_bytes, static_cast<size_t>(i));
    const bool ignored = (y == HIAST_IGNORE_LABEL);
    const bool cst = (a.terms & HIAST_TERM_CST) && in_region(a.region, ignored);
    float T = 0.f, h = 0.f;
    const int kind = cst_kind(a.terms);
    float tm = 0.f, tlogs = 0.f, tinv = 0.f, tsum = 1.f;
    const bool t_is_logits = (kind == HIAST_CST_KLDIV || kind == HIAST_CST_SOFTCE_LOGITS);
    if (cst && t_is_logits) {
      generic_ls(tp, a.HW, C, tm, tlogs, tinv);
      tsum = 0.f;
      for (int c = 0; c < C; ++c) tsum += expf(tp[c * a.HW] - tm);
    }
    auto tval = [&](int c) {
      return t_is_logits ? __fdiv_rn(expf(tp[c * a.HW] - tm), tsum) : tp[c * a.HW];
    };
    if (cst && kind != HIAST_CST_MSE) for (int c = 0; c < C; ++c) T += tval(c);
    if (ignored && (a.terms & HIAST_TERM_ENT))
      for (int c = 0; c < C; ++c) {
        const float zc = zp[c * a.HW];
        h += (expf(zc - m) * inv_s) * ((zc - m) - logs);
      }
    for (int c = 0; c < C; ++c) {
      const float zc = zp[c * a.HW];
      const float pc = expf(zc - m) * inv_s;
      float g = 0.f;
      if (!ignored) g = (s_ce + s_kld) * pc - s_kld * (1.0f / C) - ((c == y) ? s_ce : 0.f);
      else if (a.terms & HIAST_TERM_ENT) g = -s_ent * pc * (((zc - m) - logs) - h);
      if (cst) g += (kind == HIAST_CST_MSE) ? s_cst * 2.0f * (zc - tp[c * a.HW]) : s_cst * (pc * T - tval(c));
      gp[c * a.HW] = g + poison;
    }
  }
}

// Final deterministic reduction of the per-CTA partials (one CTA, fixed order).
// `losses` != nullptr: also the four unweighted loss terms f32[4] = float(sum / divisor) (0 for a disabled term, NaN for an empty
// region exactly like the reference's 0/0) and the divisors f64[4] = (n_conf, C n_conf, C n_ign, #non-zero SoftCE products) --
// the arithmetic of hiast_b200/losses.py FusedSelfTrainingLoss.forward, in this launch instead of a dozen tiny torch kernels.
__global__ void k_loss_finalize(const Partial* __restrict__ partials, int n, double* __restrict__ sums,
                                long long* __restrict__ counts, float* __restrict__ losses, double* __restrict__ divisors,
                                int C, int terms, const float* __restrict__ term_weights) {
  __shared__ Partial s_part[kThreadsL / 32];
  pdl_wait();                              // no-op unless launched as a programmatic dependent (hiast_st_loss_fused)
  PixelSums acc = {0.0, 0.0, 0.0, 0.0, 0, 0, 0};
  for (int i = threadIdx.x; i < n; i += kThreadsL) {
    const Partial p = partials[i];
    acc.ce += p.s[0]; acc.kld += p.s[1]; acc.ent += p.s[2]; acc.cst += p.s[3];
    acc.n_conf += p.n[0]; acc.n_ign += p.n[1]; acc.n_nz += p.n[2];
  }
  double s0 = warp_sum(acc.ce), s1 = warp_sum(acc.kld), s2 = warp_sum(acc.ent), s3 = warp_sum(acc.cst);
  long long n0 = warp_sum(acc.n_conf), n1 = warp_sum(acc.n_ign), n2 = warp_sum(acc.n_nz);
  if (lane_id() == 0) {
    Partial& p = s_part[threadIdx.x >> 5];
    p.s[0] = s0; p.s[1] = s1; p.s[2] = s2; p.s[3] = s3;
    p.n[0] = n0; p.n[1] = n1; p.n[2] = n2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    Partial r = s_part[0];
    for (int w = 1; w < kThreadsL / 32; ++w) {
      for (int k = 0; k < 4; ++k) r.s[k] += s_part[w].s[k];
      for (int k = 0; k < 3; ++k) r.n[k] += s_part[w].n[k];
    }
    for (int k = 0; k < 4; ++k) sums[k] = r.s[k];
    for (int k = 0; k < 3; ++k) counts[k] = r.n[k];
    if (losses) {
      const double den[4] = {static_cast<double>(r.n[0]), static_cast<double>(C) * static_cast<double>(r.n[0]),
                             static_cast<double>(C) * static_cast<double>(r.n[1]), static_cast<double>(r.n[2])};
      for (int k = 0; k < 4; ++k) {
        divisors[k] = den[k];
        const float v = ((terms >> k) & 1) ? static_cast<float>(r.s[k] / den[k]) : 0.f;
        losses[k] = term_weights ? __fmul_rn(term_weights[k], v) : v;      // w_k * loss_k in f32, as torch multiplies them
      }
    }
  }
}

int loss_grid(long long work_items) {
  const long long want = (work_items + kThreadsL - 1) / kThreadsL;
  const long long cap = static_cast<long long>(sm_count()) * 2 * 4;  // 2 resident CTAs/SM, 4 waves of work each
  return static_cast<int>(std::max<long long>(1, std::min(want, cap)));
}

bool loss_vector_ok(const LossArgs& a, const void* extra) {
  return (a.C == 19 || a.C == 16) && (a.HW % kPxL == 0) && (reinterpret_cast<uintptr_t>(a.z) % 8 == 0) &&
         (!a.t || reinterpret_cast<uintptr_t>(a.t) % 8 == 0) && (reinterpret_cast<uintptr_t>(extra) % 8 == 0);
}

size_t loss_stage_bytes(int C) { return sizeof(float2) * 2 * C * kThreadsL; }

int loss_configure_smem(int C, size_t smem) {
  if (C == 19) {
    HIAST_TRY(ensure_dyn_smem(k_loss_fwd<19>, smem));
    HIAST_TRY(ensure_dyn_smem(k_loss_bwd<19>, smem));
    HIAST_TRY(ensure_dyn_smem(k_loss_fwd_pk<19>, smem));
    HIAST_TRY(ensure_dyn_smem(k_loss_bwd_pk<19>, smem));
  } else {
    HIAST_TRY(ensure_dyn_smem(k_loss_fwd<16>, smem));
    HIAST_TRY(ensure_dyn_smem(k_loss_bwd<16>, smem));
    HIAST_TRY(ensure_dyn_smem(k_loss_fwd_pk<16>, smem));
    HIAST_TRY(ensure_dyn_smem(k_loss_bwd_pk<16>, smem));
  }
  return HIAST_OK;
}

inline int cst_kind_host(int terms) { return terms & HIAST_CST_SOFTCE_LOGITS; }
bool g_loss_scalar = false;   // development switch: force the scalar vector kernels (hiast_debug_loss_scalar)

}  // namespace hiast

using namespace hiast;

extern "C" int hiast_debug_loss_scalar(int on) {
  g_loss_scalar = on != 0;
  return HIAST_OK;
}

extern "C" size_t hiast_st_loss_workspace_bytes(int B, int C, int64_t HW) {
  (void)C;
  if (B < 0 || HW < 0) return 0;
  return static_cast<size_t>(loss_grid(static_cast<long long>(B) * HW)) * sizeof(Partial);
}

static int check_loss_args(const float* z, const float* t, const void* plbl, int plbl_bytes, int B, int C, int64_t HW,
                           int region, int terms) {
  if (!z || !plbl) return HIAST_ERR_INVALID_ARG;
  if (plbl_bytes != 1 && plbl_bytes != 8) return HIAST_ERR_INVALID_ARG;
  if (B < 0 || C < 1 || C > kMaxCGeneric || HW < 1) return HIAST_ERR_INVALID_ARG;
  if (region < HIAST_REGION_IGNORED || region > HIAST_REGION_ALL) return HIAST_ERR_INVALID_ARG;
  if (terms < 0 || terms > 63) return HIAST_ERR_INVALID_ARG;
  if ((terms & HIAST_TERM_CST) && !t) return HIAST_ERR_INVALID_ARG;
  return HIAST_OK;
}

extern "C" int hiast_st_loss_fwd(const float* z, const float* t, const void* plbl, int plbl_bytes, int B, int C,
                                 int64_t HW, int region, int terms, double* sums, int64_t* counts, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  const int rc = check_loss_args(z, t, plbl, plbl_bytes, B, C, HW, region, terms);
  if (rc != HIAST_OK) return rc;
  if (!sums || !counts || !workspace) return HIAST_ERR_INVALID_ARG;
  if (workspace_bytes < hiast_st_loss_workspace_bytes(B, C, HW)) return HIAST_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  LossArgs a = {z, t, plbl, plbl_bytes, B, C, HW, region, terms};
  Partial* parts = static_cast<Partial*>(workspace);
  int grid = loss_grid(static_cast<long long>(B) * HW);
  if (loss_vector_ok(a, nullptr)) {
    grid = std::min(grid, sm_count() * 2);   // persistent: one resident wave, every thread pipelines its own sequence
    const size_t smem = loss_stage_bytes(C);
    HIAST_TRY(loss_configure_smem(C, smem));
    const bool packed = cst_kind_host(terms) == HIAST_CST_SOFTCE && !g_loss_scalar;
    if (packed) {
      if (C == 19) k_loss_fwd_pk<19><<<grid, kThreadsL, smem, st>>>(a, parts);
      else k_loss_fwd_pk<16><<<grid, kThreadsL, smem, st>>>(a, parts);
    } else {
      if (C == 19) k_loss_fwd<19><<<grid, kThreadsL, smem, st>>>(a, parts);
      else k_loss_fwd<16><<<grid, kThreadsL, smem, st>>>(a, parts);
    }
  } else {
    k_loss_fwd_generic<<<grid, kThreadsL, 0, st>>>(a, parts);
  }
  HIAST_CHECK_LAUNCH();
  k_loss_finalize<<<1, kThreadsL, 0, st>>>(parts, grid, sums, reinterpret_cast<long long*>(counts), nullptr, nullptr, C, terms, nullptr);
  HIAST_CHECK_LAUNCH();
  return HIAST_OK;
}

extern "C" int hiast_st_loss_bwd(const float* z, const float* t, const void* plbl, int plbl_bytes, int B, int C,
                                 int64_t HW, int region, int terms, const float* scales, float* grad_z, void* stream) {
  const int rc = check_loss_args(z, t, plbl, plbl_bytes, B, C, HW, region, terms);
  if (rc != HIAST_OK) return rc;
  if (!scales || !grad_z) return HIAST_ERR_INVALID_ARG;
  if (B == 0) return HIAST_OK;
  cudaStream_t st = as_stream(stream);
  LossArgs a = {z, t, plbl, plbl_bytes, B, C, HW, region, terms};
  int grid = loss_grid(static_cast<long long>(B) * HW);
  if (loss_vector_ok(a, grad_z)) {
    grid = std::min(grid, sm_count() * 2);
    const size_t smem = loss_stage_bytes(C);
    HIAST_TRY(loss_configure_smem(C, smem));
    const bool packed = cst_kind_host(terms) == HIAST_CST_SOFTCE && !g_loss_scalar;
    if (packed) {
      const GoutArgs none = {{nullptr, nullptr, nullptr, nullptr}, nullptr, nullptr, nullptr, 0, nullptr};
      if (C == 19) k_loss_bwd_pk<19><<<grid, kThreadsL, smem, st>>>(a, scales, grad_z, nullptr, none);
      else k_loss_bwd_pk<16><<<grid, kThreadsL, smem, st>>>(a, scales, grad_z, nullptr, none);
    } else {
      if (C == 19) k_loss_bwd<19><<<grid, kThreadsL, smem, st>>>(a, scales, grad_z);
      else k_loss_bwd<16><<<grid, kThreadsL, smem, st>>>(a, scales, grad_z);
    }
  } else {
    k_loss_bwd_generic<<<grid, kThreadsL, 0, st>>>(a, scales, grad_z);
  }
  HIAST_CHECK_LAUNCH();
  return HIAST_OK;
}

constexpr int kCountBlocksMax = 1024;                                   // per-block label counts in front of the partial sums
constexpr size_t kFusedHeaderBytes = kCountBlocksMax * sizeof(LabelCounts);

extern "C" size_t hiast_st_loss_fused_workspace_bytes(int B, int C, int64_t HW) {
  return hiast_st_loss_workspace_bytes(B, C, HW) + kFusedHeaderBytes;
}

// kernel<<<grid, block, smem, st>>>(args...) as a programmatic dependent of the kernel launched before it on `st`
template <class... KArgs, class... Args>
static cudaError_t launch_dependent(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(static_cast<unsigned>(block));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

extern "C" int hiast_st_loss_fused(const float* z, const float* t, const void* plbl, int plbl_bytes, int B, int C, int64_t HW,
                                   int region, int terms, const float* grad_weights, double* sums, int64_t* counts,
                                   float* scales_used, float* grad_z, void* workspace, size_t workspace_bytes, void* stream) {
  return hiast_st_loss_fused_terms(z, t, plbl, plbl_bytes, B, C, HW, region, terms, grad_weights, sums, counts, scales_used, grad_z,
                                   nullptr, nullptr, nullptr, workspace, workspace_bytes, stream);
}

extern "C" int hiast_st_loss_fused_terms(const float* z, const float* t, const void* plbl, int plbl_bytes, int B, int C,
                                         int64_t HW, int region, int terms, const float* grad_weights, double* sums,
                                         int64_t* counts, float* scales_used, float* grad_z, float* losses, double* divisors,
                                         const float* term_weights, void* workspace, size_t workspace_bytes, void* stream) {
  if ((losses == nullptr) != (divisors == nullptr)) return HIAST_ERR_INVALID_ARG;
  const int rc = check_loss_args(z, t, plbl, plbl_bytes, B, C, HW, region, terms);
  if (rc != HIAST_OK) return rc;
  if (!grad_weights || !sums || !counts || !scales_used || !grad_z || !workspace) return HIAST_ERR_INVALID_ARG;
  if (workspace_bytes < hiast_st_loss_fused_workspace_bytes(B, C, HW)) return HIAST_ERR_WORKSPACE;
  LossArgs a = {z, t, plbl, plbl_bytes, B, C, HW, region, terms};
  if (B == 0 || !loss_vector_ok(a, grad_z) || cst_kind_host(terms) != HIAST_CST_SOFTCE || g_loss_scalar ||
      reinterpret_cast<uintptr_t>(workspace) % 16 != 0)
    return HIAST_ERR_UNSUPPORTED;         // callers take hiast_st_loss_fwd + hiast_st_loss_bwd
  cudaStream_t st = as_stream(stream);
  LabelCounts* lc = static_cast<LabelCounts*>(workspace);
  Partial* parts = reinterpret_cast<Partial*>(static_cast<char*>(workspace) + kFusedHeaderBytes);
  const long long n = static_cast<long long>(B) * HW;
  // Three launches, chained as programmatic dependents: [label counts] -> [one-pass kernel] -> [finalize].  The pre-pass is one
  // wave of blocks with a few 16-byte loads per thread, all in flight together (a latency-bound 8 MB read at 2x19x512x1024);
  // the one-pass kernel starts beside it and the finalize block is scheduled as the first one-pass block exits.
  const long long vecs = plbl_bytes == 8 ? n / 2 : n / 16;
  const int cgrid = static_cast<int>(std::max<long long>(
      1, std::min<long long>((vecs + 256 * 4 - 1) / (256 * 4), std::min(sm_count() * 4, kCountBlocksMax))));
  k_label_count<<<cgrid, 256, 0, st>>>(plbl, plbl_bytes, n, lc);
  HIAST_CHECK_LAUNCH();
  const int grid = std::min(loss_grid(n), sm_count() * 2);
  const size_t smem = loss_stage_bytes(C) + 7 * sizeof(double) * kThreadsL;     // + the forward accumulator columns
  HIAST_TRY(loss_configure_smem(C, loss_stage_bytes(C)));
  if (C == 19) {
    HIAST_TRY(ensure_dyn_smem(k_loss_fused_pk<19>, smem));
    HIAST_CUDA_TRY(launch_dependent(k_loss_fused_pk<19>, grid, kThreadsL, smem, st, a, lc, grad_weights, scales_used, grad_z, parts,
                                    cgrid));
  } else {
    HIAST_TRY(ensure_dyn_smem(k_loss_fused_pk<16>, smem));
    HIAST_CUDA_TRY(launch_dependent(k_loss_fused_pk<16>, grid, kThreadsL, smem, st, a, lc, grad_weights, scales_used, grad_z, parts,
                                    cgrid));
  }
  HIAST_CUDA_TRY(launch_dependent(k_loss_finalize, 1, kThreadsL, 0, st, parts, grid, sums, reinterpret_cast<long long*>(counts),
                                  losses, divisors, C, terms, term_weights));
  return HIAST_OK;
}

static int bwd_checked_launch(const float* z, const float* t, const void* plbl, int plbl_bytes, int B, int C, int64_t HW,
                              int region, int terms, const float* scales, const float* scales_used, float* grad_z,
                              const GoutArgs& go, void* stream) {
  const int rc = check_loss_args(z, t, plbl, plbl_bytes, B, C, HW, region, terms);
  if (rc != HIAST_OK) return rc;
  if ((!scales && !go.divisors) || !scales_used || !grad_z) return HIAST_ERR_INVALID_ARG;
  if (B == 0) return HIAST_OK;
  LossArgs a = {z, t, plbl, plbl_bytes, B, C, HW, region, terms};
  if (!loss_vector_ok(a, grad_z) || cst_kind_host(terms) != HIAST_CST_SOFTCE) return HIAST_ERR_UNSUPPORTED;
  cudaStream_t st = as_stream(stream);
  const int grid = std::min(loss_grid(static_cast<long long>(B) * HW), sm_count() * 2);
  const size_t smem = loss_stage_bytes(C);
  HIAST_TRY(loss_configure_smem(C, smem));
  if (C == 19) k_loss_bwd_pk<19><<<grid, kThreadsL, smem, st>>>(a, scales, grad_z, scales_used, go);
  else k_loss_bwd_pk<16><<<grid, kThreadsL, smem, st>>>(a, scales, grad_z, scales_used, go);
  HIAST_CHECK_LAUNCH();
  return HIAST_OK;
}

extern "C" int hiast_st_loss_bwd_checked(const float* z, const float* t, const void* plbl, int plbl_bytes, int B, int C,
                                         int64_t HW, int region, int terms, const float* scales, const float* scales_used,
                                         float* grad_z, void* stream) {
  if (!scales) return HIAST_ERR_INVALID_ARG;
  const GoutArgs none = {{nullptr, nullptr, nullptr, nullptr}, nullptr, nullptr, nullptr, 0, nullptr};
  return bwd_checked_launch(z, t, plbl, plbl_bytes, B, C, HW, region, terms, scales, scales_used, grad_z, none, stream);
}

extern "C" int hiast_st_loss_bwd_checked_terms(const float* z, const float* t, const void* plbl, int plbl_bytes, int B, int C,
                                               int64_t HW, int region, int terms, const float* gout_ce, const float* gout_kld,
                                               const float* gout_ent, const float* gout_cst, const double* divisors,
                                               const float* scales_used, float* grad_z, const float* term_weights,
                                               const float* hint_weights, int k0, float* upstream_out, void* stream) {
  if (!divisors || k0 < 0 || k0 > 3 || (upstream_out && !hint_weights)) return HIAST_ERR_INVALID_ARG;
  const GoutArgs go = {{gout_ce, gout_kld, gout_ent, gout_cst}, divisors, hint_weights, upstream_out, k0, term_weights};
  return bwd_checked_launch(z, t, plbl, plbl_bytes, B, C, HW, region, terms, nullptr, scales_used, grad_z, go, stream);
}
