// Exact arithmetic of one IAS threshold update (one class, one group), shared by the device
// scan kernel and the host test hook so that the recipe can be checked against numpy on CPU.
//
// Reference: workflows/pseudo_label_generator.py:171-179 (get_ias_threshold -> np.quantile,
// method 'linear'), :198-201 (sample list = [thr] + fp16 confidences), :207-209 (EMA + clamp).
// numpy 2.x _quantile/_get_indexes/_lerp semantics (see oracle/ias.py:threshold_from_hist).
//
// Every floating-point operation here is an individually rounded IEEE operation: on the device
// they are written with __dmul_rn/__dadd_rn/... so nvcc cannot contract them into FMAs; on the
// host the file is compiled with -ffp-contract=off.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDA_ARCH__)
#define HIAST_HD __host__ __device__ __forceinline__
#define HIAST_DMUL(a, b) __dmul_rn((a), (b))
#define HIAST_DADD(a, b) __dadd_rn((a), (b))
#define HIAST_DSUB(a, b) __dsub_rn((a), (b))
#define HIAST_FMA(a, b, c) __fma_rn((a), (b), (c))
#define HIAST_FMULF(a, b) __fmul_rn((a), (b))
#else
#if defined(__CUDACC__)
#define HIAST_HD __host__ __device__ inline
#else
#define HIAST_HD inline
#endif
#define HIAST_DMUL(a, b) ((a) * (b))
#define HIAST_DADD(a, b) ((a) + (b))
#define HIAST_DSUB(a, b) ((a) - (b))
#define HIAST_FMA(a, b, c) fma((a), (b), (c))
#define HIAST_FMULF(a, b) ((a) * (b))
#endif

namespace hiast {

// Value of a non-negative, finite fp16 bit pattern as a double (exact), by bit manipulation.
HIAST_HD double half_bits_to_double(unsigned bits) {
  const unsigned e = (bits >> 10) & 0x1fu;
  const unsigned m = bits & 0x3ffu;
  if (e == 0) return static_cast<double>(m) * 5.9604644775390625e-08;  // subnormal: m * 2^-24
  const unsigned long long d = (static_cast<unsigned long long>(e + (1023 - 15)) << 52) |
                               (static_cast<unsigned long long>(m) << 42);
#if defined(__CUDA_ARCH__)
  return __longlong_as_double(static_cast<long long>(d));
#else
  double r;
  memcpy(&r, &d, sizeof(r));
  return r;
#endif
}

struct dd {
  double hi, lo;
};

HIAST_HD dd dd_mul(dd a, dd b) {
  const double p = HIAST_DMUL(a.hi, b.hi);
  double e = HIAST_FMA(a.hi, b.hi, -p);
  e = HIAST_DADD(e, HIAST_DADD(HIAST_DMUL(a.hi, b.lo), HIAST_DMUL(a.lo, b.hi)));
  dd r;
  r.hi = HIAST_DADD(p, e);
  r.lo = HIAST_DSUB(e, HIAST_DSUB(r.hi, p));
  return r;
}

// x^n, n >= 1, in double-double (~104 bits), rounded once to double = the correctly rounded power
// (tests compare with exact rational arithmetic).  glibc's pow, which numpy's float64 scalar **
// calls, is within 1 ulp of that and differs from it in ~0.08 % of arguments; the step below
// certifies that the float32 quantile does not depend on that last bit (err bit 2).
HIAST_HD double powi_dd(double x, int n) {
  dd base = {x, 0.0};
  dd acc = {1.0, 0.0};
  bool have = false;
  while (n > 0) {
    if (n & 1) {
      acc = have ? dd_mul(acc, base) : base;
      have = true;
    }
    n >>= 1;
    if (n) base = dd_mul(base, base);
  }
  return HIAST_DADD(acc.hi, acc.lo);
}

// x^8 by three double-double squarings (gamma = 8 in every shipped config).
HIAST_HD double pow8_dd(double x) {
  dd a;
  a.hi = HIAST_DMUL(x, x);
  a.lo = HIAST_FMA(x, x, -a.hi);
  a = dd_mul(a, a);
  a = dd_mul(a, a);
  return HIAST_DADD(a.hi, a.lo);
}

HIAST_HD double ias_pow(double thr, double gamma) {
  if (gamma == 8.0) return pow8_dd(thr);
  const int gi = static_cast<int>(gamma);
  if (static_cast<double>(gi) == gamma && gi >= 1 && gi <= 64) return powi_dd(thr, gi);
  if (gamma == 0.0) return 1.0;
  return pow(thr, gamma);  // non-integer gamma: CUDA pow (<= 2 ulp) / glibc pow on the host
}

// smallest bin b in [0, nb) with prefix[b] > j   (prefix = inclusive prefix sums, prefix[nb-1] > j)
template <typename P>
HIAST_HD int upper_bin(const P* prefix, int nb, long long j) {
  int lo = 0, hi = nb - 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (static_cast<long long>(prefix[mid]) > j) hi = mid; else lo = mid + 1;
  }
  return lo;
}

// Serial searcher (host hook, single-thread device use).
template <typename P>
struct SerialSearch {
  const P* prefix;
  int nb;
  HIAST_HD int operator()(long long j) const { return upper_bin(prefix, nb, j); }
};

// number of bins (keys key_lo + 0 .. key_lo + nb - 1) whose fp16 value is < thr: the smallest fp16 key whose
// value is >= thr, found directly from the bits of thr (truncate to fp16, step up if that lost anything).
HIAST_HD int bins_below(int key_lo, int nb, double thr) {
  if (!(thr > 0.0)) return 0;                       // every key >= 0 has value >= thr
  if (thr > 65504.0) return nb;
  long long db;
#if defined(__CUDA_ARCH__)
  db = __double_as_longlong(thr);
#else
  memcpy(&db, &thr, sizeof(db));
#endif
  const int e = static_cast<int>((db >> 52) & 0x7ff) - 1023;     // unbiased exponent of thr
  int key;                                                       // largest key with value <= thr
  if (e >= -14) {
    key = ((e + 15) << 10) | static_cast<int>((db >> 42) & 0x3ff);
  } else {
    key = static_cast<int>(thr * 16777216.0);                   // subnormal half: floor(thr * 2^24), exact product
  }
  if (half_bits_to_double(static_cast<unsigned>(key)) < thr) key += 1;
  int b = key - key_lo;
  if (b < 0) b = 0;
  if (b > nb) b = nb;
  return b;
}

HIAST_HD double bits_step(double x, int dir) {  // next representable double above (dir>0) / below a positive x
#if defined(__CUDA_ARCH__)
  return __longlong_as_double(__double_as_longlong(x) + dir);
#else
  long long b;
  memcpy(&b, &x, sizeof(b));
  b += dir;
  memcpy(&x, &b, sizeof(b));
  return x;
#endif
}

struct QuantilePos {
  double vi, fl, g;
};

HIAST_HD QuantilePos quantile_pos(long long m, double alpha, double p) {
  QuantilePos r;
  const double q = HIAST_DSUB(1.0, HIAST_DMUL(alpha, p));
  r.vi = HIAST_DMUL(static_cast<double>(m), q);  // (n-1)*q with n = m+1
  r.fl = floor(r.vi);
  r.g = HIAST_DSUB(r.vi, r.fl);
  return r;
}

HIAST_HD double lerp_np(double a, double b, double g) {  // numpy >= 1.22 _lerp
  const double d = HIAST_DSUB(b, a);
  if (g >= 0.5) return HIAST_DSUB(b, HIAST_DMUL(d, HIAST_DSUB(1.0, g)));
  return HIAST_DADD(a, HIAST_DMUL(d, g));
}

// One update.  prefix: inclusive prefix sums of the class's key histogram (nb bins from key_lo).
// Returns the new threshold; *temp_out = float32 quantile.
//   *err |= 1  if q is outside [0,1] (numpy raises ValueError there);
//   *err |= 2  if the float32 quantile would change were thr^gamma one ulp larger or smaller, i.e.
//              the result is not certified independent of the last-bit rounding of the host's pow()
//              (glibc's pow is within 1 ulp but not always correctly rounded; this implementation
//              is correctly rounded for integer gamma).  Never observed on real data; see DESIGN.md.
template <typename P, typename Search>
HIAST_HD double ias_threshold_step(const P* prefix, int nb, int key_lo, double thr,
                                   double alpha, double beta, double gamma,
                                   float* temp_out, int* err, const Search& search) {
  const long long m = static_cast<long long>(prefix[nb - 1]);
  const double p = ias_pow(thr, gamma);
  const double q = HIAST_DSUB(1.0, HIAST_DMUL(alpha, p));
  if (!(q >= 0.0 && q <= 1.0)) *err |= 1;
  double t64;
  if (m == 0) {
    t64 = thr;  // the list is [thr] alone: every quantile is thr (a + 0*g)
  } else {
    const QuantilePos pos = quantile_pos(m, alpha, p);
    long long lo_i, hi_i;
    if (pos.vi >= static_cast<double>(m)) {
      lo_i = hi_i = m;
    } else if (pos.vi < 0.0) {
      lo_i = hi_i = 0;
    } else {
      lo_i = static_cast<long long>(pos.fl);
      hi_i = lo_i + 1;
    }
    const int kb = bins_below(key_lo, nb, thr);
    const long long r = kb > 0 ? static_cast<long long>(prefix[kb - 1]) : 0;  // rank of thr in the merged list
    double a, b;
    int bin_a = -1;
    if (lo_i == r) a = thr;
    else {
      bin_a = search(lo_i < r ? lo_i : lo_i - 1);
      a = half_bits_to_double(static_cast<unsigned>(key_lo + bin_a));
    }
    if (hi_i == lo_i) b = a;
    else if (hi_i == r) b = thr;
    else {
      const long long j = hi_i < r ? hi_i : hi_i - 1;
      // the next order statistic usually sits in the same bin as the previous one
      if (bin_a >= 0 && static_cast<long long>(prefix[bin_a]) > j) b = a;
      else b = half_bits_to_double(static_cast<unsigned>(key_lo + search(j)));
    }
    t64 = lerp_np(a, b, pos.g);
    if (p > 0.0 && alpha != 0.0 && gamma != 1.0) {
      // certificate: same order statistics and same float32 result for p -/+ 1 ulp
      const QuantilePos plo = quantile_pos(m, alpha, bits_step(p, -1));
      const QuantilePos phi = quantile_pos(m, alpha, bits_step(p, +1));
      bool fragile = (plo.fl != pos.fl) || (phi.fl != pos.fl);
      if (!fragile && a != b) {
        const float t = static_cast<float>(t64);
        fragile = (static_cast<float>(lerp_np(a, b, plo.g)) != t) || (static_cast<float>(lerp_np(a, b, phi.g)) != t);
      }
      if (fragile) *err |= 2;
    }
  }
  const float temp = static_cast<float>(t64);                 // stored into a float32 array (:175)
  *temp_out = temp;
  const float one_minus_beta = static_cast<float>(1.0 - beta);  // python float weak-cast to f32 (NEP 50)
  const float prod = HIAST_FMULF(one_minus_beta, temp);
  double nt = HIAST_DADD(HIAST_DMUL(beta, thr), static_cast<double>(prod));
  if (nt >= 1.0) nt = 0.999;
  return nt;
}

template <typename P>
HIAST_HD double ias_threshold_step(const P* prefix, int nb, int key_lo, double thr,
                                   double alpha, double beta, double gamma,
                                   float* temp_out, int* err) {
  const SerialSearch<P> search = {prefix, nb};
  return ias_threshold_step(prefix, nb, key_lo, thr, alpha, beta, gamma, temp_out, err, search);
}

}  // namespace hiast
