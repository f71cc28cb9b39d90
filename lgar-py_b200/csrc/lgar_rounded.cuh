// =====================================================================================
// lgar_rounded.cuh -- exact result of k successive rounded additions (see below).
// Compiles for the device (CUDA intrinsics) and as plain C++ (bit casts) with the same IEEE
// operations, so tests/test_rounded_cpu.py checks on the host exactly what the kernels run.
// =====================================================================================
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#ifdef __CUDACC__
#define LGAR_RD_INLINE __device__ __forceinline__
#define LGAR_RD_NOINLINE __device__ __noinline__
#else
#define LGAR_RD_INLINE inline
#define LGAR_RD_NOINLINE inline
#endif

namespace lgar {

LGAR_RD_INLINE int f64_hi_word(double x) {
#ifdef __CUDA_ARCH__
  return __double2hiint(x);
#else
  uint64_t u; std::memcpy(&u, &x, 8); return (int)(u >> 32);
#endif
}
LGAR_RD_INLINE double f64_from_hi_word(int hi) {
#ifdef __CUDA_ARCH__
  return __hiloint2double(hi, 0);
#else
  const uint64_t u = (uint64_t)(uint32_t)hi << 32; double d; std::memcpy(&d, &u, 8); return d;
#endif
}

// ------------------------------------------------------------------------------------
// Exact result of k successive ROUNDED additions x = fl(x + s) (s may be negative), in
// O(number of binades crossed) instead of O(k).  Inside one binade every x is a multiple of
// ulp, so fl(x + s) - x is the same representable increment c for every step unless the
// discarded part of s is exactly half an ulp (tie -> round-to-even alternates; those steps are
// taken one by one).  Used by the root finder to jump along a monotone run of psi steps while
// visiting exactly the psi values the reference's `psi = psi +/- 0.1*factor` loop visits.
// ------------------------------------------------------------------------------------
LGAR_RD_INLINE int f64_exponent(double x) { return (f64_hi_word(x) >> 20) & 0x7ff; }
LGAR_RD_NOINLINE double advance_rounded_pos(double x, double s, long long k) {
  while (k > 0) {
    const double t = x + s;
    k--;
    if (k == 0 || !(t > 0.0)) return t;  // callers reject non-positive results
    const int e0 = f64_exponent(x), e1 = f64_exponent(t);
    const double c = t - x;    // exact (|s| << |x| in every caller)
    const double err = s - c;  // exact rounding residual of this step
    if (e0 != e1 || e1 <= 53 || e1 >= 0x7fe || !(t > 0.0)) {
      x = t;
      continue;
    }
    const double ulp = f64_from_hi_word((e1 - 52) << 20);
    if (fabs(err) * 2.0 == ulp || c == 0.0) {
      if (c == 0.0) return t;  // x + s == x from here on
      x = t;
      continue;
    }
    // steps that stay strictly inside the binade of t with the constant increment c
    const double lim = (s > 0.0) ? f64_from_hi_word((e1 + 1) << 20) : f64_from_hi_word(e1 << 20);
    const double room = (s > 0.0) ? (lim - t) : (t - lim);
    // common case: all remaining steps fit ((k + 1) |c| <= room, tested conservatively) -- no division
    if ((double)(k + 1) * fabs(c) * (1.0 + 0x1p-40) <= room) return fma((double)k, c, t);
    long long n = (long long)floor(room / fabs(c)) - 1;
    if (n > k) n = k;
    if (n < 0) n = 0;
    x = fma((double)n, c, t);  // exact: the result is a multiple of ulp inside the binade
    k -= n;
  }
  return x;
}

// round-to-nearest-even is symmetric under negation, so negative x mirror the positive case
LGAR_RD_INLINE double advance_rounded(double x, double s, long long k) {
  const bool neg = x < 0.0;
  const double r = advance_rounded_pos(neg ? -x : x, neg ? -s : s, k);
  return neg ? -r : r;
}

}  // namespace lgar
