// =====================================================================================
// lgar_pow.cuh -- table-driven double-double pow(x, y) for the LGAR kernels.
//
// Why not the CUDA math library's pow: (a) accuracy -- libdevice's pow is a <= 2 ulp routine,
// the reference runs on glibc's ~0.52 ulp pow, and the LGAR root finders amplify ulp-level
// differences (measured: 2.9e-9 relative on per-step AET over a year, above the 1e-9 bar);
// (b) cost -- libdevice's pow occupies ~150 FP64-pipe issue slots and has a 926-cycle dependent
// latency on B200 (tools/fp64_microbench.cu); this routine needs ~55 FP64 operations, two small
// table look-ups (6 KB, L1-resident), and inlines so that independent evaluations overlap.
//
// Algorithm (tables from tools/gen_pow_tables.py, mpmath 200 bit):
//   log:  x = 2^k z; i = sub-interval of z (128 in mantissa-bit space, 1.0 at the centre of
//         interval 64); r = z*invc_i - 1 as an exact two-term sum (p = fl(z*invc), e = fma error);
//         log x = k ln2 + log c_i + log1p(r) evaluated as hi + lo with |err| ~ 2^-66 relative.
//   mul:  ehi + elo = y * (hi + lo)      (fma for the exact product error)
//   exp:  exp(ehi + elo) = 2^(k'/128) * (1 + tail + r + r^2/2 + ...), k' = rint(ehi*128/ln2)
// Fast path: x positive normal, result comfortably inside the normal range; everything else
// (zero, negative, inf, NaN, subnormal, overflow/underflow) goes to the library pow so all
// special-value semantics are the library's.
//
// The same source compiles as plain C++ for the CPU accuracy harness (tools/pow_accuracy.cpp):
// every operation is an IEEE fp64 add/mul/fma, so CPU and GPU results are bit-identical.
// =====================================================================================
#pragma once
#include <cstdint>
#include <cstring>
#include <cmath>
#include "lgar_pow_tables.h"

#ifdef __CUDACC__
#define LGAR_HD __host__ __device__ __forceinline__
#else
#define LGAR_HD inline
#endif

namespace lgar {

struct PowLogEntry { double invc, logc, logctail, pad; };
struct PowExpEntry { double t, tail; };

#ifdef __CUDACC__
__device__ const PowLogEntry g_pow_log_table[LGAR_POW_N] = {LGAR_POW_LOG_TABLE};
__device__ const PowExpEntry g_pow_exp_table[LGAR_POW_N] = {LGAR_POW_EXP_TABLE};
#endif
static const PowLogEntry h_pow_log_table[LGAR_POW_N] = {LGAR_POW_LOG_TABLE};
static const PowExpEntry h_pow_exp_table[LGAR_POW_N] = {LGAR_POW_EXP_TABLE};

// Scalar constants.  On the device they live in __constant__ memory so that FP64 instructions take them as
// constant-bank operands instead of materialising each 64-bit immediate with two moves.
enum PowConst { PC_LN2HI, PC_LN2LO, PC_INVLN2N, PC_LN2N_HI, PC_LN2N_LO, PC_A3, PC_A4, PC_A5, PC_A6, PC_A7, PC_A8, PC_A9,
                PC_C2, PC_C3, PC_C4, PC_C5, PC_C6, PC_SHIFT, PC_COUNT };
#define LGAR_POW_CONSTS {LGAR_LN2HI, LGAR_LN2LO, LGAR_INVLN2N, -LGAR_LN2N_HI, -LGAR_LN2N_LO, LGAR_LOG_A3, LGAR_LOG_A4, \
                         LGAR_LOG_A5, LGAR_LOG_A6, LGAR_LOG_A7, LGAR_LOG_A8, LGAR_LOG_A9, LGAR_EXP_C2, LGAR_EXP_C3, \
                         LGAR_EXP_C4, LGAR_EXP_C5, LGAR_EXP_C6, 0x1.8p52}
#ifdef __CUDACC__
__constant__ double c_pow_consts[PC_COUNT] = LGAR_POW_CONSTS;
#endif
static const double h_pow_consts[PC_COUNT] = LGAR_POW_CONSTS;
#ifdef __CUDA_ARCH__
#define PCK(i) c_pow_consts[i]
#else
#define PCK(i) h_pow_consts[i]
#endif

LGAR_HD uint64_t pow_bits(double x) {
#ifdef __CUDA_ARCH__
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t u; std::memcpy(&u, &x, 8); return u;
#endif
}
LGAR_HD double pow_from_bits(uint64_t u) {
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)u);
#else
  double d; std::memcpy(&d, &u, 8); return d;
#endif
}
LGAR_HD double pow_fma(double a, double b, double c) {
#ifdef __CUDA_ARCH__
  return __fma_rn(a, b, c);
#else
  return std::fma(a, b, c);
#endif
}

// Branch-free core.  Returns the fast-path value and sets ok = false when the arguments are outside
// the fast path (the caller then uses the library pow).  Being branch-free, two or more calls in one
// basic block are interleaved by the scheduler (ILP): see pow_x2 in lgar_device.cuh.
LGAR_HD double pow_core(double x, double y, bool& ok) {
  const uint64_t ix = pow_bits(x);
  // x must be a positive normal number
  const bool x_ok = (ix - 0x0010000000000000ULL) < (0x7ff0000000000000ULL - 0x0010000000000000ULL);
  // ---- log(x) = hi + lo
  const uint64_t tmp = ix - LGAR_POW_OFF;
  const int i = (int)((tmp >> 45) & (LGAR_POW_N - 1));
  const int k = (int)((int64_t)tmp >> 52);
  const double z = pow_from_bits(ix - (tmp & 0xfff0000000000000ULL));
  const double kd = (double)k;
#ifdef __CUDA_ARCH__
  const double2 le0 = __ldg(reinterpret_cast<const double2*>(g_pow_log_table) + 2 * i);
  const double2 le1 = __ldg(reinterpret_cast<const double2*>(g_pow_log_table) + 2 * i + 1);
  const double invc = le0.x, logc = le0.y, logctail = le1.x;
#else
  const double invc = h_pow_log_table[i].invc, logc = h_pow_log_table[i].logc, logctail = h_pow_log_table[i].logctail;
#endif
  const double p = z * invc;
  const double rl = pow_fma(z, invc, -p);  // exact: z*invc = p + rl
  const double r = p - 1.0;                // exact (Sterbenz)
  const double t1 = pow_fma(kd, PCK(PC_LN2HI), logc);  // exact: both are multiples of 2^-42 below 2^10
  // Fast2Sum t1 + r: exact because t1 == 0 (interval 64, k = 0) or |t1| >= |log c_63| = 0.0059 > |r|
  const double t2 = t1 + r;
  const double lo2 = (t1 - t2) + r;
  const double lo1 = pow_fma(kd, PCK(PC_LN2LO), logctail);
  // -r^2/2 in two terms
  const double ar = -0.5 * r;
  const double ar2 = r * ar;
  const double lo3 = pow_fma(ar, r, -ar2);
  const double hi = t2 + ar2;
  const double lo4 = (t2 - hi) + ar2;
  const double r2 = r * r;
  // log1p(r) - r + r^2/2 = r^3 (A3 + A4 r + ... + A9 r^6): |r| <= 0.0046, truncation < 2^-70 relative
  double q = pow_fma(r, PCK(PC_A9), PCK(PC_A8));
  q = pow_fma(r, q, PCK(PC_A7));
  q = pow_fma(r, q, PCK(PC_A6));
  q = pow_fma(r, q, PCK(PC_A5));
  q = pow_fma(r, q, PCK(PC_A4));
  q = pow_fma(r, q, PCK(PC_A3));
  const double pl = (r2 * r) * q;
  // contribution of the product rounding error rl: rl * d/dr log1p(r) = rl (1 - r + r^2)
  const double lo5 = pow_fma(rl, r2 - r, rl);
  const double lo = ((lo1 + lo2) + (lo3 + lo4)) + (pl + lo5);
  const double lhi = hi + lo;
  const double llo = (hi - lhi) + lo;
  // ---- ehi + elo = y * log(x)
  const double ehi = y * lhi;
  const double elo = pow_fma(y, llo, pow_fma(y, lhi, -ehi));
  // result must stay well inside the normal range (also rejects NaN / inf in y)
  const double aeh = ehi < 0.0 ? -ehi : ehi;
  ok = x_ok && (aeh < 700.0);
  // ---- exp(ehi + elo)
  const double zz = ehi * PCK(PC_INVLN2N);
  const double kf = (zz + PCK(PC_SHIFT)) - PCK(PC_SHIFT);  // rint(zz), |zz| < 2^17
  const int ki = ok ? (int)kf : 0;
  double rr = pow_fma(kf, PCK(PC_LN2N_HI), ehi);   // exact: kf*LN2N_HI has <= 52 bits
  rr = pow_fma(kf, PCK(PC_LN2N_LO), rr);
  rr = rr + elo;
  const int j = (int)(ki & (LGAR_POW_N - 1));
#ifdef __CUDA_ARCH__
  const double2 ee = __ldg(reinterpret_cast<const double2*>(g_pow_exp_table) + j);
  const double tj = ee.x, tailj = ee.y;
#else
  const double tj = h_pow_exp_table[j].t, tailj = h_pow_exp_table[j].tail;
#endif
  const uint64_t sbits = pow_bits(tj) + ((uint64_t)(int64_t)(ki >> 7) << 52);
  const double scale = pow_from_bits(sbits);
  const double rr2 = rr * rr;
  double e2 = pow_fma(rr, PCK(PC_C3), PCK(PC_C2));
  double e4 = pow_fma(rr, PCK(PC_C5), PCK(PC_C4));
  e4 = pow_fma(rr2, PCK(PC_C6), e4);
  const double tmp2 = tailj + (rr + (rr2 * e2 + (rr2 * rr2) * e4));
  const double res = pow_fma(scale, tmp2, scale);
  // |y log x| tiny: pow = 1 + y log x to well below half an ulp
  return (aeh < 0x1p-60) ? (1.0 + ehi) : res;
}

// returns true and sets *out on the fast path; false -> caller must use the library pow
LGAR_HD bool pow_fast(double x, double y, double* out) {
  bool ok;
  *out = pow_core(x, y, ok);
  return ok;
}

}  // namespace lgar
