// zz_fastmath.cuh — table-driven fp64 exp / log for the production sweep (host+device so the CPU test-suite can
// check them against libm without a GPU; tests/test_fastmath.py).
//
// Why: the faithful sweep spends ~4000 FP64-pipe instructions per gene-update, almost all inside CUDA libm's exp/log
// (profiles/r1_sweep_v0.md).  B200 issues 64 FP64 lanes/clk/SM, so the HBM roofline (188 B per gene-update) is only
// reachable below ~860 FP64 instructions per gene-update.  These versions cost 9 (exp) and 10 (log) FP64 instructions
// plus one shared-memory table read, with <= 2 ulp error (measured, tests/test_fastmath.py) — eight orders of magnitude
// inside the 1e-10 relative tolerance of north-star check 1.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "zz_tables.h"

#if defined(__CUDACC__)
#define ZZ_HD __host__ __device__ __forceinline__
#else
#define ZZ_HD inline
#endif

namespace zz {

ZZ_HD int d_hi(double x) {
#if defined(__CUDA_ARCH__)
  return __double2hiint(x);
#else
  uint64_t b; memcpy(&b, &x, 8); return (int)(b >> 32);
#endif
}
ZZ_HD int d_lo(double x) {
#if defined(__CUDA_ARCH__)
  return __double2loint(x);
#else
  uint64_t b; memcpy(&b, &x, 8); return (int)(uint32_t)b;
#endif
}
ZZ_HD double d_make(int hi, int lo) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double(hi, lo);
#else
  uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double x; memcpy(&x, &b, 8); return x;
#endif
}

// Polynomial coefficients.  sm_100a FP64 instructions encode a double immediate only through its high 32 bits, any other
// literal costs two UMOV (or an LDC) per use (profiles/r1_sweep_v1.md).  High-order coefficients are therefore rounded
// to 20 fraction bits (their rounding error is multiplied by r^4..r^6 and vanishes below 1e-19); the few that need full
// precision are kept in registers by the compiler.
#define ZZ_E4_IMM 0x1.55555p-5    /* 1/24  (error * r^4 < 1e-19) */
#define ZZ_L5_IMM 0x1.9999ap-3    /* 1/5   (error * r^5 < 1e-20 for |r| <= 2^-9) */
#define ZZ_P4_IMM 0x1.3b2abp-7    /* ln2^4/24 (error * r^4 < 1e-19 for |r| <= 2^-9) */
#define ZZ_Q4_IMM 0x1.3b2abp-39   /* the same coefficient for r in units of 1/256 (det_term) */

struct MathTabs {          // lives in shared memory on the device
  double exp_t[256];       // 2^(j/256) with the high word biased by -(j << 12): only meaningful through scale_tab()
  double2 log_t[512];      // {IC[i], LC[i]}
};

// 2^(n/256) = 2^k * 2^(j/256), n = 256 k + j: the table entry's high word carries -(j << 12), so adding (n << 12) =
// (k << 20) + (j << 12) restores the mantissa and adds k to the exponent field in ONE integer instruction, off the
// critical path of the polynomial (valid while the result is a normal number)
ZZ_HD double scale_tab(const double *__restrict__ exp_t, int n) {
#if defined(__CUDA_ARCH__)
  int j;
  asm("and.b32 %0, %1, 255;" : "=r"(j) : "r"(n));      // opaque to the optimiser: keeps (n & 255) separate so that the address
  const double tj = exp_t[j];                          // is ONE lea (base + (j << 3)) instead of shl + and + add
#else
  const double tj = exp_t[n & 255];
#endif
  return d_make(d_hi(tj) + (int)((unsigned)n << 12), d_lo(tj));
}

// exp(x) for -708 <= x <= 709 (callers clamp or take the libm path outside).  9 FP64 instructions.
ZZ_HD double fexp_core(double x, const double *__restrict__ exp_t) {
  const double magic = 6755399441055744.0;            // 1.5 * 2^52: the low word of (x*c + magic) is rint(x*c)
  const double t = fma(x, ZZ_KINVL, magic);
  const int n = d_lo(t);
  const double nd = t - magic;
  double r = fma(nd, -ZZ_KLHI, x);                    // exact: n has < 21 bits, L_hi has 32
  r = fma(nd, -ZZ_KLLO, r);
  double p = fma(r, ZZ_E4_IMM, ZZ_KE3);               // e^r = 1 + r (1 + r (1/2 + r (E3 + r/24))); E3 carries the economised r^5 term
  p = fma(r, p, 0.5);
  p = fma(r, p, 1.0);                                 // (the economised linear coefficient is 1 - 8.9e-15: taken as 1, error 1e-17)
  p = fma(r, p, 1.0);
  return scale_tab(exp_t, n) * p;                     // 2^k T[j] is exact, the product is normal here
}

// libm fallbacks, kept out of line so the (rare) special cases do not bloat the unrolled loops
#if defined(__CUDACC__)
__device__ __noinline__ double slow_exp(double x) { return exp(x); }
__device__ __noinline__ double slow_log(double x) { return log(x); }
#endif
ZZ_HD double libm_exp(double x) {
#if defined(__CUDA_ARCH__)
  return slow_exp(x);
#else
  return exp(x);
#endif
}
ZZ_HD double libm_log(double x) {
#if defined(__CUDA_ARCH__)
  return slow_log(x);
#else
  return log(x);
#endif
}

// general exp: libm outside the range where the exponent-field trick is valid (rare, warp-divergent)
ZZ_HD double fexp(double x, const double *__restrict__ exp_t) {
  if ((unsigned)(d_hi(x) & 0x7FFFFFFF) > 0x40862000u) return libm_exp(x);   // |x| > 708 (or NaN): integer test, no FP64 slot
  return fexp_core(x, exp_t);
}

// exp(-a) for the detection probability p = 1 - exp(-a), a >= 0: for a > 40, exp(-a) < 2^-57 and 1 - exp(-a) rounds to
// 1 whatever its value, so a is clamped at 700 through an integer min on the high word (a >= 0: high words order like
// the values; +Inf clamps too; NaN cannot pass through here unnoticed because (x - y)^2 is NaN as well).
ZZ_HD double fexp_neg_clamped(double a, const double *__restrict__ exp_t) {
  const int hi = d_hi(a);
  const int hc = hi < 0x4085E000 ? hi : 0x4085E000;   // 0x4085E000_00000000 = 700.0; the low word is kept (< 700.0000002)
  return fexp_core(-d_make(hc, d_lo(a)), exp_t);
}

// 2^(-b) for b >= 0 (b = t * alpha_r * log2(e), the detection exponent in base 2): the reduction step nd/256 is exact, so
// one fma fewer than the natural-base version.  b is clamped at 1010 (2^-1010 is still a normal number and 1 - 2^-b == 1
// for every b > 54).  8 FP64 instructions.
ZZ_HD double fexp2_neg_clamped(double b, const double *__restrict__ exp_t) {
  const int hi = d_hi(b);
  const int hc = hi < 0x408F9000 ? hi : 0x408F9000;   // 0x408F9000_00000000 = 1010.0
  const double x = -d_make(hc, d_lo(b));
  const double magic = 6755399441055744.0;
  const double t = fma(x, 256.0, magic);
  const int n = d_lo(t);
  const double nd = t - magic;
  const double r = fma(nd, -0.00390625, x);           // exact, |r| <= 2^-9 (log2 units)
  double p = fma(r, ZZ_P4_IMM, ZZ_KP3);               // 2^r = 1 + r (P1 + r (P2 + r (P3 + r P4)))
  p = fma(r, p, ZZ_KP2);
  p = fma(r, p, ZZ_KP1);
  p = fma(r, p, 1.0);
  return scale_tab(exp_t, n) * p;
}

// log(x) for positive normal finite x.  8 FP64 instructions + one int->double conversion.
ZZ_HD double flog_core(double x, const double2 *__restrict__ log_t) {
  const int hi = d_hi(x);
  const int k = ((hi + 0x00096000) >> 20) - 1023;     // exponent, +1 when the mantissa is >= 1 + 212/512 (table entry is for m/2)
  const int i = (hi >> 11) & 511;
  const double m = d_make((hi & 0x000FFFFF) | 0x3FF00000, d_lo(x));
  const double2 tc = log_t[i];
  const double r = fma(m, tc.x, -1.0);                // exact (IC has <= 11 significant bits), |r| <= 2^-9
  const double r2 = r * r;
  double s = fma(r, ZZ_L5_IMM, -0.25);                // log1p(r) = r + r^2 (-1/2 + r/3 - r^2/4 + r^3/5), |r^6/6| < 1e-17
  s = fma(r, s, 1.0 / 3);
  s = fma(r, s, -0.5);
  const double lo = fma(r2, s, r);
  return fma((double)k, ZZ_KLN2, tc.y) + lo;           // (double)k: I2F on the otherwise idle XU pipe
}

#if defined(__CUDACC__)
// 0x3FF00000 read back from shared memory: a value neither the compiler nor ptxas can fold into an immediate, so it
// stays in a register for the whole tile (see det_term)
__device__ __forceinline__ int opaque_one_hi(const volatile int *slot) { return *slot; }
#endif

// One library's detection factor for the straight-line tile:  q = undetected ? 1 - p : p,  p = 1 - 2^(-t a),
// a = alpha_r log2(e), passed as am256 = -256 a.  The caller clamps t so that t a <= 1010 for every library (then 2^(-t a)
// is a normal number; beyond 54 the value no longer matters because p rounds to 1).  q is +0 or a multiple of 2^-53 in
// [2^-53, 1].  The caller multiplies the factors of all libraries and takes ONE logarithm of the product:
//   A(y) = sum_r log q_r = log prod_r q_r.
// With <= 18 libraries the product of non-zero factors is >= 2^-954, still a normal number, and each multiplication adds
// <= 2^-53 relative error to the product, i.e. <= 1.1e-16 absolute error to A — the size of the rounding of a single
// log q_r in the per-library form (the reference's own log(1 - p_x) carries the same absolute error through 1 - p_x).
// 10 FP64 instructions, ~6 integer/select, 1 shared-memory read per library (the per-library logarithm cost 8 more FP64,
// ~6 integer and a second table read).
ZZ_HD double det_q(double t, double am256, unsigned undetected, const MathTabs &T) {
  const double magic = 6755399441055744.0;
  const double w = fma(t, am256, magic);               // low word: n = rint(-256 t a)
  const int n = d_lo(w);
  const double nd = w - magic;
  const double r = fma(t, am256, -nd);                 // -256 t a - n from the exact product: |r| <= 1/2 (units of 1/256)
  double p = fma(r, ZZ_Q4_IMM, ZZ_KQ3);                // 2^(r/256) = 1 + r (Q1 + r (Q2 + r (Q3 + r Q4)))
  p = fma(r, p, ZZ_KQ2);
  p = fma(r, p, ZZ_KQ1);
  p = fma(r, p, 1.0);
  const double e = scale_tab(T.exp_t, n) * p;          // 2^(-t a)
  double q = 1.0 - e;                                  // p_x (U:157)
  if (undetected) q = 1.0 - q;                         // 1 - p_x as the reference forms it (P:84), not e itself
  return q;
}

// log of a product of det_q factors: -Inf when some factor was +0 (the product of <= 18 non-zero factors is >= 2^-954)
ZZ_HD double det_log_product(double P, const double2 *__restrict__ log_t) {
  const double l = flog_core(P, log_t);
  return (d_hi(P) < 0x00100000) ? -INFINITY : l;       // P == +0 (or subnormal, which no legitimate product is)
}

ZZ_HD double flog(double x, const double2 *__restrict__ log_t) {
  const unsigned hi = (unsigned)d_hi(x);
  if (hi - 0x00100000u >= 0x7FE00000u) return libm_log(x);  // x <= 0 (incl. -0), subnormal, Inf, NaN: libm semantics
  return flog_core(x, log_t);
}

// log(q) for q known to be +0 or a normal number in (0, 1] (a detection probability or its complement): branch-free
ZZ_HD double flog_prob(double q, const double2 *__restrict__ log_t) {
  const double l = flog_core(q, log_t);
  return ((d_hi(q) | d_lo(q)) == 0) ? -INFINITY : l;
}

// ---- branch-free variants (the production tile is one straight-line block so that the scheduler can overlap its chains) ----

// exp(x) without a libm path: |x| is clamped at 708 through the high word, and x < -745.2 (true result 0, below the smallest
// subnormal) returns exactly 0.  Between -745.2 and -708 libm would return a subnormal; this returns exp(-708) = 3.3e-308.
ZZ_HD double fexp_nb(double x, const double *__restrict__ exp_t) {
  const int hi = d_hi(x);
  const int mag = hi & 0x7FFFFFFF;
  const int magc = mag < 0x40862000 ? mag : 0x40862000;            // 708.0
  const double r = fexp_core(d_make(magc | (hi & 0x80000000), d_lo(x)), exp_t);
  return ((unsigned)hi > 0xC0874999u) ? 0.0 : r;                   // x < -745.2 (negative: larger bit pattern = more negative)
}

// log(x) for x that the caller knows to be a positive normal number (no libm path; anything else gives finite garbage)
ZZ_HD double flog_nb(double x, const double2 *__restrict__ log_t) { return flog_core(x, log_t); }

// true when x is a positive normal finite double (integer test on the high word)
ZZ_HD bool is_pos_normal(double x) { return (unsigned)d_hi(x) - 0x00100000u < 0x7FE00000u; }

// 1/x for positive normal x: split off the exponent, float reciprocal seed of the mantissa, two Newton steps (<= 1 ulp)
ZZ_HD double frcp_nb(double x) {
  const int hi = d_hi(x);
  const double m = d_make((hi & 0x000FFFFF) | 0x3FF00000, d_lo(x));   // [1, 2)
#if defined(__CUDA_ARCH__)
  double r = (double)__frcp_rn((float)m);
#else
  double r = (double)(1.0f / (float)m);
#endif
  r = r * fma(-m, r, 2.0);
  r = r * fma(-m, r, 2.0);
  r = fma(r, fma(-m, r, 1.0), r);                                   // final correction: r + r (1 - m r)
  return d_make(d_hi(r) + ((1023 - ((hi >> 20) & 0x7FF)) << 20), d_lo(r));
}

// sqrt(x) for x in the float range (here x = -2 log u in (4.6e-10, 44.4]): float rsqrt seed + two Newton steps
ZZ_HD double fsqrt_nb(double x) {
#if defined(__CUDA_ARCH__)
  double y = (double)rsqrtf((float)x);
#else
  double y = (double)(1.0f / sqrtf((float)x));
#endif
  const double h = 0.5 * x;
  y = fma(y, fma(-h, y * y, 0.5), y);
  y = fma(y, fma(-h, y * y, 0.5), y);
  const double s = x * y;
  return fma(fma(-s, s, x), 0.5 * y, s);                            // one correction step on the root itself
}

// (w + 0.5) * 2^-32 in one fma (both products are exact, so this equals the two-step form bit for bit)
ZZ_HD double u32_to_open01(uint32_t w) { return fma((double)w, 0x1p-32, 0x1p-33); }

}  // namespace zz
