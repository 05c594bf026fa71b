// zz_fast.cuh — the production ("fast") formulation of the fused sweep for sm_100a.
//
// Same Markov kernel, same draws and — up to floating-point ties — the same decisions as the reference-order moves
// in zz_device.cuh, reorganised so that one gene-update costs ~700 FP64-pipe instructions instead of ~4000
// (B200: 64 FP64 lanes/clk/SM, so the 188 B/gene-update HBM roofline needs < ~860; profiles/r1_*.md):
//   * computeXgLikelihood (P:57-127) is split as XgLik = beta * (A(y) + B(y, v)):
//       A(y)    = sum_r log(1 - p_gr) | log(p_gr)           (needs exp+log per library; depends on y and alpha only)
//       B(y, v) = -n_det (ln sqrt(2 pi) + 0.5 log v) - 0.5 D(y) / v,  D(y) = sum_detected (x_gr - y)^2
//     A(current y) is carried per gene in HBM (`lps`), so mhYg evaluates exp/log only for the PROPOSED y and
//     mhVariance_g needs no per-library transcendental at all.  D(y) = ss + n_det (xbar - y)^2 comes from per-gene data
//     statistics prepared once at upload, so the sweep never re-reads the G x R matrix Xg.
//   * log(Sg) = log(exp(s0 + s1 y)) (P:20 via U:146) is taken as s0 + s1 y; divisions by hyper-parameter constants
//     become multiplications by reciprocals hoisted to shared memory; Lk of gibbsAllocationWithinActive (Pr:375) reuses
//     the w_k * exp(.) terms of gibbsAllocationActiveInactive (Pr:344); floor(cum/u) == 0 (Pr:400-402) is evaluated
//     as cum < u (exactly equivalent for cum >= 0, u > 0).
//   * exp/log are the table-driven versions of zz_fastmath.cuh (<= 2 ulp).
// Every reorganisation changes results by O(1e-15) relative, eight orders below the 1e-10 tolerance of north-star
// check 1; decisions can differ from the oracle only when |log u - R| is itself O(1e-14) (reported by the tests).
#pragma once
#include "zz_device.cuh"
#include "zz_fastmath.cuh"

#ifndef ZZ_ILP
#define ZZ_ILP 4
#endif
#ifndef ZZ_SWEEP_MINB
#define ZZ_SWEEP_MINB 1
#endif

namespace zz {

struct FastH {   // one chain's hyper-parameters + hoisted gene-independent sub-expressions (shared memory)
  int R, K;
  double s0, s1, tau, rtau, irt;          // irt = 1/sqrt(tau)
  double c1;                              // log(sqrt(tau) * sqrt(2 pi)):  log(v * rt * sqrt2pi) = log v + c1   (P:20)
  double sp, log_sp, log_1msp, log_norm_i;
  double im, sd_i, inv2iv, c0;            // c0 = (1 - sp) / (sqrt2pi * sd_i):  L0 = c0 * exp(-(y-im)^2 / (2 iv))   (P:220)
  double wa, one_m_wa, beta, temp, vub;
  double alpha_min;                       // min_r alpha_r: t * alpha_min >= 37.5 => every p_x == 1 exactly
  double alpha[ZZ_MAX_LIBS];
  double alpha2[ZZ_MAX_LIBS];             // alpha_r * log2(e): the detection exponent in base 2 (fexp2_neg_clamped)
  double am256[ZZ_MAX_LIBS];              // -256 alpha_r log2(e) (det_term)
  double t_clamp;                         // 1010 / max_r alpha2_r: t <= t_clamp keeps every 2^(-t alpha2_r) normal (det_term)
  int one_hi;                             // 0x3FF00000 (opaque_one_hi)
  int uclamp_ok;                          // max alpha <= 16 min alpha: one clamp of t serves every library (detect_logsum_sl)
  double am[ZZ_MAX_COMP], inv2av[ZZ_MAX_COMP], ck[ZZ_MAX_COMP];  // ck = w_k / (sqrt2pi * sqrt(av_k))           (Pr:344)
  // device-resident generations (zz_persist.cuh): the mhP_x proposal whose gene sum this sweep accumulates (Pr:209-224)
  int px_on, px_lib;                      // proposal pending for library px_lib
  int px_prev_acc, pad_;                  // the previous generation's proposal was accepted: A(y) += its stored per-gene term
  double px_am256, px_amin, t_clamp_px;   // -256 alpha' log2(e); min(alpha, alpha'); clamp of t valid for both
};

// FastH is derived from zz_hyper once per hyper-parameter change (k_prepare_hyper / k_gen_prepare: one thread per chain; inside
// k_generations: the releasing CTA, one group of fields per thread) and kept in global memory; the sweep kernels only copy it,
// together with the exp/log tables, into shared memory.  Both forms evaluate the same expressions: bit-identical constants.
__device__ __forceinline__ void derive_scalars(FastH &H, const zz_hyper *hy, int group) {
  const double s2p = sqrt(2 * PI_D);
  if (group == 0) {
    H.R = hy->num_libraries; H.K = hy->num_active_components;
    H.s0 = hy->s0; H.s1 = hy->s1; H.tau = hy->tau; H.rtau = sqrt(hy->tau); H.irt = 1.0 / H.rtau;
    H.c1 = log(H.rtau * s2p);
  } else if (group == 1) {
    H.sp = hy->spike_probability; H.log_sp = log(H.sp); H.log_1msp = log(1 - H.sp);
    H.wa = hy->weight_active; H.one_m_wa = 1 - hy->weight_active;
    H.beta = (hy->beta < 0.0000001) ? 0.0 : hy->beta;   // P:121-125: beta < 1e-7 => likelihood 0
    H.temp = hy->temperature; H.vub = hy->variance_g_upper_bound;
    H.one_hi = 0x3FF00000;
  } else {
    const double sd_i = sqrt(hy->inactive_variances);
    H.im = hy->inactive_means; H.sd_i = sd_i;
    H.log_norm_i = log(s2p * sd_i);
    H.inv2iv = 1.0 / (2 * hy->inactive_variances);
    H.c0 = (1 - hy->spike_probability) / (s2p * sd_i);
  }
}
__device__ __forceinline__ void derive_comp(FastH &H, const zz_hyper *hy, int k) {
  const double s2p = sqrt(2 * PI_D);
  H.am[k] = hy->active_means[k];
  H.inv2av[k] = 1.0 / (2 * hy->active_variances[k]);
  H.ck[k] = hy->weight_within_active[k] / (s2p * sqrt(hy->active_variances[k]));
}
__device__ __forceinline__ void derive_lib(FastH &H, const zz_hyper *hy, int r) {
  H.alpha[r] = hy->alpha_r[r]; H.alpha2[r] = hy->alpha_r[r] * ZZ_KLOG2E; H.am256[r] = -256.0 * H.alpha2[r];
}
__device__ __forceinline__ void derive_fin(FastH &H, const zz_hyper *hy) {      // after every derive_lib
  const int R = hy->num_libraries;
  double amin = INFINITY, a2max = 0, amax = 0;
  for (int r = 0; r < R; ++r) {
    const double a = hy->alpha_r[r], a2 = a * ZZ_KLOG2E;
    if (!(a >= amin)) amin = a;
    if (a2 > a2max) a2max = a2;
    amax = fmax(amax, a);
  }
  H.alpha_min = amin;
  H.t_clamp = 1010.0 / a2max;
  H.uclamp_ok = (amin > 0 && amax <= 16 * amin) ? 1 : 0;   // same test as the host's kernel choice (launch_sweep_fast)
  H.px_on = 0; H.px_lib = 0; H.px_prev_acc = 0; H.pad_ = 0; H.px_am256 = 0; H.px_amin = 0; H.t_clamp_px = H.t_clamp;
}
__device__ __forceinline__ void derive_fast(FastH &H, const zz_hyper *hy) {
  for (int g = 0; g < 3; ++g) derive_scalars(H, hy, g);
  for (int r = 0; r < hy->num_libraries; ++r) derive_lib(H, hy, r);     // entries beyond R / K are never read
  for (int k = 0; k < hy->num_active_components; ++k) derive_comp(H, hy, k);
  derive_fin(H, hy);
}
// the same by a whole CTA: call, __syncthreads(), then one thread calls derive_fin
// (one libm logarithm per thread: on a single lane it is a chain of ~120 dependent instructions, 0.7 us; the expressions are those
//  of derive_scalars, evaluated on the same operands, so the constants are bit-identical)
__device__ __forceinline__ void derive_fast_par(FastH &H, const zz_hyper *hy, int tid) {
  const double s2p = sqrt(2 * PI_D);
  const int w = tid >> 5, lane = tid & 31;
  if (lane == 0) {                          // one task per WARP: different tasks on the lanes of one warp would run one after the other
    switch (w) {
      case 0: H.R = hy->num_libraries; H.K = hy->num_active_components; H.s0 = hy->s0; H.s1 = hy->s1; H.tau = hy->tau;
              H.rtau = sqrt(hy->tau); H.irt = 1.0 / H.rtau; break;
      case 1: H.c1 = log(sqrt(hy->tau) * s2p); break;
      case 2: H.sp = hy->spike_probability; H.log_sp = log(hy->spike_probability); H.wa = hy->weight_active; H.one_m_wa = 1 - hy->weight_active;
              H.beta = (hy->beta < 0.0000001) ? 0.0 : hy->beta; H.temp = hy->temperature; H.vub = hy->variance_g_upper_bound; H.one_hi = 0x3FF00000; break;
      case 3: H.log_1msp = log(1 - hy->spike_probability); break;
      case 4: H.im = hy->inactive_means; H.sd_i = sqrt(hy->inactive_variances); H.inv2iv = 1.0 / (2 * hy->inactive_variances); break;
      case 5: H.log_norm_i = log(s2p * sqrt(hy->inactive_variances)); break;
      case 6: H.c0 = (1 - hy->spike_probability) / (s2p * sqrt(hy->inactive_variances)); break;
      default: break;
    }
  } else if (w < 3) {                       // lanes 1.. of warps 0-2: the per-library constants (three multiplications each)
    const int r = w * 31 + lane - 1;
    if (r < hy->num_libraries) derive_lib(H, hy, r);
  } else if (w == 7 && lane <= hy->num_active_components) derive_comp(H, hy, lane - 1);
}
constexpr int DERIVE_FIN_THREAD = 7 * 32;   // derive_fin (+ set_px) runs here, beside derive_fast_par (it reads the hyper-parameters only)

__device__ __forceinline__ void load_tabs(MathTabs &T, const double *__restrict__ exp_tab, const double2 *__restrict__ log_tab) {
  const int t = threadIdx.x, nt = blockDim.x;
  for (int j = t; j < 256; j += nt) T.exp_t[j] = exp_tab[j];
  for (int j = t; j < 512; j += nt) T.log_t[j] = log_tab[j];
}
__device__ __forceinline__ void load_hyper_fast(FastH &H, const FastH *__restrict__ src) {   // ends with a block barrier
  const int t = threadIdx.x, nt = blockDim.x;
  static_assert(sizeof(FastH) % 8 == 0, "FastH is copied as 64-bit words");
  const uint64_t *s = reinterpret_cast<const uint64_t *>(src);
  uint64_t *d = reinterpret_cast<uint64_t *>(&H);
  for (int j = t; j < (int)(sizeof(FastH) / 8); j += nt) d[j] = s[j];
  __syncthreads();
}
__device__ __forceinline__ void load_fast(FastH &H, MathTabs &T, const FastH *__restrict__ src,
                                          const double *__restrict__ exp_tab, const double2 *__restrict__ log_tab) {
  load_tabs(T, exp_tab, log_tab);
  load_hyper_fast(H, src);
}

// z = sqrt(-2 ln ua) * cos(2 pi ub)  (same draw as box_muller's z0; the sine branch is only needed by the spike move)
__device__ __forceinline__ double normal_cos(double ua, double ub, const MathTabs &T) {
  return sqrt(-2.0 * flog_core(ua, T.log_t)) * cospi(2.0 * ub);   // ua in (0,1): always a normal number
}

// Per-gene sufficient statistics of the DATA, computed once at zz_upload_data (k_prepare_data): the normal part of
// computeXgLikelihood depends on Xg only through  sum_detected (x - y)^2 = ss + n_det (xbar - y)^2  (all terms >= 0, so no
// cancellation), and the detection part only through which libraries are -Inf.  The sweep therefore never re-reads Xg.
struct GeneData {
  uint64_t nd_mask;   // bit r set <=> Xg[g, r] == -Inf
  double xbar, ss;    // mean of the detected x, sum of squared deviations from it
  int ndet;           // number of detected libraries
};

__device__ __forceinline__ double sqdist(const GeneData &d, double y) {   // D(y) = sum_detected (x - y)^2
  const double t = d.xbar - y;
  return fma((double)d.ndet * t, t, d.ss);
}

// A(y) through det_q: the factors q_r of all libraries are multiplied (four independent partial products, so the per-library
// chains overlap) and ONE logarithm is taken; one clamp of t instead of one per library.  Requires max alpha <= 16 min alpha
// (checked by the host before this kernel is chosen): then t_clamp * alpha2_r >= 1010/16 > 54 for every r, so a clamped t
// yields p_x == 1 exactly where the unclamped one does.
template <int RT>
__device__ __forceinline__ double detect_logsum_sl(const FastH &H, const MathTabs &T, uint64_t nd_mask, double t) {
  static_assert(RT > 0 && RT <= 18, "the product of the factors must stay a normal number");
  const double tc = fmin(t, H.t_clamp);                                 // +Inf clamps too
  const uint32_t mlo = (uint32_t)nd_mask;
  double P[4] = {1.0, 1.0, 1.0, 1.0};
#pragma unroll
  for (int r = 0; r < RT; ++r) {
    const double q = det_q(tc, H.am256[r], mlo & (1u << r), T);
    P[r & 3] = (r < 4) ? q : P[r & 3] * q;
  }
  return det_log_product((P[0] * P[1]) * (P[2] * P[3]), T.log_t);
}

// A(y) = sum_r log(nd_r ? 1 - p_r : p_r),  p_r = 1 - exp(-t alpha_r),  t = gl * exp(y)     (U:157, P:82-84)
template <int RT>
__device__ __forceinline__ double detect_logsum(const FastH &H, const MathTabs &T, uint64_t nd_mask, double t) {
  const int R = RT > 0 ? RT : H.R;
  double A = 0;
  // every path that produces A(y) for the same hyper-parameters must use the same arithmetic (the value is cached in `lps`
  // and compared bit for bit across kernels): the product form whenever it is applicable
  if constexpr (RT > 0 && RT <= 18) { if (H.uclamp_ok) return detect_logsum_sl<RT>(H, T, nd_mask, t); }
  if (RT > 0 && RT <= 18) {
    // q is +0 or a normal number in [2^-53, 1].  flog_core(+0) returns -1023 ln 2 = -709.1 instead of -Inf, while a sum of
    // at most 18 legitimate terms is >= 18 log(2^-53) = -661.3: "A < -700" therefore detects a zero q exactly, without a
    // per-library test.  Libraries are processed ZZ_ILP at a time with separate temporaries so the scheduler can overlap
    // their (long, strictly dependent) exp -> log chains; the additions into A stay in library order.
#pragma unroll
    for (int r0 = 0; r0 < R; r0 += ZZ_ILP) {
      double l[ZZ_ILP];
#pragma unroll
      for (int j = 0; j < ZZ_ILP; ++j) {
        if (r0 + j < R) {
          const int r = r0 + j;
          const double e = fexp2_neg_clamped(t * H.alpha2[r], T.exp_t);
          const double p = 1 - e;
          const double q = ((nd_mask >> r) & 1ull) ? (1 - p) : p;
          l[j] = flog_core(q, T.log_t);
        }
      }
#pragma unroll
      for (int j = 0; j < ZZ_ILP; ++j)
        if (r0 + j < R) A += l[j];
    }
    return A < -700.0 ? -INFINITY : A;
  }
  unsigned zero_seen = 0xFFFFFFFFu;
  for (int r = 0; r < R; ++r) {
    const double e = fexp2_neg_clamped(t * H.alpha2[r], T.exp_t);
    const double p = 1 - e;
    const double q = ((nd_mask >> r) & 1ull) ? (1 - p) : p;
    A += flog_core(q, T.log_t);
    zero_seen = min(zero_seen, (unsigned)(d_hi(q) | d_lo(q)));
  }
  return zero_seen == 0 ? -INFINITY : A;              // some q == 0: log(0) = -Inf (flog_core(0) itself is finite garbage)
}

__device__ __forceinline__ double level2_quad(const FastH &H, double y, uint32_t f) {   // (y - mean)^2 / (2 var) of the gene's component
  if (f & F_ACTIVE) { const int k = flag_k(f) - 1; const double d = y - H.am[k]; return d * d * H.inv2av[k]; }
  const double d = y - H.im;
  return d * d * H.inv2iv;
}

struct FastGene {
  double y, v, lv, rv;     // Yg, variance_g, log(variance_g), 1/variance_g
  double A;                // sum of the detection log-terms at y
  uint32_t f;
};

// computeVarianceGPriorProbability's standardised residual  q = (log v - log Sg - tau) / sqrt(tau),  log Sg = s0 + s1 y
__device__ __forceinline__ double vg_q(const FastH &H, double lv, double y) { return (lv - fma(H.s1, y, H.s0) - H.tau) * H.irt; }
__device__ __forceinline__ double vg_prior_from_q(const FastH &H, double lv, double q) { return -(lv + H.c1) - 0.5 * (q * q); }

__device__ __noinline__ double plnorm_term(const FastH &H, double y) {   // only when variance_g_upper_bound < Inf (P:24-25); rare: kept out of line
  return plnorm_log(H.vub, fma(H.s1, y, H.s0) + H.tau, H.rtau);
}

// beta * x with the reference's convention beta < 1e-7 => likelihood exactly 0 even where x is -Inf (P:121-125)
__device__ __forceinline__ double times_beta(const FastH &H, double x) { return H.beta == 0.0 ? 0.0 : H.beta * x; }

// XgLikelihood = beta (A + B),  B = -n_det (ln sqrt(2 pi) + 0.5 log v) - 0.5 D(y) / v      (P:57-127 re-expressed)
__device__ __forceinline__ double xg_lik_fast(const FastH &H, const FastGene &g, const GeneData &d) {
  return times_beta(H, g.A + (-(double)d.ndet * (LN_SQRT_2PI + 0.5 * g.lv) - 0.5 * sqdist(d, g.y) * g.rv));
}

// mhYg (Pr:250-333)
// `yp`, `tp` = proposed Yg (Pr:257) and gl * exp(yp) (U:157).  `saturated` is warp-uniform: every proposing lane of the warp has
// tp * alpha_r >= 37.5 > 54 ln 2 for all r, so exp(-tp alpha_r) <= 2^-54 and p_x = 1 - exp(.) rounds to exactly 1: the
// detection term is log(1) = 0 per detected library and log(0) = -Inf if any library is undetected — no exp/log needed.
// (Genes are stored sorted by expression, so whole warps of highly expressed genes take this path.)
template <int RT>
__device__ __forceinline__ int fast_yg(const FastH &H, const MathTabs &T, FastGene &g, const GeneData &d, double yp, double tp,
                                       bool saturated, double u, double &ratio) {
  const double Ap = saturated ? (d.nd_mask ? -INFINITY : 0.0) : detect_logsum<RT>(H, T, d.nd_mask, tp);   // Pr:261-265
  const double dLX = times_beta(H, (Ap - g.A) - 0.5 * (sqdist(d, yp) - sqdist(d, g.y)) * g.rv);   // XgLik' - XgLik
  const double q = vg_q(H, g.lv, g.y), qp = vg_q(H, g.lv, yp);
  double dPV = -0.5 * (qp * qp - q * q);                                // Pr:267 - cache
  if (H.vub != INFINITY) dPV -= plnorm_term(H, yp) - plnorm_term(H, g.y);
  const double dL2 = -(level2_quad(H, yp, g.f) - level2_quad(H, g.y, g.f));  // Pr:271-291
  ratio = H.temp * (dLX + dL2 + dPV);                                   // Pr:296
  const int acc = (flog_core(u, T.log_t) < ratio);                      // Pr:302 (u in (0,1): normal)
  if (acc) { g.y = yp; g.A = Ap; }
  return acc;
}

// mhVariance_g (Pr:7-56): no per-library work, p_x does not depend on variance_g
__device__ __forceinline__ int fast_vg(const FastH &H, const MathTabs &T, FastGene &g, const GeneData &d, double tv, double u1,
                                       double u2, double &ratio) {
  const double arg = tv * (u1 - 0.5);                                   // MP:5; HR = log(sf) = arg (Pr:15)
  const double sf = fexp(arg, T.exp_t);
  const double vp = g.v * sf;                                           // MP:7
  const double rvp = g.rv * fexp(-arg, T.exp_t);                        // 1 / v'
  const double lvp = g.lv + arg;
  const double dB = -(double)d.ndet * (0.5 * arg) - 0.5 * sqdist(d, g.y) * (rvp - g.rv);
  const double q = vg_q(H, g.lv, g.y), qp = q + arg * H.irt;
  const double dPV = -arg - 0.5 * (qp * qp - q * q);                    // the plnorm term (P:25) does not depend on v
  ratio = H.temp * (times_beta(H, dB) + dPV) + arg;                     // Pr:24
  // a current XgLikelihood of -Inf (or NaN) makes Lp - Lc NaN in the reference (both sides carry the same -Inf detection
  // term): keep that behaviour (reject + NaN flag) although the term cancels analytically here
  if (!(g.A > -INFINITY) && H.beta != 0.0) ratio = NAN;
  if (vp > H.vub) ratio = -INFINITY;                                    // Pr:26
  const int acc = (flog_core(u2, T.log_t) < ratio);                     // Pr:35
  if (acc) { g.v = vp; g.lv = lvp; g.rv = rvp; }
  return acc;
}

// gibbsAllocationActiveInactive + gibbsAllocationWithinActive (Pr:335-410); returns the allocation word (0 or k)
__device__ __forceinline__ int fast_alloc(const FastH &H, const MathTabs &T, double y, uint32_t &f, double u_ai, double u_w) {
  const bool in_spike = (f & F_INSPIKE) != 0;
  const double d0 = y - H.im;
  const double L0 = in_spike ? H.sp : H.c0 * fexp(-(d0 * d0) * H.inv2iv, T.exp_t);   // Pr:339
  double Lk[ZZ_MAX_COMP];
  double L1 = 0;
#pragma unroll
  for (int k = 0; k < ZZ_MAX_COMP; ++k) {
    if (k < H.K) {
      const double d = y - H.am[k];
      Lk[k] = H.ck[k] * fexp(-(d * d) * H.inv2av[k], T.exp_t);          // Pr:344 / Pr:375-395
      L1 += Lk[k];
    }
  }
  const double num = in_spike ? 0.0 : H.wa * L1;
  double pa = num / (H.one_m_wa * L0 + num);                            // Pr:348
  if (isnan(pa)) pa = H.wa;                                             // Pr:349
  int a = (u_ai < pa) ? 1 : 0;                                          // Pr:353
  if (f & F_PINNED) a = 1;                                              // Pr:355
  f = (f & ~F_ACTIVE) | (a ? F_ACTIVE : 0u);
  if (!a) return 0;
  if (L1 > 0 && L1 < INFINITY) {                                        // NaN rows keep the old k, Pr:403
    const double thr = u_w * L1;
    double cum = 0; int cnt = 0;
#pragma unroll
    for (int k = 0; k < ZZ_MAX_COMP; ++k)
      if (k < H.K) { cum += Lk[k]; cnt += (cum < thr); }                // Pr:398-402
    const int kn = cnt + 1 > H.K ? H.K : cnt + 1;
    f = flag_set_k(f, kn);
  }
  return flag_k(f);
}

// ---- sufficient statistics accumulated by the sweep itself (layout: zz_suffstats in include/zigzag_gpu.h) --------------
constexpr int NSTAT_FIXED_ = 10;
constexpr int MAX_NSTAT_ = 3 * (ZZ_MAX_COMP + 1) + NSTAT_FIXED_;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sufficient statistics accumulated by the sweep: every thread owns one column of a [nstat][BLOCK] shared-memory array and
// adds its gene's contribution after each tile (3 instructions per statistic, no shuffles); the block reduces the columns
// once at the end.  Tiles are assigned to warps statically, so every sum has a fixed order and the result is reproducible.
template <int BLOCK_>
__device__ __forceinline__ void stats_add(double *acc /*[nstat][BLOCK_]*/, int K, bool valid, double y, double lv, uint32_t f,
                                          double xglik, double vgprob, bool finite_state) {
  if (!valid) return;
  const int K1 = K + 1;
  double *col = acc + threadIdx.x;
  const bool act = f & F_ACTIVE, sp = f & F_INSPIKE;
  const int comp = act ? flag_k(f) : 0;                                 // Pr:415 all_allocations
  col[comp * BLOCK_] += 1.0;
  if (act || !sp) { col[(K1 + comp) * BLOCK_] += y; col[(2 * K1 + comp) * BLOCK_] += y * y; }
  double *fx = col + 3 * K1 * BLOCK_;
  if (sp) fx[0] += 1.0;
  else {
    fx[1 * BLOCK_] += 1.0; fx[2 * BLOCK_] += lv; fx[3 * BLOCK_] += lv * lv; fx[4 * BLOCK_] += y; fx[5 * BLOCK_] += y * y;
    fx[6 * BLOCK_] += y * lv; fx[8 * BLOCK_] += vgprob;
  }
  fx[7 * BLOCK_] += xglik;
  if (!finite_state) fx[9 * BLOCK_] += 1.0;
}

// column sums -> one row of `partials`: after one barrier, warp w reduces statistics w, w + nwarps, ...; each lane adds its
// BLOCK_/32 columns in a fixed order, then a butterfly.  `red` is unused (kept for the signature).
template <int BLOCK_>
__device__ __forceinline__ void stats_flush(const double *acc, int nstat, double *row, double *red) {
  (void)red;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int s = w; s < nstat; s += BLOCK_ / 32) {
    double v = 0;
#pragma unroll
    for (int j = 0; j < BLOCK_ / 32; ++j) v += acc[s * BLOCK_ + j * 32 + lane];
    v = warp_sum(v);
    if (lane == 0) row[s] = v;
  }
}

}  // namespace zz
