// zz_device.cuh — device-side arithmetic of zigzag's per-gene moves (sm_100a).
//
// Each __device__ function is the per-gene reading of one reference routine, kept in the reference's operation
// order (compile with -fmad=false so a*b+c is not contracted).  Citations: P = R/zigzagProbability.R,
// Pr = R/zigzagProposals.R, MP = R/zigzagMoveProposals.R, U = R/zigzagUtilities.R, I = R/zigzagInitialize.R.
// Derived fields the reference materialises for all genes (Sg[G], p_x[GxR], set_sigmaX_pX U:133-142) are pure
// functions of (Yg, hyper) and are recomputed in registers instead of being stored in HBM.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/zigzag_gpu.h"

namespace zz {

// flag word, one uint16 per gene
constexpr uint32_t F_ACTIVE = 1u;    // allocation_active_inactive
constexpr uint32_t F_INSPIKE = 2u;   // inactive_spike_allocation
constexpr uint32_t F_ALLZERO = 4u;   // no_detect_spike / all_zero_idx membership
constexpr uint32_t F_PINNED = 8u;    // active_gene_set membership
constexpr uint32_t F_K_SHIFT = 8u;   // allocation_within_active - 1 in bits 8..11
constexpr uint32_t F_K_MASK = 0xFu << F_K_SHIFT;

__device__ __forceinline__ int flag_k(uint32_t f) { return (int)((f & F_K_MASK) >> F_K_SHIFT) + 1; }
__device__ __forceinline__ uint32_t flag_set_k(uint32_t f, int k) { return (f & ~F_K_MASK) | ((uint32_t)(k - 1) << F_K_SHIFT); }

constexpr double LN_SQRT_2PI = 0.918938533204672741780329736406;  // nmath/dnorm.c M_LN_SQRT_2PI
constexpr double PI_D = 3.141592653589793238462643383280;

// Block-shared copy of one chain's hyper-parameters plus the gene-independent sub-expressions of the reference
// formulas (each is evaluated exactly as the reference writes it, just once per block instead of once per gene).
struct HyperS {
  int R, K;
  double s0, s1, tau, rtau;            // rtau = sqrt(tau)                    P:16
  double spike_p, im, iv, wa, beta, temp, vub;
  double sqrt2pi;                      // sqrt(2*pi)                          I:126
  double log_sp, log_1msp;             // log(spike_p), log(1 - spike_p)      P:218-220
  double log_norm_i, two_iv;           // log(sqrt2pi * sqrt(iv)), 2*iv       P:220
  double log_rtau;                     // unused by the faithful path (kept for the fast path)
  double alpha[ZZ_MAX_LIBS];
  double am[ZZ_MAX_COMP], two_av[ZZ_MAX_COMP], log_norm_a[ZZ_MAX_COMP];  // P:252
  double w[ZZ_MAX_COMP], logw[ZZ_MAX_COMP];                                // Pr:344, Pr:375
  double sd_i;                         // sqrt(iv)                            Pr:545
};

__device__ __forceinline__ void load_hyper(HyperS &H, const zz_hyper *__restrict__ hy) {
  // called by all threads of the block; thread-strided so the copy is coalesced
  const int t = threadIdx.x, nt = blockDim.x;
  if (t == 0) {
    H.R = hy->num_libraries; H.K = hy->num_active_components;
    H.s0 = hy->s0; H.s1 = hy->s1; H.tau = hy->tau; H.rtau = sqrt(hy->tau);
    H.spike_p = hy->spike_probability; H.im = hy->inactive_means; H.iv = hy->inactive_variances;
    H.wa = hy->weight_active; H.beta = hy->beta; H.temp = hy->temperature; H.vub = hy->variance_g_upper_bound;
    H.sqrt2pi = sqrt(2 * PI_D);
    H.log_sp = log(hy->spike_probability);
    H.log_1msp = log(1 - hy->spike_probability);
    H.sd_i = sqrt(hy->inactive_variances);
    H.log_norm_i = log(H.sqrt2pi * H.sd_i);
    H.two_iv = 2 * hy->inactive_variances;
    H.log_rtau = log(H.rtau);
  }
  for (int r = t; r < ZZ_MAX_LIBS; r += nt) H.alpha[r] = hy->alpha_r[r];
  if (t < ZZ_MAX_COMP) {
    const double s2p = sqrt(2 * PI_D);
    H.am[t] = hy->active_means[t];
    H.two_av[t] = 2 * hy->active_variances[t];
    H.log_norm_a[t] = log(s2p * sqrt(hy->active_variances[t]));
    H.w[t] = hy->weight_within_active[t];
    H.logw[t] = log(hy->weight_within_active[t]);
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------ densities

// U:144-148 get_sigmax
__device__ __forceinline__ double get_sigmax(double s0, double s1, double y) { return exp(s0 + s1 * y); }

// stats::dnorm(x, mu, sd, log=TRUE) (nmath/dnorm.c).  For finite x and sd > 0 the generic expression below propagates
// Inf/NaN exactly like nmath's explicit branches; sd == 0 is the one case that needs them.
__device__ __forceinline__ double dnorm_log(double x, double mu, double sd, double log_sd) {
  if (sd == 0) return (isnan(x) || isnan(mu)) ? x + mu : ((x == mu) ? INFINITY : -INFINITY);
  double z = (x - mu) / sd;
  return -(LN_SQRT_2PI + 0.5 * z * z + log_sd);
}

// P:57-127 computeXgLikelihood for an out-of-spike gene, with get_px (U:150-159) fused in:
//   p_gr = 1 - exp(-(gl*exp(y)) * alpha_r);  ll = x==-Inf ? log(1 - p) : log(p) + dnorm(x; y, sqrt(v))
// `xg` points at this gene's entry of library 0; libraries are `stride` doubles apart (library-major SoA).
__device__ __forceinline__ double xg_lik_out(const HyperS &H, const double *__restrict__ xg, int64_t stride,
                                             double y, double v, double gl) {
  const double t = gl * exp(y);
  const double sd = sqrt(v);
  const double lsd = log(sd);
  double acc = 0;
  for (int r = 0; r < H.R; ++r) {
    const double x = __ldg(xg + (int64_t)r * stride);
    const double p = 1 - exp(-(t * H.alpha[r]));
    const bool nd = (x == -INFINITY);
    double ll = log(nd ? (1 - p) : p);           // P:82 / P:84
    if (!nd) ll = ll + dnorm_log(x, y, sd, lsd); // P:85
    acc += ll;                                   // rowSums, P:73
  }
  if (H.beta < 0.0000001) acc = 0;               // P:121-123
  return H.beta * acc;                           // P:125
}

// in-spike rows: log(spike_allocation * nd_spike), P:117
__device__ __forceinline__ double xg_lik_inspike(const HyperS &H, bool all_zero) {
  double lnl = all_zero ? 0.0 : -INFINITY;
  if (H.beta < 0.0000001) lnl = 0;
  return H.beta * lnl;
}

// single-library column of computeXgLikelihood as mhP_x uses it, Pr:213-215
__device__ __forceinline__ double xg_lik_onelib(const HyperS &H, double x, double y, double v, double gl, double alpha,
                                                bool in_spike, bool all_zero) {
  if (in_spike) return xg_lik_inspike(H, all_zero);
  const double p = 1 - exp(-((gl * exp(y)) * alpha));
  const bool nd = (x == -INFINITY);
  double ll = log(nd ? (1 - p) : p);
  if (!nd) { const double sd = sqrt(v); ll = ll + dnorm_log(x, y, sd, log(sd)); }
  if (H.beta < 0.0000001) ll = 0;
  return H.beta * ll;
}

// log of the lognormal CDF (stats::plnorm(..., log.p=TRUE)); only used when variance_g_upper_bound < Inf
__device__ __forceinline__ double plnorm_log(double q, double meanlog, double sdlog) {
  if (q <= 0) return -INFINITY;
  double zz_ = (log(q) - meanlog) / sdlog;
  return log(0.5 * erfc(-zz_ / sqrt(2.0)));
}

// P:11-31 computeVarianceGPriorProbability(x = variance_g, sx = Sg, gtau)
__device__ __forceinline__ double varg_prior(const HyperS &H, double x, double sx, double gtau, double rtgtau) {
  const double q = (log(x) - log(sx) - gtau) / rtgtau;
  double lnp = -log(x * rtgtau * H.sqrt2pi) - 0.5 * (q * q);
  if (H.vub != INFINITY) lnp = lnp - plnorm_log(H.vub, log(sx) + gtau, sqrt(gtau));
  return lnp;
}

// P:210-238 computeInactiveLikelihood
__device__ __forceinline__ double inactive_lik(const HyperS &H, double y, bool in_spike) {
  if (in_spike) return H.log_sp;
  const double d = y - H.im;
  return H.log_1msp - H.log_norm_i - (d * d) / H.two_iv;
}
// same with explicit (im, iv): the hyper moves evaluate it at proposed values
__device__ __forceinline__ double inactive_lik_at(const HyperS &H, double y, bool in_spike, double im, double iv) {
  if (in_spike) return H.log_sp;
  const double d = y - im;
  return H.log_1msp - log(H.sqrt2pi * sqrt(iv)) - (d * d) / (2 * iv);
}

// P:240-273 computeActiveLikelihood, k in 1..K
__device__ __forceinline__ double active_lik(const HyperS &H, double y, int k) {
  const double d = y - H.am[k - 1];
  return -H.log_norm_a[k - 1] - (d * d) / H.two_av[k - 1];
}

__device__ __forceinline__ double level2_lik(const HyperS &H, double y, uint32_t f, bool in_spike) {
  return (f & F_ACTIVE) ? active_lik(H, y, flag_k(f)) : inactive_lik(H, y, in_spike);
}

// ------------------------------------------------------------------------------------------------ RNG

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                               uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ double u32_open(uint32_t w) {  // (w + 0.5) * 2^-32: open interval, 32-bit resolution (like R's unif_rand)
  return ((double)w + 0.5) * (1.0 / 4294967296.0);
}

// One Philox block: counter = (global gene index, blk | chain<<8, step lo, step hi); key = seed.
// Draw map of a sweep (specified in DESIGN.md "RNG"; the test oracle implements the same map):
//   blk 0: w0,w1 -> Box-Muller -> z of mhYg (cos branch); w2 -> u of mhYg; w3 -> u_proposal of mhVariance_g
//   blk 1: w0 -> u_accept of mhVariance_g; w1 -> u of the active/inactive draw; w2 -> u of the within-active draw
//   blk 2: w0,w1 -> Box-Muller -> (z_y, z_v) of mhSpikeAllocation; w2 -> its u
__device__ __forceinline__ void philox_block(uint64_t seed, uint32_t chain, uint64_t step, uint32_t blk, uint64_t gene, uint32_t o[4]) {
  philox4x32_10((uint32_t)gene, (blk & 255u) | (chain << 8), (uint32_t)step, (uint32_t)(step >> 32), (uint32_t)seed,
                (uint32_t)(seed >> 32), o);
}

__device__ __forceinline__ void box_muller(double ua, double ub, double &z0, double &z1) {
  const double r = sqrt(-2.0 * log(ua));
  double s, c;
  sincospi(2.0 * ub, &s, &c);
  z0 = r * c; z1 = r * s;
}

// ------------------------------------------------------------------------------------------------ accept windows

// 100-wide sliding accept window (Yg_trace / variance_g_trace rows) packed in 2 x u64; cbind(trace[,-1], bit) Pr:316-320
__device__ __forceinline__ void window_push(ulonglong2 &w, int bit) {
  w.y = ((w.y << 1) | (w.x >> 63)) & ((1ULL << 36) - 1);
  w.x = (w.x << 1) | (unsigned long long)(bit & 1);
}
__device__ __forceinline__ int window_sum(const ulonglong2 &w) { return __popcll(w.x) + __popcll(w.y & ((1ULL << 36) - 1)); }

// ------------------------------------------------------------------------------------------------ per-gene moves

struct Gene {           // one gene's mutable state, held in registers across the moves of a sweep
  double y, v;          // Yg, variance_g
  double xglik, vgprob; // caches XgLikelihood, variance_g_probability
  uint32_t f;           // flag word
};

// mhYg, Pr:250-333.  Returns the accept bit; `ratio` receives R (NaN when the gene is not proposed).
__device__ __forceinline__ int move_yg(const HyperS &H, Gene &g, const double *__restrict__ xg, int64_t stride, double gl,
                                       double ty, double z, double u, double &ratio) {
  ratio = NAN;
  if (g.f & F_INSPIKE) return 0;                                   // out_spike_idx only, Pr:252-257
  const double yp = g.y + ty * z;                                  // Pr:257
  const double Sgp = get_sigmax(H.s0, H.s1, yp);                   // Pr:259
  const double LXp = xg_lik_out(H, xg, stride, yp, g.v, gl);       // Pr:261-265
  const double PVp = varg_prior(H, g.v, Sgp, H.tau, H.rtau);       // Pr:267
  double Lp = LXp;
  Lp = Lp + level2_lik(H, yp, g.f, false);                         // Pr:271-277
  Lp = Lp + PVp;                                                   // Pr:279
  double Lc = g.xglik;
  Lc = Lc + level2_lik(H, g.y, g.f, false);                        // Pr:285-291
  Lc = Lc + g.vgprob;                                              // Pr:293
  ratio = H.temp * (Lp - Lc);                                      // Pr:296
  const int acc = (log(u) < ratio);                                // Pr:302 (NaN => reject)
  if (acc) { g.y = yp; g.xglik = LXp; g.vgprob = PVp; }            // Pr:310,325,331
  return acc;
}

// mhVariance_g, Pr:7-56 with scale_move MP:3-13
__device__ __forceinline__ int move_variance_g(const HyperS &H, Gene &g, const double *__restrict__ xg, int64_t stride, double gl,
                                               double tv, double u1, double u2, double &ratio) {
  ratio = NAN;
  if (g.f & F_INSPIKE) return 0;
  const double sf = exp(tv * (u1 - 0.5));                          // MP:5
  const double vp = g.v * sf;                                      // MP:7
  const double HR = log(sf);                                       // Pr:15
  const double LXp = xg_lik_out(H, xg, stride, g.y, vp, gl);       // Pr:17 (field p_x == get_px(current Yg))
  const double Sg = get_sigmax(H.s0, H.s1, g.y);                   // field Sg, U:135
  const double PVp = varg_prior(H, vp, Sg, H.tau, H.rtau);         // Pr:19
  const double Lp = LXp + PVp;                                     // Pr:21
  const double Lc = g.xglik + g.vgprob;                            // Pr:22
  ratio = H.temp * (Lp - Lc) + HR;                                 // Pr:24
  if (vp > H.vub) ratio = -INFINITY;                               // Pr:26
  const int acc = (log(u2) < ratio);                               // Pr:35
  if (acc) { g.v = vp; g.xglik = LXp; g.vgprob = PVp; }            // Pr:38,52,54
  return acc;
}

// gibbsAllocationWithinActive for one active gene, Pr:363-410.  Returns k in 1..K, or 0 for a NaN row (keep old k).
__device__ __forceinline__ int within_active(const HyperS &H, double y, double u) {
  double Lk[ZZ_MAX_COMP];
  double rs = 0;
#pragma unroll
  for (int k = 0; k < ZZ_MAX_COMP; ++k) {
    if (k < H.K) {
      Lk[k] = exp(H.logw[k] + active_lik(H, y, k + 1) - 0);        // Pr:375-376, Pr:395
      rs += Lk[k];                                                 // rowSums, Pr:398
    }
  }
  double cum = 0;
  int cnt = 0;
  bool nan_row = false;
#pragma unroll
  for (int k = 0; k < ZZ_MAX_COMP; ++k) {
    if (k < H.K) {
      cum += Lk[k] / rs;                                           // Pr:398-399
      const double q = floor(cum / u);                             // Pr:400
      nan_row |= isnan(q);
      cnt += (q == 0);                                             // Pr:401-402
    }
  }
  if (nan_row) return 0;                                           // Pr:403
  int kn = cnt + 1;
  return kn > H.K ? H.K : kn;                                      // clamp the K+1 rounding case (documented deviation)
}

// gibbsAllocationActiveInactive + chained gibbsAllocationWithinActive, Pr:335-361.  Returns the allocation word (0 or k).
__device__ __forceinline__ int move_alloc(const HyperS &H, Gene &g, double u_ai, double u_w, double &pa) {
  const bool in_spike = (g.f & F_INSPIKE) != 0;
  const double L0 = exp(inactive_lik(H, g.y, in_spike));           // Pr:339
  double L1 = 0;
#pragma unroll
  for (int k = 0; k < ZZ_MAX_COMP; ++k)
    if (k < H.K) L1 = L1 + H.w[k] * exp(active_lik(H, g.y, k + 1)); // Pr:342-346
  const double one_m_sa = in_spike ? 0.0 : 1.0;
  pa = H.wa * L1 * one_m_sa / ((1 - H.wa) * L0 + H.wa * L1 * one_m_sa); // Pr:348
  if (isnan(pa)) pa = H.wa;                                        // Pr:349
  int a = (u_ai < pa) ? 1 : 0;                                     // Pr:353
  if (g.f & F_PINNED) a = 1;                                       // Pr:355
  g.f = (g.f & ~F_ACTIVE) | (a ? F_ACTIVE : 0u);
  if (a) {
    const int kn = within_active(H, g.y, u_w);                     // Pr:359
    if (kn > 0) g.f = flag_set_k(g.f, kn);
    return flag_k(g.f);
  }
  return 0;
}

// mhSpikeAllocation, Pr:528-638
__device__ __forceinline__ int move_spike(const HyperS &H, Gene &g, const double *__restrict__ xg, int64_t stride, double gl,
                                          double zy, double zv, double u, double &ratio) {
  ratio = NAN;
  if (!((g.f & F_ALLZERO) && !(g.f & F_ACTIVE))) return 0;        // Pr:533
  const bool cur_in = (g.f & F_INSPIKE) != 0;
  const bool prop_in = !cur_in;                                    // Pr:535
  const double yp = H.im + H.sd_i * zy;                            // Pr:545
  const double Sgp = get_sigmax(H.s0, H.s1, yp);                   // Pr:547
  const double vp = exp((log(Sgp) + H.tau) + H.rtau * zv);         // Pr:549
  const double PVp = varg_prior(H, vp, Sgp, H.tau, H.rtau);        // Pr:556
  const double PVc = varg_prior(H, g.v, get_sigmax(H.s0, H.s1, g.y), H.tau, H.rtau); // Pr:558
  const double pygp = inactive_lik(H, yp, prop_in);                // Pr:560
  const double pyg = inactive_lik(H, g.y, cur_in);                 // Pr:563
  const double pp = prop_in ? H.log_sp : (pygp + PVp);             // Pr:569-570
  const double pc = cur_in ? H.log_sp : (pyg + PVc);               // Pr:572-573
  double Lp = prop_in ? xg_lik_inspike(H, true) : xg_lik_out(H, xg, stride, yp, vp, gl); // Pr:579
  double Lc = g.xglik;                                             // Pr:581
  if (Lp == -INFINITY && Lc == -INFINITY) { Lp = 0; Lc = 0; }      // Pr:583-585
  const double HR = prop_in ? (pyg - H.log_1msp + PVc)             // Pr:594-595
                            : -(pygp - H.log_1msp + PVp);          // Pr:592-593
  ratio = H.temp * (Lp - Lc + pp - pc) + HR;                       // Pr:601
  if (vp > H.vub) ratio = -INFINITY;                               // Pr:603
  const int acc = (log(u) < ratio);                                // Pr:605
  const bool new_in = acc ? prop_in : cur_in;                      // Pr:616
  g.f = (g.f & ~F_INSPIKE) | (new_in ? F_INSPIKE : 0u);
  if (!prop_in) {                                                  // [proposed_out_spike_idx]
    g.y = acc ? yp : g.y;                                          // Pr:620
    g.v = acc ? vp : g.v;                                          // Pr:624
    g.vgprob = acc ? PVp : PVc;                                    // Pr:636
  }
  g.xglik = acc ? Lp : Lc;                                         // Pr:630-632
  return acc;
}

// ------------------------------------------------------------------------------------------------ block reduction

// Deterministic block sum: warp shuffle tree, then the warps' partials are added in warp order by thread 0.
template <int BLOCK>
__device__ __forceinline__ double block_sum(double v, double *smem /*BLOCK/32*/) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) smem[w] = v;
  __syncthreads();
  double r = 0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < BLOCK / 32; ++i) r += smem[i];
  }
  return r;  // valid on thread 0
}

}  // namespace zz
