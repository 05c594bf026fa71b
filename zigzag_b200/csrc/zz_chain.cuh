// zz_chain.cuh — the scalar half of a generation, on the device.
//
// The reference's ten hyper-parameter moves (R/zigzagProposals.R) are scalar Metropolis / Gibbs moves whose only O(G) part is a
// sum over genes.  Here a single device thread (k_chain_stage) computes each move's proposal, prior ratio, Hastings ratio,
// accept/reject and tuning window between the gene-sum kernels; the proposal travels to those kernels through a mailbox in
// global memory (DevChain), accepted moves raise gates that the cache-commit kernels test.  The host only enqueues.
// Same moves, same order and the same private Philox stream as the fused schedule of the CPU test chain, so a device-resident
// segment can be replayed on the CPU decision for decision (tests/test_gpu_chain.py).
#pragma once
#include "zz_device.cuh"
#include "zz_fast.cuh"

namespace zz {

struct L2Params { double im, iv, am[ZZ_MAX_COMP], av[ZZ_MAX_COMP]; };

struct DevChain {
  zz_priors pr;
  zz_chain_scalars sc;
  // mailbox read by the gene-sum kernels
  L2Params l2[2];                 // RED_LEVEL2 with grid.y = 2: [0] proposed, [1] current parameters (Pr:444-447, 653-654, 780-784)
  double p_s0, p_s1, p_tau;       // RED_VARG_HYPER proposal (Pr:84-87, 126-130, 167-170)
  double p_alpha, p_alpha_old; int p_lib, pad0_;   // RED_PXDELTA proposal (Pr:209-224); the replaced alpha_r for the XgLikelihood commit
  int gate_vg, gate_px, pad_;     // accepted mhTau/mhSg/mhS0Tau -> commit variance_g_probability; accepted mhP_x -> commit XgLikelihood
  // the pending move, between its "propose" and its "decide" stage
  double HR, pp, pc, difp, prop0, prop1;
  int sel, k;
  int chain, pad2_;               // index of this chain in the handle (chain field of its Philox counters)
};

// ---- scalar draws: private Philox stream, four uniforms per block (Marsaglia-Tsang gamma, beta from two gammas) ----
__device__ inline double c_unif(DevChain &c, uint64_t seed) {
  if ((c.sc.hyper_ctr & 3) == 0) {
    uint32_t o[4];
    philox_block(seed, (uint32_t)c.chain, c.sc.hyper_ctr >> 2, 255, 0xFFFFFFFFull, o);
    for (int j = 0; j < 4; ++j) c.sc.hyper_buf[j] = u32_open(o[j]);
  }
  return c.sc.hyper_buf[c.sc.hyper_ctr++ & 3];
}
__device__ inline double c_norm(DevChain &c, uint64_t seed) {
  const double a = c_unif(c, seed), b = c_unif(c, seed);
  return sqrt(-2.0 * log(a)) * cos(2.0 * PI_D * b);
}
__device__ inline double c_gamma(DevChain &c, uint64_t seed, double shape, double rate) {   // Marsaglia & Tsang 2000
  double boost = 1.0;
  if (shape < 1) { const double u = c_unif(c, seed); boost = pow(u, 1.0 / shape); shape += 1; }
  const double d = shape - 1.0 / 3, cc = 1 / sqrt(9 * d);
  for (;;) {
    const double x = c_norm(c, seed);
    double v = 1 + cc * x;
    if (v <= 0) continue;
    v = v * v * v;
    const double u = c_unif(c, seed);
    if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) return d * v / rate * boost;
  }
}
__device__ inline double c_beta(DevChain &c, uint64_t seed, double a, double b) {
  const double x = c_gamma(c, seed, a, 1), y = c_gamma(c, seed, b, 1);
  return x / (x + y);
}

__device__ inline double dgamma_log(double x, double shape, double rate) {
  if (x < 0) return -INFINITY;
  return shape * log(rate) + (shape - 1) * log(x) - rate * x - lgamma(shape);
}
// dgamma_log(xp) - dgamma_log(xc): the normalising constant (lgamma, log rate) cancels; -Inf proposal outside the support
__device__ inline double dgamma_log_diff(double xp, double xc, double shape, double rate) {
  if (xp < 0) return -INFINITY;
  return (shape - 1) * (log(xp) - log(xc)) - rate * (xp - xc);
}
__device__ inline double dunif_log(double x, double lo, double hi) { return (x >= lo && x <= hi) ? -log(hi - lo) : -INFINITY; }
__device__ inline double dnorm_log(double x, double mu, double sigma) {
  const double z = (x - mu) / sigma;
  return -(0.918938533204672741780329736406 + 0.5 * z * z + log(sigma));     // nmath dnorm, give_log (finite arguments)
}
__device__ inline double scale_move(DevChain &c, uint64_t seed, double param, double t, double &sf) {   // MP:3-13
  sf = exp(t * (c_unif(c, seed) - 0.5));
  return param * sf;
}
__device__ inline double two_boundary_slide(DevChain &c, uint64_t seed, double p, double lo, double hi, double t) {   // MP:29-46
  double np = (p - t) + ((p + t) - (p - t)) * c_unif(c, seed);
  if (np > hi) np = 2 * hi - np;
  return lo + fabs(np - lo);
}
__device__ inline void win_push0(uint8_t *w) { for (int i = 0; i < 99; ++i) w[i] = w[i + 1]; w[99] = 0; }   // c(trace[-1], 0)
__device__ inline int win_sum100(const uint8_t *w) { int s = 0; for (int i = 0; i < 100; ++i) s += w[i]; return s; }
__device__ inline double x_tune_s(double wsum, double t, double target, double lo, double hi) {   // U:161-181
  const double pjump = wsum / 100;
  double nt = t * (tan(PI_D * pjump / 2) / tan(PI_D * target / 2));
  if (nt < lo) nt = lo;
  if (nt > hi) nt = hi;
  return nt;
}

__device__ inline void set_l2_current(DevChain &c, const zz_hyper &h, int K) {
  for (int j = 0; j < 2; ++j) {
    c.l2[j].im = h.inactive_means; c.l2[j].iv = h.inactive_variances;
    for (int k = 0; k < ZZ_MAX_COMP; ++k) { c.l2[j].am[k] = k < K ? h.active_means[k] : 0.0; c.l2[j].av[k] = k < K ? h.active_variances[k] : 1.0; }
  }
}

// ---- the moves, split at the gene sum.  `red` = output of the preceding reduction ([y][2]) ----
__device__ inline void mv_mixture_weights(DevChain &c, zz_hyper &h, uint64_t seed, const double *stats, int K) {   // Pr:412-431
  double g_[ZZ_MAX_COMP + 1], tot = 0;
  for (int k = 0; k <= K; ++k) {
    const double a = (k == 0 ? c.pr.weight_active_shape_2 : c.pr.weight_active_shape_1 / K) + stats[k] * h.beta;
    g_[k] = c_gamma(c, seed, a, 1); tot += g_[k];
  }
  const double nw0 = g_[0] / tot;
  if (!isnan(nw0)) {
    h.weight_active = 1 - nw0;
    double s = 0;
    for (int k = 1; k <= K; ++k) s += g_[k] / tot;
    for (int k = 1; k <= K; ++k) h.weight_within_active[k - 1] = (g_[k] / tot) / s;
  }
}
__device__ inline void mv_spike_prob(DevChain &c, zz_hyper &h, uint64_t seed, const double *stats, int K) {   // Pr:511-526
  const double nin = stats[3 * (K + 1) + 0], ninact = stats[0];
  const double ns = c_beta(c, seed, nin + c.pr.spike_prior_shape_1, (ninact - nin) + c.pr.spike_prior_shape_2);
  if (!isnan(ns)) h.spike_probability = ns;
}

__device__ inline void propose_inactive_means(DevChain &c, const zz_hyper &h, uint64_t seed, int K) {   // Pr:433-471
  const double thr = c.pr.threshold_i, im = h.inactive_means;
  const double imp = thr - fabs(thr - im - c.sc.t_im * c_norm(c, seed));
  c.pp = dgamma_log_diff(thr - imp, thr - im, c.pr.inactive_means_prior_shape, c.pr.inactive_means_prior_rate);   // prior ratio
  c.pc = 0;
  c.prop0 = imp;
  set_l2_current(c, h, K);
  c.l2[0].im = imp;
}
__device__ inline void decide_inactive_means(DevChain &c, zz_hyper &h, uint64_t seed, const double *red, int tune) {
  const double Rr = h.temperature * (red[0] - red[2] + c.pp - c.pc);       // inactive sums: proposed - current
  if (tune) win_push0(c.sc.w_im);
  if (log(c_unif(c, seed)) < Rr) { h.inactive_means = c.prop0; if (tune) c.sc.w_im[99] = 1; }
}

__device__ inline void propose_active_means_dif(DevChain &c, const zz_hyper &h, uint64_t seed, int K) {   // Pr:640-705, one of K rounds
  int k = (int)(c_unif(c, seed) * K); if (k >= K) k = K - 1;
  const double difp = fabs(c.sc.active_means_dif[k] + c.sc.t_am[k] * c_norm(c, seed));
  set_l2_current(c, h, K);
  for (int j = 0; j < K; ++j) c.l2[0].am[j] = (j == k ? difp : c.sc.active_means_dif[j]) + c.pr.threshold_a[j];
  c.pp = dgamma_log_diff(difp, c.sc.active_means_dif[k], c.pr.active_means_dif_prior_shape, c.pr.active_means_dif_prior_rate);
  c.pc = 0;
  c.k = k; c.difp = difp;
}
__device__ inline void decide_active_means_dif(DevChain &c, zz_hyper &h, uint64_t seed, const double *red, int tune, int K) {
  const double Rr = h.temperature * (red[1] + c.pp - red[3] - c.pc);       // active sums
  const int k = c.k;
  if (tune) win_push0(c.sc.w_am[k]);
  if (log(c_unif(c, seed)) < Rr) {
    c.sc.active_means_dif[k] = c.difp; if (tune) c.sc.w_am[k][99] = 1;
    for (int j = 0; j < K; ++j) h.active_means[j] = c.l2[0].am[j];
  }
}

__device__ inline void propose_shared_variance(DevChain &c, const zz_hyper &h, uint64_t seed, int tune, int K) {   // Pr:768-803
  const double lo = c.pr.variances_prior_log_min, hi = c.pr.variances_prior_log_max;
  const double avp1 = pow(10.0, two_boundary_slide(c, seed, log10(h.active_variances[0]), lo, hi, c.sc.t_av));
  if (tune) win_push0(c.sc.w_av);
  set_l2_current(c, h, K);
  c.l2[0].iv = avp1;
  for (int k = 0; k < K; ++k) c.l2[0].av[k] = avp1;
  c.pp = dunif_log(log10(avp1), lo, hi); c.pc = dunif_log(log10(h.active_variances[0]), lo, hi);
  c.prop0 = avp1;
}
__device__ inline void decide_shared_variance(DevChain &c, zz_hyper &h, uint64_t seed, const double *red, int tune, int K) {
  const double Rr = h.temperature * (red[1] + red[0] + c.pp - red[3] - red[2] - c.pc);
  if (log(c_unif(c, seed)) < Rr) {
    for (int k = 0; k < K; ++k) h.active_variances[k] = c.prop0;
    h.inactive_variances = c.prop0; if (tune) c.sc.w_av[99] = 1;
  }
}

__device__ inline double s0_prior(const DevChain &c, double x) {   // P:33-39, up to the constant -log(sigma sqrt(2 pi)) that every use subtracts out
  const double z = (x - c.pr.s0_mu) / c.pr.s0_sigma;
  return -0.5 * z * z;
}
__device__ inline double s1_prior(const DevChain &c, double x) { return dgamma_log(fabs(x), c.pr.s1_shape, c.pr.s1_rate); }   // P:41-47
__device__ inline double tau_prior(const DevChain &c, double x) { return dgamma_log(x, c.pr.tau_shape, c.pr.tau_rate); }       // P:3-9

__device__ inline void propose_tau(DevChain &c, const zz_hyper &h, uint64_t seed) {   // Pr:119-152
  double sf; const double taup = scale_move(c, seed, h.tau, c.sc.t_tau, sf);
  c.HR = log(sf); c.p_s0 = h.s0; c.p_s1 = h.s1; c.p_tau = taup;
  c.pp = dgamma_log_diff(taup, h.tau, c.pr.tau_shape, c.pr.tau_rate); c.pc = 0;   // P:3-9
}
__device__ inline void decide_tau(DevChain &c, zz_hyper &h, uint64_t seed, const double *red, int tune) {
  const double Rr = h.temperature * (red[0] - red[1] + c.pp - c.pc) + c.HR;
  if (tune) win_push0(c.sc.w_tau);
  c.gate_vg = 0;
  if (log(c_unif(c, seed)) < Rr) { h.tau = c.p_tau; if (tune) c.sc.w_tau[99] = 1; c.gate_vg = 1; }
}
__device__ inline void propose_sg(DevChain &c, const zz_hyper &h, uint64_t seed, int tune) {   // Pr:58-117
  double s0p = h.s0, s1p = h.s1; c.HR = 0;
  c.sel = (c_unif(c, seed) < 0.5) ? 1 : 2;
  if (c.sel == 1) { s0p = h.s0 + c.sc.t_s0 * c_norm(c, seed); if (tune) win_push0(c.sc.w_s0); }
  else { double sf; s1p = scale_move(c, seed, h.s1, c.sc.t_s1, sf); c.HR = log(sf); if (tune) win_push0(c.sc.w_s1); }
  c.p_s0 = s0p; c.p_s1 = s1p; c.p_tau = h.tau;
  // only one of s0 / s1 moved: the other prior term is identical on both sides
  c.pp = c.sel == 1 ? s0_prior(c, s0p) - s0_prior(c, h.s0) : dgamma_log_diff(fabs(s1p), fabs(h.s1), c.pr.s1_shape, c.pr.s1_rate);
  c.pc = 0;
}
__device__ inline void decide_sg(DevChain &c, zz_hyper &h, uint64_t seed, const double *red, int tune) {
  const double Rr = h.temperature * (red[0] - red[1] + c.pp - c.pc) + c.HR;
  c.gate_vg = 0;
  if (log(c_unif(c, seed)) < Rr) {
    h.s0 = c.p_s0; h.s1 = c.p_s1;
    if (tune) { if (c.sel == 1) c.sc.w_s0[99] = 1; else c.sc.w_s1[99] = 1; }
    c.gate_vg = 1;
  }
}
__device__ inline void propose_s0tau(DevChain &c, const zz_hyper &h, uint64_t seed, int tune) {   // Pr:154-191
  double sf; const double taup = scale_move(c, seed, h.tau, c.sc.t_s0tau, sf), s0p = h.s0 - log(sf);
  c.HR = log(sf);
  if (tune) win_push0(c.sc.w_s0tau);
  c.p_s0 = s0p; c.p_s1 = h.s1; c.p_tau = taup;
  c.pp = s0_prior(c, s0p) - s0_prior(c, h.s0) + dgamma_log_diff(taup, h.tau, c.pr.tau_shape, c.pr.tau_rate); c.pc = 0;
}
__device__ inline void decide_s0tau(DevChain &c, zz_hyper &h, uint64_t seed, const double *red, int tune) {
  const double Rr = h.temperature * (red[0] - red[1] + c.pp - c.pc) + c.HR;
  c.gate_vg = 0;
  if (log(c_unif(c, seed)) < Rr) { h.s0 = c.p_s0; h.tau = c.p_tau; if (tune) c.sc.w_s0tau[99] = 1; c.gate_vg = 1; }
}
__device__ inline void propose_px(DevChain &c, const zz_hyper &h, uint64_t seed, int tune, int R) {   // Pr:193-243
  int lib = (int)(c_unif(c, seed) * R); if (lib >= R) lib = R - 1;
  double sf; const double ap = scale_move(c, seed, h.alpha_r[lib], c.sc.t_alpha[lib], sf);
  c.HR = log(sf);
  if (tune) win_push0(c.sc.w_alpha[lib]);
  // prior ratio: the reference sums dgamma over all R libraries for both vectors; the R - 1 unchanged terms are identical in the
  // two sums and are left out here (64 logs + 32 lgammas per generation for a single thread otherwise)
  c.pp = dgamma_log_diff(ap, h.alpha_r[lib], c.pr.alpha_r_shape, c.pr.alpha_r_rate);
  c.pc = 0;
  c.p_lib = lib; c.p_alpha = ap;
}
__device__ inline void decide_px(DevChain &c, zz_hyper &h, uint64_t seed, const double *red, int tune) {
  const double Rr = h.temperature * (red[0] + c.pp - c.pc) + c.HR;
  c.gate_px = 0;
  c.p_alpha_old = h.alpha_r[c.p_lib];                   // k_px_commit_dev evaluates ll(alpha') - ll(alpha) from the mailbox (Pr:213-215)
  if (log(c_unif(c, seed)) < Rr && Rr != INFINITY) { c.gate_px = 1; if (tune) c.sc.w_alpha[c.p_lib][99] = 1; }
}
__device__ inline void tune_all_scalars(DevChain &c, double target, int K, int R) {   // U:183-240, scalar part
  zz_chain_scalars &s = c.sc;
  for (int r = 0; r < R; ++r) s.t_alpha[r] = x_tune_s(win_sum100(s.w_alpha[r]), s.t_alpha[r], target, 0.01, 2);
  s.t_s0 = x_tune_s(win_sum100(s.w_s0), s.t_s0, target, 0.0001, 5);
  s.t_s1 = x_tune_s(win_sum100(s.w_s1), s.t_s1, target, 0.0001, 5);
  s.t_tau = x_tune_s(win_sum100(s.w_tau), s.t_tau, target, 0.0001, 5);
  s.t_s0tau = x_tune_s(win_sum100(s.w_s0tau), s.t_s0tau, target, 0.0001, 5);
  s.t_im = x_tune_s(win_sum100(s.w_im), s.t_im, target, 0.001, 10);
  for (int k = 0; k < K; ++k) s.t_am[k] = x_tune_s(win_sum100(s.w_am[k]), s.t_am[k], target, 0.01, 10);
  s.t_av = x_tune_s(win_sum100(s.w_av), s.t_av, target, 0.01, c.pr.variances_prior_log_max - c.pr.variances_prior_log_min);
}

// ---- closed forms of the gene sums from the sweep's sufficient statistics (layout: zz_suffstats in include/zigzag_gpu.h) ----
// sum over inactive genes of computeInactiveLikelihood and over active genes of computeActiveLikelihood (P:210-273) are quadratic
// in the per-component (n, sum y, sum y^2); the sum of computeVarianceGPriorProbability over out-of-spike genes (P:11-31, no upper
// bound) is quadratic in (n, sum z, sum z^2, sum y, sum y^2, sum y z), z = log variance_g.  With them a hyper move needs no pass
// over the genes at all.
__device__ inline void closed_level2(const zz_hyper &h, const double *st, int K, const L2Params &p, double *out2) {
  const int K1 = K + 1;
  const double n_in = st[3 * K1], n0out = st[0] - n_in, s2p = sqrt(2 * PI_D);
  out2[0] = n_in * log(h.spike_probability) + n0out * (log(1 - h.spike_probability) - log(s2p * sqrt(p.iv)))
            - (st[2 * K1] - 2 * p.im * st[K1] + n0out * p.im * p.im) / (2 * p.iv);
  double la = 0;
  for (int k = 1; k <= K; ++k) {
    const double n = st[k], sy = st[K1 + k], sy2 = st[2 * K1 + k], am = p.am[k - 1], av = p.av[k - 1];
    la += -n * log(s2p * sqrt(av)) - (sy2 - 2 * am * sy + n * am * am) / (2 * av);
  }
  out2[1] = la;
}
// The same sums as DIFFERENCES proposed - current (what the moves need): the logarithms cancel unless a variance changes, and a
// single thread pays ~0.4 us per double-precision libm call.  r[0] = inactive part, r[1] = active part; r[2] = r[3] = 0.
__device__ inline void closed_level2_diff(const double *st, int K, const L2Params &p, const L2Params &c, double *r) {
  const int K1 = K + 1;
  const double n0out = st[0] - st[3 * K1];
  const double qp = st[2 * K1] - 2 * p.im * st[K1] + n0out * p.im * p.im, qc = st[2 * K1] - 2 * c.im * st[K1] + n0out * c.im * c.im;
  double li = -(qp / (2 * p.iv) - qc / (2 * c.iv));
  if (p.iv != c.iv) li += -0.5 * n0out * (log(p.iv) - log(c.iv));
  double la = 0;
  for (int k = 1; k <= K; ++k) {
    const double n = st[k], sy = st[K1 + k], sy2 = st[2 * K1 + k];
    const double ap = p.am[k - 1], ac = c.am[k - 1], vp = p.av[k - 1], vc = c.av[k - 1];
    la += -((sy2 - 2 * ap * sy + n * ap * ap) / (2 * vp) - (sy2 - 2 * ac * sy + n * ac * ac) / (2 * vc));
    if (vp != vc) la += -0.5 * n * (log(vp) - log(vc));
  }
  r[0] = li; r[1] = la; r[2] = 0; r[3] = 0;
}
__device__ inline double closed_varg(const double *st, int K, double s0, double s1, double tau) {
  const double *fx = st + 3 * (K + 1);
  const double n = fx[1], sz = fx[2], sz2 = fx[3], syo = fx[4], sy2o = fx[5], syz = fx[6], c = s0 + tau;
  const double q = sz2 + n * c * c + s1 * s1 * sy2o - 2 * c * sz - 2 * s1 * syz + 2 * c * s1 * syo;
  return -sz - n * log(sqrt(tau) * sqrt(2 * PI_D)) - q / (2 * tau);
}

// Stages of one generation after the fused sweep (hyper slots 1, 4, 6, 7, 8, 12, 13, 14, 15 of proposal_list, I:207-223):
enum ChainStage {
  ST_BEGIN = 0,     // gibbsMixtureWeights; propose mhInactiveMeans                      -> level-2 sums
  ST_AMD,           // decide the pending level-2 move (iter 0: mhInactiveMeans, else mhActiveMeansDif); propose mhActiveMeansDif -> level-2 sums
  ST_SHARED,        // decide the last mhActiveMeansDif; propose mhSharedVariance         -> level-2 sums
  ST_TAU,           // decide mhSharedVariance; gibbsSpikeProb; propose mhTau             -> variance-prior sums
  ST_SG,            // decide mhTau (gate commit); propose mhSg                           -> commit, variance-prior sums
  ST_S0TAU,         // decide mhSg (gate commit); propose mhS0Tau                         -> commit, variance-prior sums
  ST_PX,            // decide mhS0Tau (gate commit); propose mhP_x                        -> commit, detection delta of one library
  ST_PX_DECIDE,     // decide mhP_x (gate commit)                                         -> commit XgLikelihood
  ST_END,           // install the accepted alpha_r; counters; tune_all (scalar part)
  ST_ALL_CLOSED     // slots 1, 4, 6, 7, 8, 12, 13, 14 in one go from the closed forms, then propose mhP_x -> detection delta
};

__global__ void k_chain_stage(int stage, int iter, int tune, int K, int R, uint64_t seed, DevChain *cp_all, zz_hyper *hp_all,
                              const double *stats_all, const double *red_all, int red_stride, double target, int do_tune_all, int sampled,
                              FastH *fasth_all) {
  // one block per chain of the handle (batched chains run their scalar stages side by side)
  DevChain *cp = cp_all + blockIdx.x; zz_hyper *hp = hp_all + blockIdx.x;
  const double *stats_g = stats_all + (size_t)blockIdx.x * (3 * (K + 1) + 10), *red = red_all + (size_t)blockIdx.x * red_stride;
  // One thread does the scalar work, but on a shared-memory copy of the chain state: the moves touch a few thousand bytes (100-wide
  // windows shifted element by element, Philox buffer, mailbox) and a single thread pays the full DRAM/L2 latency for each access
  // (measured: 150 us per generation on the global-memory version, most of it window shifts).
  __shared__ DevChain c_s;
  __shared__ zz_hyper h_s;
  __shared__ double stats[3 * (ZZ_MAX_COMP + 1) + 10];
  static_assert(sizeof(DevChain) % 8 == 0 && sizeof(zz_hyper) % 8 == 0, "copied as 64-bit words");
  {
    const uint64_t *s0 = reinterpret_cast<const uint64_t *>(cp), *s1 = reinterpret_cast<const uint64_t *>(hp);
    uint64_t *d0 = reinterpret_cast<uint64_t *>(&c_s), *d1 = reinterpret_cast<uint64_t *>(&h_s);
    for (int j = threadIdx.x; j < (int)(sizeof(DevChain) / 8); j += blockDim.x) d0[j] = s0[j];
    for (int j = threadIdx.x; j < (int)(sizeof(zz_hyper) / 8); j += blockDim.x) d1[j] = s1[j];
    for (int j = threadIdx.x; j < 3 * (K + 1) + 10; j += blockDim.x) stats[j] = stats_g[j];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
  DevChain &c = c_s; zz_hyper &h = h_s;
  switch (stage) {
    case ST_BEGIN:
      mv_mixture_weights(c, h, seed, stats, K);
      propose_inactive_means(c, h, seed, K);
      break;
    case ST_AMD:
      if (iter == 0) decide_inactive_means(c, h, seed, red, tune); else decide_active_means_dif(c, h, seed, red, tune, K);
      propose_active_means_dif(c, h, seed, K);
      break;
    case ST_SHARED:
      decide_active_means_dif(c, h, seed, red, tune, K);
      propose_shared_variance(c, h, seed, tune, K);
      break;
    case ST_TAU:
      decide_shared_variance(c, h, seed, red, tune, K);
      mv_spike_prob(c, h, seed, stats, K);
      propose_tau(c, h, seed);
      break;
    case ST_SG: decide_tau(c, h, seed, red, tune); propose_sg(c, h, seed, tune); break;
    case ST_S0TAU: decide_sg(c, h, seed, red, tune); propose_s0tau(c, h, seed, tune); break;
    case ST_PX: decide_s0tau(c, h, seed, red, tune); propose_px(c, h, seed, tune, R); break;
    case ST_PX_DECIDE:                                   // ... and the end of the generation in the same launch
      decide_px(c, h, seed, red, tune);
      if (c.gate_px) h.alpha_r[c.p_lib] = c.p_alpha;
      c.sc.step += 10;                                   // the sweep and nine hyper slots, one `step` each
      c.sc.gen += 1;
      if (sampled) c.sc.n_samples += 1;
      if (do_tune_all) tune_all_scalars(c, target, K, R);
      derive_fast(fasth_all[blockIdx.x], &h);            // the sweep kernels' derived constants (k_prepare_hyper, fused here)
      break;
    case ST_ALL_CLOSED: {
      double r[4];
      mv_mixture_weights(c, h, seed, stats, K);
      propose_inactive_means(c, h, seed, K);
      closed_level2_diff(stats, K, c.l2[0], c.l2[1], r);
      decide_inactive_means(c, h, seed, r, tune);
      for (int it = 0; it < K; ++it) {
        propose_active_means_dif(c, h, seed, K);
        closed_level2_diff(stats, K, c.l2[0], c.l2[1], r);
        decide_active_means_dif(c, h, seed, r, tune, K);
      }
      propose_shared_variance(c, h, seed, tune, K);
      closed_level2_diff(stats, K, c.l2[0], c.l2[1], r);
      decide_shared_variance(c, h, seed, r, tune, K);
      mv_spike_prob(c, h, seed, stats, K);
      int any = 0;
      propose_tau(c, h, seed);
      r[0] = closed_varg(stats, K, c.p_s0, c.p_s1, c.p_tau); r[1] = closed_varg(stats, K, h.s0, h.s1, h.tau);
      decide_tau(c, h, seed, r, tune); any |= c.gate_vg;
      propose_sg(c, h, seed, tune);
      r[0] = closed_varg(stats, K, c.p_s0, c.p_s1, c.p_tau); r[1] = closed_varg(stats, K, h.s0, h.s1, h.tau);
      decide_sg(c, h, seed, r, tune); any |= c.gate_vg;
      propose_s0tau(c, h, seed, tune);
      r[0] = closed_varg(stats, K, c.p_s0, c.p_s1, c.p_tau); r[1] = closed_varg(stats, K, h.s0, h.s1, h.tau);
      decide_s0tau(c, h, seed, r, tune); any |= c.gate_vg;
      c.gate_vg = any;                                   // one commit of variance_g_probability for the three moves
      propose_px(c, h, seed, tune, R);
      break;
    }
    case ST_END:
      if (c.gate_px) h.alpha_r[c.p_lib] = c.p_alpha;
      c.sc.step += 10;                                   // the sweep and nine hyper slots, one `step` each
      c.sc.gen += 1;
      if (sampled) c.sc.n_samples += 1;
      if (do_tune_all) tune_all_scalars(c, target, K, R);
      break;
  }
  }   // thread 0
  __syncthreads();
  {
    uint64_t *d0 = reinterpret_cast<uint64_t *>(cp), *d1 = reinterpret_cast<uint64_t *>(hp);
    const uint64_t *s0 = reinterpret_cast<const uint64_t *>(&c_s), *s1 = reinterpret_cast<const uint64_t *>(&h_s);
    for (int j = threadIdx.x; j < (int)(sizeof(DevChain) / 8); j += blockDim.x) d0[j] = s0[j];
    for (int j = threadIdx.x; j < (int)(sizeof(zz_hyper) / 8); j += blockDim.x) d1[j] = s1[j];
  }
}

}  // namespace zz
