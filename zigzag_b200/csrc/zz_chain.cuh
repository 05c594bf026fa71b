// zz_chain.cuh — the scalar half of a generation, on the device.
//
// The reference's ten hyper-parameter moves (R/zigzagProposals.R) are scalar Metropolis / Gibbs moves whose only O(G) part is a
// sum over genes.  In the device-resident loop (k_generations, zz_persist.cuh) those sums are closed forms of the sufficient
// statistics the per-gene sweep has just accumulated, so after the grid barrier ONE CTA — the last one to arrive — runs all of them
// and publishes the new hyper-parameters (and the constants derived from them) to the rest of the grid.  The moves fall into
// groups that do not depend on each other within a generation and run side by side on different warps of that CTA:
//   warp 0  gibbsMixtureWeights (Pr:412-431) and gibbsSpikeProb (Pr:511-526): K + 3 gamma draws, one lane each
//   warp 1  mhInactiveMeans -> mhActiveMeansDif (K rounds) -> mhSharedVariance   (Pr:433-471, 640-705, 768-803)
//   warp 2  mhTau -> mhSg -> mhS0Tau                                              (Pr:119-152, 58-117, 154-191)
//   warp 3  the decision of the pending mhP_x (Pr:193-243; its gene sum was accumulated by the sweep)
// and afterwards tune_all (U:183-240, scalar part) and the proposal of the next generation's mhP_x.
// Scalar draws are counter-based: move `slot` of a generation reads the Philox blocks (gene 0xFFFFFF00 + slot, blk 0, 1, ...,
// step = the step of that generation's sweep) in order (RNG) — the map the CPU test chain implements too, so a
// device-resident segment is replayed on the CPU decision for decision (tests/test_gpu_chain.py).
#pragma once
#include "zz_device.cuh"
#include "zz_fast.cuh"

namespace zz {

enum { SUB_MIXW = 0x00 /* + component 0..K */, SUB_SP0 = 0x10, SUB_SP1 = 0x11, SUB_IM = 0x20, SUB_AMD = 0x21, SUB_SV = 0x22,
       SUB_TAU = 0x23, SUB_SG = 0x24, SUB_S0TAU = 0x25, SUB_PX = 0x26, SUB_SWAP = 0x30 };

// 100-wide accept window of a scalar move (the *_trace fields) as 2 x u64, bit 0 = newest entry (like the per-gene windows)
struct Win100 { unsigned long long lo, hi; };
__host__ __device__ inline void win_push0(Win100 &w) {                     // c(trace[-1], 0)
  w.hi = ((w.hi << 1) | (w.lo >> 63)) & ((1ULL << 36) - 1);
  w.lo = w.lo << 1;
}
__host__ __device__ inline void win_set_last(Win100 &w) { w.lo |= 1ULL; }  // trace[100] <- 1
__device__ inline int win_sum100(const Win100 &w) { return __popcll(w.lo) + __popcll(w.hi & ((1ULL << 36) - 1)); }

// scalar chain state as the device keeps it (zz_chain_scalars with the windows packed) + the pending mhP_x proposal
struct ChainDev {
  zz_priors pr;
  double active_means_dif[ZZ_MAX_COMP];
  double t_im, t_av, t_am[ZZ_MAX_COMP], t_s0, t_s1, t_tau, t_s0tau, t_alpha[ZZ_MAX_LIBS];
  Win100 w_im, w_av, w_am[ZZ_MAX_COMP], w_s0, w_s1, w_tau, w_s0tau, w_alpha[ZZ_MAX_LIBS];
  uint64_t hyper_ctr; double hyper_buf[4];     // position in the sequential stream (constructor, reference schedule): carried, not used
  uint64_t step; int64_t gen, n_samples;
  // mhP_x of the generation whose sweep runs next: proposed before the sweep, which accumulates its gene sum (Pr:209-224)
  int px_lib, px_valid; double px_alpha, px_HR, px_pp;
  int px_prev_acc, pad_;                       // the previous generation accepted its mhP_x: the sweep adds the stored per-gene delta to A(y)
  int chain, pad2_;                            // index of this chain in the handle (chain field of its Philox counters)
};

// The scalar moves run on single lanes, where a libm log / exp is a chain of ~120 dependent FP64 instructions at 9 cycles each
// (0.7 us; 45 us per generation with libm throughout, measured): they use the sweep's table-driven versions (<= 2 ulp) instead.
struct SMath {
  const MathTabs *T;
  __device__ double log(double x) const { return flog(x, T->log_t); }
  __device__ double exp(double x) const { return fexp(x, T->exp_t); }
};

// ---- scalar draws -----------------------------------------------------------------------------------------------------
struct SubStream {                             // sequential consumption of one move's blocks
  uint64_t seed, step; uint32_t chain, slot, blk; int pos; double buf[4];
  __device__ void begin(uint64_t seed_, uint32_t chain_, uint64_t step_, uint32_t slot_) { seed = seed_; chain = chain_; step = step_; slot = slot_; blk = 0; pos = 4; }
  __device__ double unif() {
    if (pos == 4) {
      uint32_t o[4];
      philox_block(seed, chain, step, blk++ & 255u, 0xFFFFFF00ull + slot, o);
      for (int j = 0; j < 4; ++j) buf[j] = u32_open(o[j]);
      pos = 0;
    }
    return buf[pos++];
  }
  __device__ double norm(const SMath &m) { const double a = unif(), b = unif(); return sqrt(-2.0 * m.log(a)) * cospi(2.0 * b); }
};

// The draws of one Metropolis move = block `blk` of its sub-stream, prepared ahead of the (sequential) chain of moves by a lane of
// its own: the four uniforms, the normals the move may form from (w0, w1) or (w1, w2), and the logarithms of w1..w3 (one of
// which is the move's accept uniform).
struct MoveDraws { double u[4], n01, n12, l1, l2, l3; };
enum { MD_IM = 0, MD_SV = 1, MD_TAU = 2, MD_SG = 3, MD_S0TAU = 4, MD_AMD = 5 /* + round */, MD_MAX = 5 + ZZ_MAX_COMP };
__device__ inline void make_move_draws(const SMath &m, MoveDraws &d, uint64_t seed, uint32_t chain, uint64_t step, int idx) {
  static const uint32_t slot_of[5] = {SUB_IM, SUB_SV, SUB_TAU, SUB_SG, SUB_S0TAU};
  const uint32_t slot = idx < MD_AMD ? slot_of[idx] : SUB_AMD, blk = idx < MD_AMD ? 0u : (uint32_t)(idx - MD_AMD);
  uint32_t o[4];
  philox_block(seed, chain, step, blk, 0xFFFFFF00ull + slot, o);
  for (int j = 0; j < 4; ++j) d.u[j] = u32_open(o[j]);
  const double l0 = m.log(d.u[0]);
  d.l1 = m.log(d.u[1]); d.l2 = m.log(d.u[2]); d.l3 = m.log(d.u[3]);
  d.n01 = sqrt(-2.0 * l0) * cospi(2.0 * d.u[1]);
  d.n12 = sqrt(-2.0 * d.l1) * cospi(2.0 * d.u[2]);
}

// Marsaglia & Tsang 2000 on a sub-stream of its own: attempt j uses block j (w0, w1 -> normal, w2 -> accept), the boost uniform of
// shape < 1 is w3 of block 0
__device__ inline double gamma_slot(const SMath &m, uint64_t seed, uint32_t chain, uint64_t step, uint32_t slot, double shape, double rate) {
  uint32_t o[4];
  double boost = 1;
  if (shape < 1) { philox_block(seed, chain, step, 0, 0xFFFFFF00ull + slot, o); boost = pow(u32_open(o[3]), 1.0 / shape); shape += 1; }
  const double d = shape - 1.0 / 3, cc = 1 / sqrt(9 * d);
  for (uint32_t j = 0; j < 255; ++j) {
    philox_block(seed, chain, step, j, 0xFFFFFF00ull + slot, o);
    const double x = sqrt(-2.0 * m.log(u32_open(o[0]))) * cospi(2.0 * u32_open(o[1]));
    double v = 1 + cc * x;
    if (v <= 0) continue;
    v = v * v * v;
    if (m.log(u32_open(o[2])) < 0.5 * x * x + d - d * v + d * m.log(v)) return rate == 1.0 ? d * v * boost : d * v / rate * boost;
  }
  return d / rate * boost;
}

// The shape-independent part of attempts 0 and 1 of gamma_slot — the normal x and the logarithm of the accept uniform — can be
// drawn before the shape (a gene count) is known: GammaPre holds them, gamma_slot_fin finishes the draw with the same arithmetic.
struct GammaPre { double x[2], lu[2]; };
__device__ inline void gamma_predraw(const SMath &m, uint64_t seed, uint32_t chain, uint64_t step, uint32_t slot, int j, GammaPre &g) {
  uint32_t o[4];
  philox_block(seed, chain, step, (uint32_t)j, 0xFFFFFF00ull + slot, o);
  g.x[j] = sqrt(-2.0 * m.log(u32_open(o[0]))) * cospi(2.0 * u32_open(o[1]));
  g.lu[j] = m.log(u32_open(o[2]));
}
__device__ inline double gamma_slot_fin(const SMath &m, uint64_t seed, uint32_t chain, uint64_t step, uint32_t slot, double shape, double rate, const GammaPre &g) {
  if (shape < 1) return gamma_slot(m, seed, chain, step, slot, shape, rate);      // boosted draw: rare (beta = 0 with a small prior shape)
  const double d = shape - 1.0 / 3, cc = 1 / sqrt(9 * d);
  for (uint32_t j = 0; j < 255; ++j) {
    double x, lu;
    if (j < 2) { x = g.x[j]; lu = g.lu[j]; }
    else {
      uint32_t o[4];
      philox_block(seed, chain, step, j, 0xFFFFFF00ull + slot, o);
      x = sqrt(-2.0 * m.log(u32_open(o[0]))) * cospi(2.0 * u32_open(o[1]));
      lu = m.log(u32_open(o[2]));
    }
    double v = 1 + cc * x;
    if (v <= 0) continue;
    v = v * v * v;
    if (lu < 0.5 * x * x + d - d * v + d * m.log(v)) return rate == 1.0 ? d * v : d * v / rate;
  }
  return d / rate;
}

// dgamma_log(xp) - dgamma_log(xc): the normalising constant (lgamma, log rate) cancels; -Inf proposal outside the support
template <class M>
__device__ inline double dgamma_log_diff(const M &m, double xp, double xc, double shape, double rate) {
  if (xp < 0) return -INFINITY;
  return (shape - 1) * (m.log(xp) - m.log(xc)) - rate * (xp - xc);
}
struct LibM { __device__ double log(double x) const { return ::log(x); } __device__ double exp(double x) const { return ::exp(x); } };
// dunif(x, lo, hi, log = TRUE): only its finiteness can differ between the proposed and the current value
__device__ inline double dunif_log_rel(double x, double lo, double hi) { return (x >= lo && x <= hi) ? 0.0 : -INFINITY; }
__device__ inline double x_tune_s(double wsum, double t, double target, double lo, double hi) {   // U:161-181
  const double pjump = wsum / 100;
  double nt = t * (tan(PI_D * pjump / 2) / tan(PI_D * target / 2));
  if (nt < lo) nt = lo;
  if (nt > hi) nt = hi;
  return nt;
}

// ---- closed forms of the gene sums from the sweep's sufficient statistics (layout: zz_suffstats in include/zigzag_gpu.h) ----
// The sums over inactive genes of computeInactiveLikelihood and over active genes of computeActiveLikelihood (P:210-273) are
// quadratic in the per-component (n, sum y, sum y^2); the sum of computeVarianceGPriorProbability over out-of-spike genes
// (P:11-31, no upper bound) is quadratic in (n, sum z, sum z^2, sum y, sum y^2, sum y z), z = log variance_g.  The moves need the
// DIFFERENCE proposed - current, in which the logarithms cancel unless a variance changes.
struct L2Params { double im, iv, am[ZZ_MAX_COMP], av[ZZ_MAX_COMP]; };
__device__ inline void l2_current(L2Params &p, const zz_hyper &h, int K) {
  p.im = h.inactive_means; p.iv = h.inactive_variances;
  for (int k = 0; k < K; ++k) { p.am[k] = h.active_means[k]; p.av[k] = h.active_variances[k]; }
}
// A single lane pays ~25 dependent FP64 instructions (0.12 us) per division: the moves carry the reciprocals 1 / (2 variance) of the
// CURRENT parameters along (RcpL2, refreshed when a variance is accepted) and divide only for a proposed variance.
struct RcpL2 { double iv, av[ZZ_MAX_COMP]; };           // 1 / (2 inactive_variances), 1 / (2 active_variances[k])
// inactive part of (sum at p) - (sum at c): computeInactiveLikelihood over the inactive genes (P:210-238)
__device__ inline double closed_inactive_diff(const SMath &m, const double *st, int K, double imp, double ivp, double rivp, double imc, double ivc, double rivc) {
  const int K1 = K + 1;
  const double n0out = st[0] - st[3 * K1];
  const double qp = st[2 * K1] - 2 * imp * st[K1] + n0out * imp * imp, qc = st[2 * K1] - 2 * imc * st[K1] + n0out * imc * imc;
  double li = -(qp * rivp - qc * rivc);
  if (ivp != ivc) li += -0.5 * n0out * (m.log(ivp) - m.log(ivc));
  return li;
}
// active part: computeActiveLikelihood over the active genes (P:240-273); `ldv` = log(vp) - log(vc) when every component shares them
template <int K>
__device__ inline double closed_active_diff(const double *st, const double *amp, const double *rvp, const double *amc, const double *rvc, double ldv) {
  constexpr int K1 = K + 1;
  double la = 0;
#pragma unroll
  for (int k = 1; k <= K; ++k) {
    const double n = st[k], sy = st[K1 + k], sy2 = st[2 * K1 + k];
    const double ap = amp[k - 1], ac = amc[k - 1];
    la += -((sy2 - 2 * ap * sy + n * ap * ap) * rvp[k - 1] - (sy2 - 2 * ac * sy + n * ac * ac) * rvc[k - 1]) - 0.5 * n * ldv;
  }
  return la;
}
// (sum of computeVarianceGPriorProbability at (s0p, s1p, taup)) - (the same at (s0, s1, tau)); rtp = 1 / (2 taup), rt = 1 / (2 tau)
__device__ inline double closed_varg_diff(const SMath &m, const double *st, int K, double s0p, double s1p, double taup, double rtp, double s0, double s1, double tau, double rt) {
  const double *fx = st + 3 * (K + 1);
  const double n = fx[1], sz = fx[2], sz2 = fx[3], syo = fx[4], sy2o = fx[5], syz = fx[6];
  const double cp = s0p + taup, cc = s0 + tau;
  const double qp = sz2 + n * cp * cp + s1p * s1p * sy2o - 2 * cp * sz - 2 * s1p * syz + 2 * cp * s1p * syo;
  const double qc = sz2 + n * cc * cc + s1 * s1 * sy2o - 2 * cc * sz - 2 * s1 * syz + 2 * cc * s1 * syo;
  double d = -(qp * rtp - qc * rt);
  if (taup != tau) d += -0.5 * n * (m.log(taup) - m.log(tau));           // -n log(sqrt(tau) sqrt(2 pi)), P:20
  return d;
}
__device__ inline double s0_prior_rel(const zz_priors &pr, double x) {   // P:33-39 up to the constant every use subtracts out
  const double z = (x - pr.s0_mu) / pr.s0_sigma;
  return -0.5 * z * z;
}

// ---- the move groups (each runs on lane 0 of its own warp; `h` is the shared-memory copy of the chain's hyper-parameters) ----

// warp 0, lanes 0 .. K + 2: the gamma draws of gibbsMixtureWeights (Pr:412-431) and gibbsSpikeProb (Pr:511-526); `g` is a
// shared-memory array of K + 3 doubles that lane 0 combines afterwards (stage_gibbs_combine)
__device__ inline uint32_t gibbs_slot(int idx, int K) { return idx <= K ? (uint32_t)(SUB_MIXW + idx) : idx == K + 1 ? (uint32_t)SUB_SP0 : (uint32_t)SUB_SP1; }
// the part that does not need the statistics: lanes 0 .. 2 (K + 3) - 1, lane = 2 * draw + attempt
__device__ inline void stage_gibbs_predraw(const SMath &m, const ChainDev &c, uint64_t seed, uint64_t step, int K, int lane, GammaPre *gp) {
  if (lane >= 2 * (K + 3)) return;
  gamma_predraw(m, seed, (uint32_t)c.chain, step, gibbs_slot(lane >> 1, K), lane & 1, gp[lane >> 1]);
}
__device__ inline void stage_gibbs_draws(const SMath &m, const ChainDev &c, const zz_hyper &h, uint64_t seed, uint64_t step, const double *stats, int K, int lane, const GammaPre *gp, double *g) {
  if (lane > K + 2) return;
  double shape;
  if (lane <= K) shape = (lane == 0 ? c.pr.weight_active_shape_2 : c.pr.weight_active_shape_1 / K) + stats[lane] * h.beta;
  else {
    const double nin = stats[3 * (K + 1) + 0], ninact = stats[0];
    shape = lane == K + 1 ? nin + c.pr.spike_prior_shape_1 : (ninact - nin) + c.pr.spike_prior_shape_2;
  }
  g[lane] = gamma_slot_fin(m, seed, (uint32_t)c.chain, step, gibbs_slot(lane, K), shape, 1.0, gp[lane]);
}
__device__ inline void stage_gibbs_combine(zz_hyper &h, int K, const double *g) {
  double tot = 0, act = 0;
  for (int k = 0; k <= K; ++k) tot += g[k];
  for (int k = 1; k <= K; ++k) act += g[k];
  const double nw0 = g[0] / tot;                                          // rdirichlet: gamma draws over their sum (Pr:420-428)
  if (!isnan(nw0)) {
    h.weight_active = 1 - nw0;
    const double ract = 1.0 / act;                                        // (g_k / tot) / (sum_{j >= 1} g_j / tot)
    for (int k = 1; k <= K; ++k) h.weight_within_active[k - 1] = g[k] * ract;
  }
  const double ns = g[K + 1] / (g[K + 1] + g[K + 2]);                     // rbeta as a ratio of gammas
  if (!isnan(ns)) h.spike_probability = ns;
}

// ---- the Metropolis moves, split in two -----------------------------------------------------------------------------------
// Everything of a move except its gene sum — proposal, Hastings ratio, prior ratio, the logarithms of a changed variance — depends
// only on the draws and on parameters that are known BEFORE the sweep's statistics are: stage_*_pre computes it (on the steward CTA:
// while the other CTAs sweep), stage_*_post adds the closed-form gene sum and decides.  Moves later in a chain depend on whether
// earlier ones were accepted: their pre-parts are tabulated for every outcome that can reach them.  The arithmetic of each
// expression is unchanged (same operands, same order), so the decisions are those of the one-piece moves.

// closed_varg_diff with the variance logarithms handed in: ldt = m.log(taup) - m.log(tau)
__device__ inline double closed_varg_diff_pre(const double *st, int K, double s0p, double s1p, double taup, double rtp, double s0, double s1, double tau, double rt, double ldt) {
  const double *fx = st + 3 * (K + 1);
  const double n = fx[1], sz = fx[2], sz2 = fx[3], syo = fx[4], sy2o = fx[5], syz = fx[6];
  const double cp = s0p + taup, cc = s0 + tau;
  const double qp = sz2 + n * cp * cp + s1p * s1p * sy2o - 2 * cp * sz - 2 * s1p * syz + 2 * cp * s1p * syo;
  const double qc = sz2 + n * cc * cc + s1 * s1 * sy2o - 2 * cc * sz - 2 * s1 * syz + 2 * cc * s1 * syo;
  double d = -(qp * rtp - qc * rt);
  if (taup != tau) d += -0.5 * n * ldt;                                  // -n log(sqrt(tau) sqrt(2 pi)), P:20
  return d;
}

// mhTau (Pr:119-152) -> mhSg (Pr:58-117) -> mhS0Tau (Pr:154-191)
struct PreVarg {
  double rt0;                                            // 1 / (2 tau)
  double t_HR, t_taup, t_rtp, t_pp, t_ldt;               // mhTau: scale w0, accept w1
  int g_sel, pad_; double g_s0p, g_s1p, g_HR, g_pp;      // mhSg: selector w0; s0: z = (w1, w2), accept w3; s1: scale w1, accept w2
  // mhS0Tau: scale w0, accept w1; [v]: tau is still h.tau (0) / mhTau was accepted (1); [w]: s0 is still h.s0 (0) / mhSg moved it (1)
  double z_HR, z_taup[2], z_rtp[2], z_dgam[2], z_ldt[2], z_s0p[2], z_s0rel[2];
};
__device__ inline void stage_varg_pre(const SMath &m, const ChainDev &c, const zz_hyper &h, const MoveDraws *md, PreVarg &P) {
  const double s0 = h.s0, s1 = h.s1, tau = h.tau;
  P.rt0 = 1.0 / (2 * tau);
  {
    const MoveDraws &d = md[MD_TAU];
    P.t_HR = c.t_tau * (d.u[0] - 0.5); P.t_taup = tau * m.exp(P.t_HR); P.t_rtp = 1.0 / (2 * P.t_taup);   // scale_move, MP:3-13; HR = log(sf)
    P.t_pp = dgamma_log_diff(m, P.t_taup, tau, c.pr.tau_shape, c.pr.tau_rate);                            // P:3-9
    P.t_ldt = m.log(P.t_taup) - m.log(tau);
  }
  {
    const MoveDraws &d = md[MD_SG];
    P.g_s0p = s0; P.g_s1p = s1; P.g_HR = 0;
    P.g_sel = (d.u[0] < 0.5) ? 1 : 2;
    if (P.g_sel == 1) { P.g_s0p = s0 + c.t_s0 * d.n12; P.g_pp = s0_prior_rel(c.pr, P.g_s0p) - s0_prior_rel(c.pr, s0); }
    else {
      P.g_HR = c.t_s1 * (d.u[1] - 0.5); P.g_s1p = s1 * m.exp(P.g_HR);
      P.g_pp = dgamma_log_diff(m, fabs(P.g_s1p), fabs(s1), c.pr.s1_shape, c.pr.s1_rate);                  // P:41-47 (the unchanged s0 term cancels)
    }
  }
  {
    const MoveDraws &d = md[MD_S0TAU];
    P.z_HR = c.t_s0tau * (d.u[0] - 0.5);
    const double e = m.exp(P.z_HR);
    for (int v = 0; v < 2; ++v) {
      const double tv = v ? P.t_taup : tau;
      P.z_taup[v] = tv * e; P.z_rtp[v] = 1.0 / (2 * P.z_taup[v]);
      P.z_dgam[v] = dgamma_log_diff(m, P.z_taup[v], tv, c.pr.tau_shape, c.pr.tau_rate);
      P.z_ldt[v] = m.log(P.z_taup[v]) - m.log(tv);
    }
    for (int w = 0; w < 2; ++w) {
      const double sw = w ? P.g_s0p : s0;
      P.z_s0p[w] = sw - P.z_HR; P.z_s0rel[w] = s0_prior_rel(c.pr, P.z_s0p[w]) - s0_prior_rel(c.pr, sw);
    }
  }
}
__device__ inline void stage_varg_post(const PreVarg &P, ChainDev &c, const zz_hyper &h, const MoveDraws *md, const double *st, int K, int tune, double &s0, double &s1, double &tau) {
  s0 = h.s0; s1 = h.s1; tau = h.tau;
  double rt = P.rt0;
  int v = 0, w = 0;
  {
    const double Rr = h.temperature * (closed_varg_diff_pre(st, K, s0, s1, P.t_taup, P.t_rtp, s0, s1, tau, rt, P.t_ldt) + P.t_pp) + P.t_HR;
    if (tune) win_push0(c.w_tau);
    if (md[MD_TAU].l1 < Rr) { tau = P.t_taup; rt = P.t_rtp; v = 1; if (tune) win_set_last(c.w_tau); }
  }
  {
    const MoveDraws &d = md[MD_SG];
    const int sel = P.g_sel;
    if (tune) { if (sel == 1) win_push0(c.w_s0); else win_push0(c.w_s1); }
    const double Rr = h.temperature * (closed_varg_diff_pre(st, K, P.g_s0p, P.g_s1p, tau, rt, s0, s1, tau, rt, 0.0) + P.g_pp) + P.g_HR;
    if ((sel == 1 ? d.l3 : d.l2) < Rr) {
      s0 = P.g_s0p; s1 = P.g_s1p; w = sel == 1;
      if (tune) { if (sel == 1) win_set_last(c.w_s0); else win_set_last(c.w_s1); }
    }
  }
  {
    if (tune) win_push0(c.w_s0tau);
    const double pp = P.z_s0rel[w] + P.z_dgam[v];
    const double Rr = h.temperature * (closed_varg_diff_pre(st, K, P.z_s0p[w], s1, P.z_taup[v], P.z_rtp[v], s0, s1, tau, rt, P.z_ldt[v]) + pp) + P.z_HR;
    if (md[MD_S0TAU].l1 < Rr) { s0 = P.z_s0p[w]; tau = P.z_taup[v]; if (tune) win_set_last(c.w_s0tau); }
  }
}

// mhInactiveMeans (Pr:433-471): z = (w0, w1), accept w2 | mhActiveMeansDif x K (Pr:640-705): round `it`: k = w0, z = (w1, w2), accept w3 |
// mhSharedVariance (Pr:768-803): slide w0, accept w1.  The first two touch disjoint parameters; the third runs at the means they left.
struct PreL2 {
  double riv, imp, im_pp;                                // mhInactiveMeans
  double rav[ZZ_MAX_COMP];                               // 1 / (2 active_variances[k])
  int amd_k[ZZ_MAX_COMP];                                // component of round it
  // round it, given the set `mask` of earlier rounds that were accepted (only those on the same component matter): proposed dif, prior ratio
  double amd_difp[3][4], amd_pp[3][4];
  double sv_np, sv_pv, sv_avp1, sv_rp1, sv_ldv, sv_ldi, sv_rvc0, sv_rivc, sv_prior; int sv_shared, pad_;
};
template <int K>
__device__ inline void stage_l2_pre(const SMath &m, const ChainDev &c, const zz_hyper &h, const MoveDraws *md, PreL2 &P) {
  static_assert(K <= 3, "the outcome tables of mhActiveMeansDif hold three rounds");
  {
    const MoveDraws &d = md[MD_IM];
    const double thr = c.pr.threshold_i, im = h.inactive_means;
    P.riv = 1.0 / (2 * h.inactive_variances);
    P.imp = thr - fabs(thr - im - c.t_im * d.n01);
    P.im_pp = dgamma_log_diff(m, thr - P.imp, thr - im, c.pr.inactive_means_prior_shape, c.pr.inactive_means_prior_rate);
  }
  for (int k = 0; k < K; ++k) P.rav[k] = 1.0 / (2 * h.active_variances[k]);
  for (int it = 0; it < K; ++it) {
    const MoveDraws &d = md[MD_AMD + it];
    int k = (int)(d.u[0] * K); if (k >= K) k = K - 1;
    P.amd_k[it] = k;
    for (int mask = 0; mask < (1 << it); ++mask) {
      double dk = c.active_means_dif[k];                                  // the current dif[k]: the proposal of the latest accepted round on k
      for (int j = it - 1; j >= 0; --j) if (((mask >> j) & 1) && P.amd_k[j] == k) { dk = P.amd_difp[j][mask & ((1 << j) - 1)]; break; }
      const double difp = fabs(dk + c.t_am[k] * d.n12);
      P.amd_difp[it][mask] = difp;
      P.amd_pp[it][mask] = dgamma_log_diff(m, difp, dk, c.pr.active_means_dif_prior_shape, c.pr.active_means_dif_prior_rate);
    }
  }
  {
    const MoveDraws &d = md[MD_SV];
    const double lo = c.pr.variances_prior_log_min, hi = c.pr.variances_prior_log_max;
    const double avc = h.active_variances[0], ivc = h.inactive_variances;
    const double LN10 = 2.302585092994045684, lavc = m.log(avc), pv = lavc * (1.0 / LN10), t = c.t_av;   // log10
    double np = (pv - t) + ((pv + t) - (pv - t)) * d.u[0];                // two_boundary_slide_move, MP:29-46
    if (np > hi) np = 2 * hi - np;
    np = lo + fabs(np - lo);
    P.sv_np = np; P.sv_pv = pv;
    P.sv_avp1 = m.exp(np * LN10); P.sv_rp1 = 1.0 / (2 * P.sv_avp1);      // 10^np
    P.sv_rvc0 = 1.0 / (2 * avc);
    bool shared = true;
    for (int k = 0; k < K; ++k) shared = shared && h.active_variances[k] == avc;
    P.sv_shared = shared;
    P.sv_ldv = np * LN10 - lavc;                                          // log(avp1) - log(avc), one pair of logarithms
    P.sv_rivc = ivc == avc ? P.sv_rvc0 : 1.0 / (2 * ivc);
    P.sv_ldi = P.sv_avp1 != ivc ? m.log(P.sv_avp1) - m.log(ivc) : 0.0;
    // Pr:786-787 dunif(log10(.), lo, hi, log = TRUE): np is inside [lo, hi] by construction, pv unless the chain started outside
    P.sv_prior = dunif_log_rel(np, lo, hi) - dunif_log_rel(pv, lo, hi);
  }
}
__device__ inline double stage_im_post(const PreL2 &P, ChainDev &c, const zz_hyper &h, const MoveDraws *md, const double *st, int K, int tune) {
  const int K1 = K + 1;
  const double im = h.inactive_means, imp = P.imp, n0out = st[0] - st[3 * K1];
  const double qp = st[2 * K1] - 2 * imp * st[K1] + n0out * imp * imp, qc = st[2 * K1] - 2 * im * st[K1] + n0out * im * im;
  const double Rr = h.temperature * (-(qp * P.riv - qc * P.riv) + P.im_pp);       // closed_inactive_diff at equal variances
  if (tune) win_push0(c.w_im);
  if (md[MD_IM].l2 < Rr) { if (tune) win_set_last(c.w_im); return imp; }
  return im;
}
template <int K>
__device__ inline void stage_amd_post(const PreL2 &P, ChainDev &c, const zz_hyper &h, const MoveDraws *md, const double *st, int tune, double *am_out /* [K] */) {
  double amp[K], am[K], dif[K], rav[K];                                   // registers: every index below is a compile-time constant
#pragma unroll
  for (int k = 0; k < K; ++k) { am[k] = h.active_means[k]; dif[k] = c.active_means_dif[k]; rav[k] = P.rav[k]; }
  int mask = 0;
#pragma unroll
  for (int it = 0; it < K; ++it) {
    const int k = P.amd_k[it];
    const double difp = P.amd_difp[it][mask], pp = P.amd_pp[it][mask];
#pragma unroll
    for (int j = 0; j < K; ++j) amp[j] = (j == k ? difp : dif[j]) + c.pr.threshold_a[j];
    const double Rr = h.temperature * (closed_active_diff<K>(st, amp, rav, am, rav, 0.0) + pp);
    if (tune) win_push0(c.w_am[k]);
    if (md[MD_AMD + it].l3 < Rr) {
      if (tune) win_set_last(c.w_am[k]);
      mask |= 1 << it;
#pragma unroll
      for (int j = 0; j < K; ++j) { am[j] = amp[j]; if (j == k) dif[j] = difp; }
    }
  }
#pragma unroll
  for (int k = 0; k < K; ++k) { am_out[k] = am[k]; c.active_means_dif[k] = dif[k]; }
}
// returns the accepted (shared) variance, NaN when the move was rejected
template <int K>
__device__ inline double stage_sv_post(const SMath &m, const PreL2 &P, ChainDev &c, const zz_hyper &h, const MoveDraws *md, const double *st, int tune, double im, const double *am_in) {
  constexpr int K1 = K + 1;
  const double LN10 = 2.302585092994045684;
  if (tune) win_push0(c.w_av);
  double rvp[K], rvc[K], am[K];
#pragma unroll
  for (int k = 0; k < K; ++k) { am[k] = am_in[k]; rvp[k] = P.sv_rp1; rvc[k] = (h.active_variances[k] == h.active_variances[0]) ? P.sv_rvc0 : 1.0 / (2 * h.active_variances[k]); }
  double la;
  if (P.sv_shared) la = closed_active_diff<K>(st, am, rvp, am, rvc, P.sv_ldv);
  else {                                                                  // (a state set from outside with unequal variances)
    la = closed_active_diff<K>(st, am, rvp, am, rvc, 0.0);
    for (int k = 1; k <= K; ++k) la += -0.5 * st[k] * (P.sv_np * LN10 - m.log(h.active_variances[k - 1]));
  }
  const double n0out = st[0] - st[3 * K1];
  const double q = st[2 * K1] - 2 * im * st[K1] + n0out * im * im;       // the same quadratic on both sides: the means do not move
  double li = -(q * P.sv_rp1 - q * P.sv_rivc);
  if (P.sv_avp1 != h.inactive_variances) li += -0.5 * n0out * P.sv_ldi;
  const double Rr = h.temperature * (la + li + P.sv_prior);
  if (md[MD_SV].l1 < Rr) { if (tune) win_set_last(c.w_av); return P.sv_avp1; }
  return NAN;                                                             // rejected: the caller keeps the current variances
}

// mhP_x (Pr:193-243), first half: library, proposed alpha, Hastings and prior ratios — before the sweep that accumulates its gene sum.
// Uniforms w0 (library), w1 (scale) of the generation's SUB_PX block 0; the decision uses w2.
__device__ inline void px_propose(ChainDev &c, const zz_hyper &h, uint64_t seed, uint64_t step, int tune, int R) {
  uint32_t o[4];
  philox_block(seed, (uint32_t)c.chain, step, 0, 0xFFFFFF00ull + SUB_PX, o);
  int lib = (int)(u32_open(o[0]) * R); if (lib >= R) lib = R - 1;
  const double HR = c.t_alpha[lib] * (u32_open(o[1]) - 0.5), ap = h.alpha_r[lib] * exp(HR);   // scale_move, MP:3-13 (libm: this runs
  if (tune) win_push0(c.w_alpha[lib]);                                                           // off the critical path, also in k_gen_prepare)
  // prior ratio: the reference sums dgamma over all R libraries for both vectors; the R - 1 unchanged terms are identical
  c.px_pp = dgamma_log_diff(LibM(), ap, h.alpha_r[lib], c.pr.alpha_r_shape, c.pr.alpha_r_rate);
  c.px_HR = HR; c.px_lib = lib; c.px_alpha = ap; c.px_valid = 1;
}
// the same in two pieces for the device-resident loop: the NEXT generation's proposal is prepared while the current sweep runs,
// for both values alpha_r[lib] can have by then (unchanged / the pending proposal on the same library was accepted)
struct PrePx { int lib, same; double HR, ap[2], pp[2]; };
__device__ inline void px_propose_pre(const ChainDev &c, const zz_hyper &h, uint64_t seed, uint64_t step, int R, PrePx &P) {
  uint32_t o[4];
  philox_block(seed, (uint32_t)c.chain, step, 0, 0xFFFFFF00ull + SUB_PX, o);
  int lib = (int)(u32_open(o[0]) * R); if (lib >= R) lib = R - 1;
  P.lib = lib; P.same = c.px_valid && lib == c.px_lib;
  P.HR = c.t_alpha[lib] * (u32_open(o[1]) - 0.5);
  const double e = exp(P.HR);
  for (int v = 0; v < 2; ++v) {
    const double ac = (v && P.same) ? c.px_alpha : h.alpha_r[lib];
    P.ap[v] = ac * e;
    P.pp[v] = dgamma_log_diff(LibM(), P.ap[v], ac, c.pr.alpha_r_shape, c.pr.alpha_r_rate);
  }
}
__device__ inline void px_propose_post(const PrePx &P, ChainDev &c, int accepted_now, int tune) {
  const int v = (accepted_now && P.same) ? 1 : 0;
  if (tune) win_push0(c.w_alpha[P.lib]);
  c.px_pp = P.pp[v]; c.px_HR = P.HR; c.px_lib = P.lib; c.px_alpha = P.ap[v]; c.px_valid = 1;
}
// second half, after the sweep: `delta` = sum over genes of ll(alpha') - ll(alpha) of the one library, `bad` = non-finite terms seen
__device__ inline double px_log_u(const ChainDev &c, uint64_t seed, uint64_t step) {   // log of the move's accept uniform (needs no statistics)
  uint32_t o[4];
  philox_block(seed, (uint32_t)c.chain, step, 0, 0xFFFFFF00ull + SUB_PX, o);
  return log(u32_open(o[2]));
}
__device__ inline int px_decide(ChainDev &c, const zz_hyper &h, double log_u, double delta, double bad, int tune) {
  const double Rr = h.temperature * (delta + c.px_pp) + c.px_HR;
  const int acc = c.px_valid && bad == 0.0 && (log_u < Rr) && Rr != INFINITY;
  if (acc && tune) win_set_last(c.w_alpha[c.px_lib]);
  return acc;
}

// the pending proposal as the sweep reads it
__device__ inline void set_px(FastH &F, const zz_hyper &h, const ChainDev &c) {
  F.px_prev_acc = c.px_prev_acc;
  F.px_on = c.px_valid; F.px_lib = c.px_lib;
  if (c.px_valid) {
    const double ac = h.alpha_r[c.px_lib], ap = c.px_alpha;
    F.px_am256 = -256.0 * (ap * ZZ_KLOG2E);
    F.px_amin = fmin(ac, ap);
    F.t_clamp_px = 1010.0 / (fmax(ac, ap) * ZZ_KLOG2E);     // valid for both: a scale move keeps alpha' within e of alpha, so t_clamp_px * min >= 54
  }
}

__device__ inline void tune_all_scalars(ChainDev &c, double target, int K, int R, int lane) {   // U:183-240, scalar part, one lane per parameter
  if (lane < R) c.t_alpha[lane] = x_tune_s(win_sum100(c.w_alpha[lane]), c.t_alpha[lane], target, 0.01, 2);
  else if (lane < R + K) { const int k = lane - R; c.t_am[k] = x_tune_s(win_sum100(c.w_am[k]), c.t_am[k], target, 0.01, 10); }
  else switch (lane - R - K) {
    case 0: c.t_s0 = x_tune_s(win_sum100(c.w_s0), c.t_s0, target, 0.0001, 5); break;
    case 1: c.t_s1 = x_tune_s(win_sum100(c.w_s1), c.t_s1, target, 0.0001, 5); break;
    case 2: c.t_tau = x_tune_s(win_sum100(c.w_tau), c.t_tau, target, 0.0001, 5); break;
    case 3: c.t_s0tau = x_tune_s(win_sum100(c.w_s0tau), c.t_s0tau, target, 0.0001, 5); break;
    case 4: c.t_im = x_tune_s(win_sum100(c.w_im), c.t_im, target, 0.001, 10); break;
    case 5: c.t_av = x_tune_s(win_sum100(c.w_av), c.t_av, target, 0.01, c.pr.variances_prior_log_max - c.pr.variances_prior_log_min); break;
    default: break;
  }
}

}  // namespace zz
