/*
 * zz_oracle.h — CPU ORACLE for zigzag's per-gene MCMC sweep.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the reference's R arithmetic (ammonthompson/zigzag v1.0.0,
 * citations are R/<file>:<lines> in the reference tree).  It is the checker for the CUDA path:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product (zigzag_b200/) never links, imports or calls anything in oracle/.
 *
 * PARITY PINNING: the reference is pure R with no tests, no golden vectors and no RNG seed, and R
 * is not installed in this image, so FUNCTION-LEVEL PARITY IS UNPINNED by the reference.  What pins
 * this oracle: (1) the run-level P(active) values in docs/tutorial.html:681-691 (statistical,
 * tests/test_oracle_lung.py), (2) an independent NumPy restatement (oracle/oracle_np.py) and an
 * mpmath evaluation of the same closed forms (tests/test_oracle_kat.py), (3) the Random123
 * known-answer vectors for Philox4x32-10.
 *
 * Conventions: Xg is G x R column-major (= R's matrix layout = library-major SoA), -Inf = not
 * detected; all indices 0-based here except alloc_within which keeps R's 1..K; every random draw
 * is an explicit input so each move is a pure function (state, hyper, draws) -> (state', decisions).
 * Sums over libraries / genes accumulate in long double like R's rowSums()/sum() do on x86
 * (R runtime fact), unless compiled with -DZZO_NO_LDOUBLE.
 */
#ifndef ZZ_ORACLE_H
#define ZZ_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZZO_MAX_LIBS 64
#define ZZO_MAX_COMP 8

typedef struct zzo_hyper {
  int32_t num_libraries;         /* R */
  int32_t num_active_components; /* K */
  double s0, s1, tau;
  double alpha_r[ZZO_MAX_LIBS];
  double spike_probability;
  double inactive_means, inactive_variances;
  double active_means[ZZO_MAX_COMP], active_variances[ZZO_MAX_COMP];
  double weight_active, weight_within_active[ZZO_MAX_COMP];
  double beta, temperature, variance_g_upper_bound;
} zzo_hyper;

/* per-gene chain state, struct of arrays, all length G (Xg: G*R) */
typedef struct zzo_state {
  int64_t G;
  const double *Xg;          /* [R][G] */
  const double *gene_lengths;/* unit mean */
  double *Yg, *variance_g;
  int32_t *alloc_active_inactive;   /* 0/1 */
  int32_t *alloc_within_active;     /* 1..K */
  int32_t *inactive_spike_allocation; /* 0/1 */
  const int32_t *no_detect_spike;   /* 0/1: all libraries -Inf (== all_zero_idx membership) */
  const int32_t *pinned_active;     /* 0/1 or NULL: active_gene_set */
  double *XgLikelihood, *variance_g_probability;  /* the reference's caches */
  double *tuningParam_yg, *tuningParam_variance_g;
  uint64_t *Yg_trace, *variance_g_trace; /* 100-wide accept windows, 2 x u64 per gene: [2g]=lo,[2g+1]=hi */
} zzo_state;

int zzo_set_threads(int n);  /* OpenMP build: set / query the thread count (returns the count in effect) */

/* ---- scalar building blocks (one gene) ---- */
void   zzo_get_px(const zzo_hyper *h, double y, double gl, double *px /*R*/);          /* U:150-159 */
double zzo_get_sigmax(double s0, double s1, double y);                                 /* U:144-148 */
double zzo_dnorm_log(double x, double mu, double sigma);                               /* nmath dnorm */
double zzo_xg_lik(const zzo_hyper *h, const double *x, int64_t stride, double y, double v,
                  const double *px, int spike_alloc, int nd_spike);                    /* P:57-127 */
double zzo_varg_prior(const zzo_hyper *h, double v, double sx, double gtau);            /* P:11-31 */
double zzo_inactive_lik(double y, double spike_p, double im, double iv, int spike_alloc);/* P:210-238 */
double zzo_active_lik(double y, int k /*1..K*/, const double *am, const double *av);    /* P:240-273 */

/* ---- evaluators over all genes (check 1) ---- */
void zzo_eval_xg_loglik(const zzo_hyper *h, const zzo_state *s, double *out);
void zzo_eval_varg_logprior(const zzo_hyper *h, const zzo_state *s, double *out);
void zzo_eval_level2_loglik(const zzo_hyper *h, const zzo_state *s, double *out);
void zzo_refresh_caches(const zzo_hyper *h, zzo_state *s);   /* I:473,504 */

/* ---- per-gene moves; draws are explicit, indexed by gene; decisions[g] optional ---- */
/* decisions: 0 = reject / not proposed, 1 = accept; for allocation moves the new allocation word */
void zzo_mh_yg(const zzo_hyper *h, zzo_state *s, const double *z, const double *u, int tune,
               uint8_t *decisions, double *log_ratio /*optional, NaN where not proposed*/);      /* Pr:250-333 */
void zzo_mh_variance_g(const zzo_hyper *h, zzo_state *s, const double *u1, const double *u2, int tune,
                       uint8_t *decisions, double *log_ratio);                                   /* Pr:7-56 */
void zzo_gibbs_alloc_active_inactive(const zzo_hyper *h, zzo_state *s, const double *u_ai,
                                     const double *u_within, uint8_t *decisions,
                                     double *prob_active /*optional*/);                          /* Pr:335-361 */
void zzo_gibbs_alloc_within_active(const zzo_hyper *h, zzo_state *s, const double *u_within,
                                   uint8_t *decisions);                                          /* Pr:363-410 */
void zzo_mh_spike_alloc(const zzo_hyper *h, zzo_state *s, const double *z_y, const double *z_v,
                        const double *u, uint8_t *decisions, double *log_ratio);                 /* Pr:528-638 */
/* fused systematic-scan sweep = mhYg -> mhVariance_g -> gibbsAllocationActiveInactive(+within) -> mhSpikeAllocation,
   draws laid out [9][G]: z_yg,u_yg,u1_vg,u2_vg,u_ai,u_within,zy_spike,zv_spike,u_spike */
void zzo_sweep(const zzo_hyper *h, zzo_state *s, const double *draws, int tune, uint8_t *decisions /*[4][G] optional*/);

/* ---- reductions that feed the hyper moves (a14) ---- */
#define ZZO_NSTAT_FIXED 10
/* out layout: [0..K] n_c (c=0 inactive incl. spike, 1..K active comp), [K+1..2K+1] sum y over c (inactive: out-of-spike only),
   [2K+2..3K+2] sum y^2, then n_inspike, n_out, sum z, sum z^2, sum y(out), sum y^2(out), sum y z, lnl(sum XgLikelihood cache),
   sum vgprob cache(out), nan_count ; z = log(variance_g) over out-of-spike genes.  length 3(K+1)+10 */
int  zzo_suffstats_len(int K);
void zzo_suffstats(const zzo_hyper *h, const zzo_state *s, double *out);
/* sum over out-of-spike genes of computeVarianceGPriorProbability at (s0,s1,tau) ; Pr:84-87,126-130,167-170 */
double zzo_varg_hyper_sum(const zzo_hyper *h, const zzo_state *s, double s0, double s1, double tau);
/* out[0] = sum computeInactiveLikelihood over inactive genes at (spike_p, im, iv); out[1] = sum computeActiveLikelihood
   over active genes at (am, av) ; Pr:444-447,485-488,653-654,729-730,780-784 */
void zzo_level2_sums(const zzo_hyper *h, const zzo_state *s, double im, double iv, const double *am,
                     const double *av, double *out2);
/* mhP_x: per library r with proposed alpha: out[r] = sum_g [ll_r(alpha_prop[r]) - ll_r(alpha_r[r])] ; Pr:209-224 */
void zzo_px_delta(const zzo_hyper *h, const zzo_state *s, const double *alpha_prop, const uint8_t *lib_mask, double *out);
/* apply an accepted alpha_r[lib] proposal to the XgLikelihood cache exactly as Pr:213-215,239 (R>2) or :219 (R==2) */
void zzo_px_commit(zzo_hyper *h, zzo_state *s, int lib, double alpha_new);
double zzo_lnl(const zzo_hyper *h, const zzo_state *s);                              /* U:70-75 */

/* ---- tuning (burn-in) ---- */
void zzo_window_init(uint64_t *win, int64_t G);                    /* 77 zeros + 23 ones, I:460,466 */
/* constructor-side initial state (I:416-504) with the draw map of zz_init_state; no_detect_spike_out may be NULL */
void zzo_init_state(const zzo_hyper *h, zzo_state *s, uint64_t seed, uint32_t chain, uint64_t step, uint64_t gene0, double rv_m, double rv_s,
                    int32_t *no_detect_spike_out);
double zzo_x_tune(double accept_sum_over_100, double tuning, double target, double lo, double hi);  /* U:161-181 */
void zzo_tune_per_gene(zzo_state *s, double target);               /* U:202-206 */

/* ---- counter-based RNG shared with the device (spec in DESIGN.md) ---- */
void zzo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* one Philox block (4 words) / four 32-bit open-interval uniforms for (seed, chain, step, blk, global gene index) */
void zzo_philox_block(uint64_t seed, uint32_t chain, uint64_t step, uint32_t blk, uint64_t gene, uint32_t out[4]);
void zzo_uniform4(uint64_t seed, uint32_t chain, uint64_t step, uint32_t blk, uint64_t gene, double out[4]);
void zzo_box_muller(double ua, double ub, double out[2]);
/* fills the [9][G] draw table zzo_sweep consumes, for genes gene0..gene0+G-1 */
void zzo_sweep_draws(uint64_t seed, uint32_t chain, uint64_t step, uint64_t gene0, int64_t G, double *draws);

/* ---- schedule-faithful chain (host restatement of burnin()/mcmc(); hyper moves included) ---- */
typedef struct zzo_priors {
  double weight_active_shape_1, weight_active_shape_2;
  double spike_prior_shape_1, spike_prior_shape_2;
  double active_means_dif_prior_shape, active_means_dif_prior_rate;
  double inactive_means_prior_shape, inactive_means_prior_rate;
  double variances_prior_log_min, variances_prior_log_max; /* shared_variance = TRUE */
  double tau_shape, tau_rate, s0_mu, s0_sigma, s1_shape, s1_rate, alpha_r_shape, alpha_r_rate;
  double threshold_i, threshold_a[ZZO_MAX_COMP];
} zzo_priors;

typedef struct zzo_chain zzo_chain;
zzo_chain *zzo_chain_new(int64_t G, int R, int K, const double *Xg, const double *gene_lengths,
                         const zzo_priors *pr, uint64_t seed);
void zzo_chain_free(zzo_chain *c);
/* reference schedule: each generation = 15 moves drawn with replacement (B:64-67 / M:169-173) */
void zzo_chain_burnin(zzo_chain *c, int ngen, double target);
/* runs ngen generations, accumulating allocation_trace every sample_frequency (M:188-199) */
void zzo_chain_mcmc(zzo_chain *c, int ngen, int sample_frequency);
/* fused-sweep schedule: each generation = zzo_sweep + one pass of every hyper move */
void zzo_chain_run_fused(zzo_chain *c, int ngen, int sample_frequency, int tune, double target);
void zzo_chain_prob_active(const zzo_chain *c, double *out /*G*/);
const zzo_hyper *zzo_chain_hyper(const zzo_chain *c);
/* the scalar chain state in the layout of zz_chain_scalars (include/zigzag_gpu.h): lets a test start the device-resident
   generation loop (zz_run_generations) from exactly this chain and compare the two afterwards */
typedef struct zzo_chain_scalars {
  double active_means_dif[ZZO_MAX_COMP];
  double t_im, t_av, t_am[ZZO_MAX_COMP], t_s0, t_s1, t_tau, t_s0tau, t_alpha[ZZO_MAX_LIBS];
  uint8_t w_im[100], w_av[100], w_am[ZZO_MAX_COMP][100], w_s0[100], w_s1[100], w_tau[100], w_s0tau[100], w_alpha[ZZO_MAX_LIBS][100];
  uint64_t hyper_ctr;
  double hyper_buf[4];
  uint64_t step;
  int64_t gen;
  int64_t n_samples;
} zzo_chain_scalars;
void zzo_chain_export_scalars(const zzo_chain *c, zzo_chain_scalars *out);
const uint32_t *zzo_chain_alloc_trace(const zzo_chain *c);   /* [K+1][G] */
/* from now on the chain draws with Philox key `seed` and chain field `chain_index` (batched chains of one GPU handle share the key) */
void zzo_chain_set_stream(zzo_chain *c, uint64_t seed, uint32_t chain_index);
zzo_state *zzo_chain_state(zzo_chain *c);
double zzo_chain_lnl(const zzo_chain *c);

/* posterior-predictive discrepancies of one simulated data set (R/zigzagPostPredictiveSimulation.R:3-156);
   out = { W_L, W_U, B_rumsfeld_mean, W_L_r[R], B_rumsfeld_r[R] } */
void zzo_post_predictive(const zzo_hyper *h, const zzo_state *s, uint64_t seed, uint64_t step, const double *bins, int nbins, double *out);

#ifdef __cplusplus
}
#endif
#endif
