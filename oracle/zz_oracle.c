/*
 * zz_oracle.c — CPU ORACLE (test infrastructure, never linked into the product).
 * See zz_oracle.h for the scope / pinning statement.  Abbreviations for citations into the reference:
 *   P  = R/zigzagProbability.R     Pr = R/zigzagProposals.R   MP = R/zigzagMoveProposals.R
 *   U  = R/zigzagUtilities.R       I  = R/zigzagInitialize.R  B  = R/zigzagBurnin.R   M = R/zigzagMCMC.R
 * The R code is vectorised over genes; every function below is the per-gene (elementwise) reading of the
 * same statements, in the same operation order.
 */
#include "zz_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifdef ZZO_NO_LDOUBLE
typedef double acc_t;
#else
typedef long double acc_t; /* R's rowSums()/sum() accumulate in LDOUBLE (R runtime fact, not in repo) */
#endif

#define M_LN_SQRT_2PI_ 0.918938533204672741780329736406 /* nmath/dnorm.c constant */
static const double PI_ = 3.141592653589793238462643383280;
static inline double sqrt2pi_(void) { return sqrt(2 * PI_); } /* I:126 */

/* threads used by the OpenMP build (torchrun exports OMP_NUM_THREADS=1, the CPU-baseline legs of bench.py override it) */
int zzo_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
#else
  (void)n; return 1;
#endif
}

/* ------------------------------------------------------------------ building blocks */

/* U:144-148  get_sigmax: exp(ss0 + ss1 * yy) */
double zzo_get_sigmax(double s0, double s1, double y) { return exp(s0 + s1 * y); }

/* U:150-159  get_px: 1 - exp(-outer((gl * exp(yy)), falpha_r)) */
void zzo_get_px(const zzo_hyper *h, double y, double gl, double *px) {
  double t = gl * exp(y);
  for (int r = 0; r < h->num_libraries; ++r) px[r] = 1 - exp(-(t * h->alpha_r[r]));
}

/* stats::dnorm(x, mu, sigma, log=TRUE) as in nmath/dnorm.c (R runtime, not in repo) */
double zzo_dnorm_log(double x, double mu, double sigma) {
  if (isnan(x) || isnan(mu) || isnan(sigma)) return x + mu + sigma;
  if (sigma < 0) return NAN;
  if (!isfinite(sigma)) return -INFINITY;
  if (!isfinite(x) && mu == x) return NAN;
  if (sigma == 0) return (x == mu) ? INFINITY : -INFINITY;
  double z = (x - mu) / sigma;
  if (!isfinite(z)) return -INFINITY;
  z = fabs(z);
  if (z >= 2 * sqrt(DBL_MAX)) return -INFINITY;
  return -(M_LN_SQRT_2PI_ + 0.5 * z * z + log(sigma));
}

/* P:57-127 computeXgLikelihood, one gene.  x is read with a stride (library-major Xg). */
double zzo_xg_lik(const zzo_hyper *h, const double *x, int64_t stride, double y, double v,
                  const double *px, int spike_alloc, int nd_spike) {
  double lnl;
  if (spike_alloc == 1) {
    lnl = log((double)(spike_alloc * nd_spike)); /* P:117 */
  } else {
    acc_t acc = 0; /* rowSums, P:73 */
    double sd = sqrt(v); /* P:85: sqrt(sx[...]) — "sx" is variance_g */
    for (int r = 0; r < h->num_libraries; ++r) {
      double xr = x[(int64_t)r * stride];
      double ll;
      if (xr == -INFINITY) ll = log(1 - px[r]);                 /* P:82 */
      else ll = log(px[r]) + zzo_dnorm_log(xr, y, sd);          /* P:84-85 */
      acc += ll;
    }
    lnl = (double)acc;
  }
  if (h->beta < 0.0000001) lnl = 0; /* P:121-123 */
  return h->beta * lnl;             /* P:125 */
}

/* log of the lognormal CDF, stats::plnorm(q, meanlog, sdlog, log.p=TRUE) = pnorm(log q,...,log.p=TRUE).
   R uses Cody's rational approximations (nmath/pnorm.c, not in repo); restated through erfc. */
static double plnorm_log(double q, double meanlog, double sdlog) {
  if (q <= 0) return -INFINITY;
  double zz = (log(q) - meanlog) / sdlog;
  double p = 0.5 * erfc(-zz / sqrt(2.0));
  return log(p);
}

/* P:11-31 computeVarianceGPriorProbability(x = variance_g, sx = Sg, gtau) */
double zzo_varg_prior(const zzo_hyper *h, double x, double sx, double gtau) {
  double rtgtau = sqrt(gtau);
  double q = (log(x) - log(sx) - gtau) / rtgtau;
  double lnp = -log(x * rtgtau * sqrt2pi_()) - 0.5 * (q * q); /* R's ^2 is x*x */
  if (h->variance_g_upper_bound != INFINITY)
    lnp = lnp - plnorm_log(h->variance_g_upper_bound, log(sx) + gtau, sqrt(gtau));
  return lnp;
}

/* P:210-238 computeInactiveLikelihood */
double zzo_inactive_lik(double y, double spike_p, double im, double iv, int spike_alloc) {
  double is = sqrt(iv);
  if (spike_alloc == 1) return log(spike_p);
  double d = y - im;
  return log(1 - spike_p) - log(sqrt2pi_() * is) - (d * d) / (2 * iv);
}

/* P:240-273 computeActiveLikelihood, k in 1..K */
double zzo_active_lik(double y, int k, const double *am, const double *av) {
  double d = y - am[k - 1];
  return -log(sqrt2pi_() * sqrt(av[k - 1])) - (d * d) / (2 * av[k - 1]);
}

static inline double level2_lik(const zzo_hyper *h, double y, int ai, int k, int spike_alloc) {
  if (ai == 0) return zzo_inactive_lik(y, h->spike_probability, h->inactive_means, h->inactive_variances, spike_alloc);
  return zzo_active_lik(y, k, h->active_means, h->active_variances);
}

/* ------------------------------------------------------------------ evaluators */

void zzo_eval_xg_loglik(const zzo_hyper *h, const zzo_state *s, double *out) {
#pragma omp parallel for schedule(static)
  for (int64_t g = 0; g < s->G; ++g) {
    double px[ZZO_MAX_LIBS];
    zzo_get_px(h, s->Yg[g], s->gene_lengths[g], px);
    out[g] = zzo_xg_lik(h, s->Xg + g, s->G, s->Yg[g], s->variance_g[g], px,
                        s->inactive_spike_allocation[g], s->no_detect_spike[g]);
  }
}

void zzo_eval_varg_logprior(const zzo_hyper *h, const zzo_state *s, double *out) {
#pragma omp parallel for schedule(static)
  for (int64_t g = 0; g < s->G; ++g)
    out[g] = zzo_varg_prior(h, s->variance_g[g], zzo_get_sigmax(h->s0, h->s1, s->Yg[g]), h->tau);
}

void zzo_eval_level2_loglik(const zzo_hyper *h, const zzo_state *s, double *out) {
#pragma omp parallel for schedule(static)
  for (int64_t g = 0; g < s->G; ++g)
    out[g] = level2_lik(h, s->Yg[g], s->alloc_active_inactive[g], s->alloc_within_active[g],
                        s->inactive_spike_allocation[g]);
}

/* I:473 and I:504 */
void zzo_refresh_caches(const zzo_hyper *h, zzo_state *s) {
  zzo_eval_xg_loglik(h, s, s->XgLikelihood);
  zzo_eval_varg_logprior(h, s, s->variance_g_probability);
}

/* ------------------------------------------------------------------ accept windows */

static inline void window_push(uint64_t *w, int bit) { /* cbind(trace[,-1], bit): Pr:316-320, 44-48 */
  uint64_t lo = w[0], hi = w[1];
  hi = ((hi << 1) | (lo >> 63)) & ((1ULL << 36) - 1);
  lo = (lo << 1) | (uint64_t)(bit & 1);
  w[0] = lo; w[1] = hi;
}
static inline int window_sum(const uint64_t *w) {
  return __builtin_popcountll(w[0]) + __builtin_popcountll(w[1] & ((1ULL << 36) - 1));
}
void zzo_window_init(uint64_t *win, int64_t G) { /* c(rep(0,77), rep(1,23)) */
  for (int64_t g = 0; g < G; ++g) { win[2 * g] = (1ULL << 23) - 1; win[2 * g + 1] = 0; }
}

/* ------------------------------------------------------------------ per-gene moves */

static inline int mh_yg_one(const zzo_hyper *h, zzo_state *s, int64_t g, double z, double u, int tune, double *lr) {
  if (s->inactive_spike_allocation[g] != 0) { if (lr) *lr = NAN; return 0; } /* out_spike_idx only, Pr:252-257 */
  double y = s->Yg[g], v = s->variance_g[g];
  int ai = s->alloc_active_inactive[g], k = s->alloc_within_active[g];
  double yp = y + s->tuningParam_yg[g] * z;                       /* Pr:257 rnorm(mean, sd) = mu + sd*z */
  double Sgp = zzo_get_sigmax(h->s0, h->s1, yp);                   /* Pr:259 */
  double px[ZZO_MAX_LIBS];
  zzo_get_px(h, yp, s->gene_lengths[g], px);                       /* Pr:261 */
  double LXp = zzo_xg_lik(h, s->Xg + g, s->G, yp, v, px, 0, s->no_detect_spike[g]); /* Pr:264 */
  double PVp = zzo_varg_prior(h, v, Sgp, h->tau);                  /* Pr:267 */
  double Lp = LXp;
  Lp = Lp + level2_lik(h, yp, ai, k, 0);                           /* Pr:271-277 */
  Lp = Lp + PVp;                                                   /* Pr:279 */
  double Lc = s->XgLikelihood[g];
  Lc = Lc + level2_lik(h, y, ai, k, 0);                            /* Pr:285-291 */
  Lc = Lc + s->variance_g_probability[g];                          /* Pr:293 */
  double Rr = h->temperature * (Lp - Lc);                          /* Pr:296 */
  int acc = (log(u) < Rr);                                         /* Pr:302 ; NaN => reject (documented deviation) */
  if (lr) *lr = Rr;
  if (acc) { s->Yg[g] = yp; s->XgLikelihood[g] = LXp; s->variance_g_probability[g] = PVp; } /* Pr:310,325,331 */
  if (tune) window_push(s->Yg_trace + 2 * g, acc);                 /* Pr:312-322 */
  return acc;
}

void zzo_mh_yg(const zzo_hyper *h, zzo_state *s, const double *z, const double *u, int tune,
               uint8_t *decisions, double *log_ratio) {
#pragma omp parallel for schedule(static)
  for (int64_t g = 0; g < s->G; ++g) {
    int a = mh_yg_one(h, s, g, z[g], u[g], tune, log_ratio ? log_ratio + g : NULL);
    if (decisions) decisions[g] = (uint8_t)a;
  }
}

static inline int mh_vg_one(const zzo_hyper *h, zzo_state *s, int64_t g, double u1, double u2, int tune, double *lr) {
  if (s->inactive_spike_allocation[g] != 0) { if (lr) *lr = NAN; return 0; }
  double y = s->Yg[g], v = s->variance_g[g];
  double sf = exp(s->tuningParam_variance_g[g] * (u1 - 0.5));      /* MP:5 */
  double vp = v * sf;                                              /* MP:7 */
  double HR = log(sf);                                             /* Pr:15 */
  double px[ZZO_MAX_LIBS];
  zzo_get_px(h, y, s->gene_lengths[g], px);                        /* field p_x == get_px(current state), U:133-142 */
  double LXp = zzo_xg_lik(h, s->Xg + g, s->G, y, vp, px, 0, s->no_detect_spike[g]); /* Pr:17 */
  double Sg = zzo_get_sigmax(h->s0, h->s1, y);                     /* field Sg, U:135 */
  double PVp = zzo_varg_prior(h, vp, Sg, h->tau);                  /* Pr:19 */
  double Lp = LXp + PVp;                                           /* Pr:21 */
  double Lc = s->XgLikelihood[g] + s->variance_g_probability[g];   /* Pr:22 */
  double Rr = h->temperature * (Lp - Lc) + HR;                     /* Pr:24 */
  if (vp > h->variance_g_upper_bound) Rr = -INFINITY;              /* Pr:26 */
  int acc = (log(u2) < Rr);                                        /* Pr:35 */
  if (lr) *lr = Rr;
  if (acc) { s->variance_g[g] = vp; s->XgLikelihood[g] = LXp; s->variance_g_probability[g] = PVp; } /* Pr:38,52,54 */
  if (tune) window_push(s->variance_g_trace + 2 * g, acc);         /* Pr:40-50 */
  return acc;
}

void zzo_mh_variance_g(const zzo_hyper *h, zzo_state *s, const double *u1, const double *u2, int tune,
                       uint8_t *decisions, double *log_ratio) {
#pragma omp parallel for schedule(static)
  for (int64_t g = 0; g < s->G; ++g) {
    int a = mh_vg_one(h, s, g, u1[g], u2[g], tune, log_ratio ? log_ratio + g : NULL);
    if (decisions) decisions[g] = (uint8_t)a;
  }
}

/* Pr:363-410, one active gene: returns new k (1..K) or 0 when the row is NaN (keep old) */
static inline int within_active_one(const zzo_hyper *h, double y, double u) {
  int K = h->num_active_components;
  double Lk[ZZO_MAX_COMP];
  acc_t rs = 0;
  for (int k = 1; k <= K; ++k) {
    double logLk = log(h->weight_within_active[k - 1]) + zzo_active_lik(y, k, h->active_means, h->active_variances) - 0; /* Pr:375-376 */
    Lk[k - 1] = exp(logLk);                                        /* Pr:395 */
    rs += Lk[k - 1];
  }
  double rsum = (double)rs;                                        /* rowSums, Pr:398 */
  acc_t cum = 0;
  int cnt = 0, nan_row = 0;
  for (int k = 0; k < K; ++k) {
    double prob = Lk[k] / rsum;                                    /* Pr:398 */
    cum += prob;                                                   /* matrixStats::rowCumsums, Pr:399 */
    double rs_k = floor((double)cum / u);                          /* Pr:400 */
    if (isnan(rs_k)) nan_row = 1;
    if (rs_k == 0) cnt++;                                          /* Pr:401-402 */
  }
  if (nan_row) return 0;                                           /* Pr:403: NA rows keep the old allocation */
  int knew = cnt + 1;
  if (knew > K) knew = K; /* rounding can give K+1 in R (SURVEY §7 hard part 5): clamp, documented deviation */
  return knew;
}

static inline int alloc_ai_one(const zzo_hyper *h, zzo_state *s, int64_t g, double u_ai, double u_w, double *pa_out) {
  double y = s->Yg[g];
  int sa = s->inactive_spike_allocation[g];
  int K = h->num_active_components;
  /* Pr:339 with all=TRUE (P:214-220) */
  double L0 = exp(zzo_inactive_lik(y, h->spike_probability, h->inactive_means, h->inactive_variances, sa));
  double L1 = 0;
  for (int k = 1; k <= K; ++k)                                     /* Pr:342-346 */
    L1 = L1 + h->weight_within_active[k - 1] * exp(zzo_active_lik(y, k, h->active_means, h->active_variances));
  double wa = h->weight_active;
  double num = wa * L1 * (1 - sa);
  double pa = num / ((1 - wa) * L0 + wa * L1 * (1 - sa));          /* Pr:348 */
  if (isnan(pa)) pa = wa;                                          /* Pr:349 */
  if (pa_out) *pa_out = pa;
  int a = (u_ai < pa) ? 1 : 0;                                     /* Pr:353 rbinom(1, p): Bernoulli by inversion of u */
  if (s->pinned_active && s->pinned_active[g]) a = 1;              /* Pr:355 */
  s->alloc_active_inactive[g] = a;
  if (a) {                                                         /* Pr:359 -> Pr:363-410 */
    int kn = within_active_one(h, y, u_w);
    if (kn > 0) s->alloc_within_active[g] = kn;
  }
  return a;
}

void zzo_gibbs_alloc_active_inactive(const zzo_hyper *h, zzo_state *s, const double *u_ai,
                                     const double *u_within, uint8_t *decisions, double *prob_active) {
#pragma omp parallel for schedule(static)
  for (int64_t g = 0; g < s->G; ++g) {
    int a = alloc_ai_one(h, s, g, u_ai[g], u_within[g], prob_active ? prob_active + g : NULL);
    if (decisions) decisions[g] = (uint8_t)(a ? s->alloc_within_active[g] : 0);
  }
}

void zzo_gibbs_alloc_within_active(const zzo_hyper *h, zzo_state *s, const double *u_within, uint8_t *decisions) {
#pragma omp parallel for schedule(static)
  for (int64_t g = 0; g < s->G; ++g) {
    if (s->alloc_active_inactive[g] == 1) {
      int kn = within_active_one(h, s->Yg[g], u_within[g]);
      if (kn > 0) s->alloc_within_active[g] = kn;
    }
    if (decisions) decisions[g] = (uint8_t)(s->alloc_active_inactive[g] ? s->alloc_within_active[g] : 0);
  }
}

static inline int spike_one(const zzo_hyper *h, zzo_state *s, int64_t g, double zy, double zv, double u, double *lr) {
  /* Pr:533: all_zero_idx[allocation_active_inactive[all_zero_idx] == 0] */
  if (!(s->no_detect_spike[g] == 1 && s->alloc_active_inactive[g] == 0)) { if (lr) *lr = NAN; return 0; }
  int cur = s->inactive_spike_allocation[g];
  int prop = abs(cur - 1);                                         /* Pr:535 */
  double y = s->Yg[g], v = s->variance_g[g];
  double sp = h->spike_probability, im = h->inactive_means, iv = h->inactive_variances;
  double yp = im + sqrt(iv) * zy;                                  /* Pr:545 */
  double Sgp = zzo_get_sigmax(h->s0, h->s1, yp);                   /* Pr:547 */
  double vp = exp((log(Sgp) + h->tau) + sqrt(h->tau) * zv);        /* Pr:549 rlnorm = exp(rnorm(meanlog, sdlog)) */
  double px[ZZO_MAX_LIBS];
  zzo_get_px(h, yp, s->gene_lengths[g], px);                       /* Pr:551 */
  double PVp = zzo_varg_prior(h, vp, Sgp, h->tau);                 /* Pr:556 */
  double PVc = zzo_varg_prior(h, v, zzo_get_sigmax(h->s0, h->s1, y), h->tau); /* Pr:558 (field Sg) */
  double pygp = zzo_inactive_lik(yp, sp, im, iv, prop);            /* Pr:560 */
  double pyg = zzo_inactive_lik(y, sp, im, iv, cur);               /* Pr:563 */
  double pp = (prop == 0) ? (pygp + PVp) : log(sp);                /* Pr:569-570 */
  double pc = (cur == 1) ? log(sp) : (pyg + PVc);                  /* Pr:572-573 */
  double Lp = zzo_xg_lik(h, s->Xg + g, s->G, yp, vp, px, prop, s->no_detect_spike[g]); /* Pr:579 */
  double Lc = s->XgLikelihood[g];                                  /* Pr:581 */
  if (Lp == -INFINITY && Lc == -INFINITY) { Lp = 0; Lc = 0; }      /* Pr:583-585 */
  double HR = (prop == 0) ? -(pygp - log(1 - sp) + PVp)            /* Pr:592-593 */
                          : (pyg - log(1 - sp) + PVc);             /* Pr:594-595 */
  double Rr = h->temperature * (Lp - Lc + pp - pc) + HR;           /* Pr:601 */
  if (vp > h->variance_g_upper_bound) Rr = -INFINITY;              /* Pr:603 */
  int acc = (log(u) < Rr);                                         /* Pr:605 */
  if (lr) *lr = Rr;
  s->inactive_spike_allocation[g] = acc ? prop : cur;              /* Pr:616 */
  if (prop == 0) {                                                 /* [proposed_out_spike_idx] only */
    s->Yg[g] = acc ? yp : y;                                       /* Pr:620 */
    s->variance_g[g] = acc ? vp : v;                               /* Pr:624 */
    s->variance_g_probability[g] = acc ? PVp : PVc;                /* Pr:636 */
  }
  s->XgLikelihood[g] = acc ? Lp : Lc;                              /* Pr:630-632 (after the -Inf pair zeroing) */
  return acc;
}

void zzo_mh_spike_alloc(const zzo_hyper *h, zzo_state *s, const double *z_y, const double *z_v,
                        const double *u, uint8_t *decisions, double *log_ratio) {
#pragma omp parallel for schedule(static)
  for (int64_t g = 0; g < s->G; ++g) {
    int a = spike_one(h, s, g, z_y[g], z_v[g], u[g], log_ratio ? log_ratio + g : NULL);
    if (decisions) decisions[g] = (uint8_t)a;
  }
}

void zzo_sweep(const zzo_hyper *h, zzo_state *s, const double *d, int tune, uint8_t *dec) {
  int64_t G = s->G;
#pragma omp parallel for schedule(static)
  for (int64_t g = 0; g < G; ++g) {
    int a0 = mh_yg_one(h, s, g, d[0 * G + g], d[1 * G + g], tune, NULL);
    int a1 = mh_vg_one(h, s, g, d[2 * G + g], d[3 * G + g], tune, NULL);
    int a2 = alloc_ai_one(h, s, g, d[4 * G + g], d[5 * G + g], NULL);
    int a3 = spike_one(h, s, g, d[6 * G + g], d[7 * G + g], d[8 * G + g], NULL);
    if (dec) {
      dec[0 * G + g] = (uint8_t)a0; dec[1 * G + g] = (uint8_t)a1;
      dec[2 * G + g] = (uint8_t)(a2 ? s->alloc_within_active[g] : 0); dec[3 * G + g] = (uint8_t)a3;
    }
  }
}

/* ------------------------------------------------------------------ reductions (a14) */

int zzo_suffstats_len(int K) { return 3 * (K + 1) + ZZO_NSTAT_FIXED; }

void zzo_suffstats(const zzo_hyper *h, const zzo_state *s, double *out) {
  int K = h->num_active_components, K1 = K + 1, n = zzo_suffstats_len(K);
  acc_t acc[3 * (ZZO_MAX_COMP + 1) + ZZO_NSTAT_FIXED];
  for (int i = 0; i < n; ++i) acc[i] = 0;
  for (int64_t g = 0; g < s->G; ++g) {
    double y = s->Yg[g], v = s->variance_g[g];
    int ai = s->alloc_active_inactive[g], sa = s->inactive_spike_allocation[g];
    int c = ai ? s->alloc_within_active[g] : 0;                   /* Pr:415 all_allocations */
    acc[c] += 1;
    if (ai || !sa) { acc[K1 + c] += y; acc[2 * K1 + c] += y * y; }
    acc[3 * K1 + 0] += (sa == 1);
    if (!sa) {
      double z = log(v);
      acc[3 * K1 + 1] += 1; acc[3 * K1 + 2] += z; acc[3 * K1 + 3] += z * z;
      acc[3 * K1 + 4] += y; acc[3 * K1 + 5] += y * y; acc[3 * K1 + 6] += y * z;
      acc[3 * K1 + 8] += s->variance_g_probability[g];
    }
    acc[3 * K1 + 7] += s->XgLikelihood[g];
    if (!isfinite(y) || !isfinite(v)) acc[3 * K1 + 9] += 1;
  }
  for (int i = 0; i < n; ++i) out[i] = (double)acc[i];
}

double zzo_varg_hyper_sum(const zzo_hyper *h, const zzo_state *s, double s0, double s1, double tau) {
  acc_t acc = 0;
  for (int64_t g = 0; g < s->G; ++g)
    if (s->inactive_spike_allocation[g] == 0)
      acc += zzo_varg_prior(h, s->variance_g[g], zzo_get_sigmax(s0, s1, s->Yg[g]), tau);
  return (double)acc;
}

void zzo_level2_sums(const zzo_hyper *h, const zzo_state *s, double im, double iv, const double *am,
                     const double *av, double *out2) {
  acc_t a0 = 0, a1 = 0;
  for (int64_t g = 0; g < s->G; ++g) {
    if (s->alloc_active_inactive[g] == 0)
      a0 += zzo_inactive_lik(s->Yg[g], h->spike_probability, im, iv, s->inactive_spike_allocation[g]);
    else
      a1 += zzo_active_lik(s->Yg[g], s->alloc_within_active[g], am, av);
  }
  out2[0] = (double)a0; out2[1] = (double)a1;
}

/* single-library computeXgLikelihood(Xg[,lib], Yg, variance_g, px[,lib]) as used by Pr:213-215 */
static inline double xg_lik_onelib(const zzo_hyper *h, const zzo_state *s, int64_t g, int lib, double alpha) {
  if (s->inactive_spike_allocation[g] == 1)
    return h->beta * ((h->beta < 0.0000001) ? 0.0 : log((double)s->no_detect_spike[g]));
  double y = s->Yg[g];
  double p = 1 - exp(-((s->gene_lengths[g] * exp(y)) * alpha));
  double xr = s->Xg[(int64_t)lib * s->G + g];
  double ll = (xr == -INFINITY) ? log(1 - p) : log(p) + zzo_dnorm_log(xr, y, sqrt(s->variance_g[g]));
  if (h->beta < 0.0000001) ll = 0;
  return h->beta * ll;
}

void zzo_px_delta(const zzo_hyper *h, const zzo_state *s, const double *alpha_prop, const uint8_t *lib_mask, double *out) {
  for (int r = 0; r < h->num_libraries; ++r) {
    out[r] = 0;
    if (lib_mask && !lib_mask[r]) continue;
    acc_t acc = 0;
    for (int64_t g = 0; g < s->G; ++g)
      acc += xg_lik_onelib(h, s, g, r, alpha_prop[r]) - xg_lik_onelib(h, s, g, r, h->alpha_r[r]);
    out[r] = (double)acc;
  }
}

void zzo_px_commit(zzo_hyper *h, zzo_state *s, int lib, double alpha_new) {
  if (h->num_libraries > 2) {                                      /* Pr:213-215 */
    for (int64_t g = 0; g < s->G; ++g)
      s->XgLikelihood[g] = s->XgLikelihood[g] + xg_lik_onelib(h, s, g, lib, alpha_new) - xg_lik_onelib(h, s, g, lib, h->alpha_r[lib]);
    h->alpha_r[lib] = alpha_new;
  } else {                                                         /* Pr:219 */
    h->alpha_r[lib] = alpha_new;
    zzo_eval_xg_loglik(h, s, s->XgLikelihood);
  }
}

/* U:70-75 calculate_lnl (num_libs > 1): sum(computeXgLikelihood(Xg, Yg, variance_g, p_x)) */
double zzo_lnl(const zzo_hyper *h, const zzo_state *s) {
  acc_t acc = 0;
  for (int64_t g = 0; g < s->G; ++g) {
    double px[ZZO_MAX_LIBS];
    zzo_get_px(h, s->Yg[g], s->gene_lengths[g], px);
    acc += zzo_xg_lik(h, s->Xg + g, s->G, s->Yg[g], s->variance_g[g], px,
                      s->inactive_spike_allocation[g], s->no_detect_spike[g]);
  }
  return (double)acc;
}

/* ------------------------------------------------------------------ tuning */

/* U:161-181 x_tune */
double zzo_x_tune(double wsum, double tuning, double target, double lo, double hi) {
  double pjump = wsum / 100;
  double nt = tuning * (tan(PI_ * pjump / 2) / tan(PI_ * target / 2));
  if (nt < lo) nt = lo;
  if (nt > hi) nt = hi;
  return nt;
}

/* U:202-206 */
void zzo_tune_per_gene(zzo_state *s, double target) {
  for (int64_t g = 0; g < s->G; ++g) {
    s->tuningParam_variance_g[g] = zzo_x_tune(window_sum(s->variance_g_trace + 2 * g), s->tuningParam_variance_g[g], target, 0.0001, 10);
    s->tuningParam_yg[g] = zzo_x_tune(window_sum(s->Yg_trace + 2 * g), s->tuningParam_yg[g], target, 0.0001, 10);
  }
}

/* ------------------------------------------------------------------ counter-based RNG (spec: DESIGN.md §RNG) */

static inline void mulhilo(uint32_t a, uint32_t b, uint32_t *hi, uint32_t *lo) {
  uint64_t p = (uint64_t)a * b; *hi = (uint32_t)(p >> 32); *lo = (uint32_t)p;
}

/* Philox4x32-10 (Salmon et al. 2011, Random123).  KAT vectors in tests/test_philox.py */
void zzo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint32_t h0, l0, h1, l1;
    mulhilo(0xD2511F53u, c0, &h0, &l0);
    mulhilo(0xCD9E8D57u, c2, &h1, &l1);
    uint32_t n0 = h1 ^ c1 ^ k0, n1 = l1, n2 = h0 ^ c3 ^ k1, n3 = l0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline double u32_open(uint32_t w) { /* (w + 0.5) * 2^-32: open interval, 32-bit resolution like R's unif_rand */
  return ((double)w + 0.5) * (1.0 / 4294967296.0);
}

/* One Philox block = 4 words for (seed, chain, step, blk, global gene index).
   counter = (gene [< 2^32], blk | chain<<8, step lo, step hi); key = (seed lo, seed hi).
   Draw map of a sweep (DESIGN.md "RNG"):
     blk 0: w0,w1 -> Box-Muller -> z of mhYg (cos branch) ; w2 -> u of mhYg ; w3 -> u_proposal of mhVariance_g
     blk 1: w0 -> u_accept of mhVariance_g ; w1 -> u of the active/inactive Gibbs draw ; w2 -> u of the within-active draw
     blk 2: w0,w1 -> Box-Muller -> (z_y, z_v) of mhSpikeAllocation ; w2 -> its u                                   */
void zzo_philox_block(uint64_t seed, uint32_t chain, uint64_t step, uint32_t blk, uint64_t gene, uint32_t out[4]) {
  uint32_t ctr[4] = { (uint32_t)gene, (blk & 255u) | (chain << 8), (uint32_t)step, (uint32_t)(step >> 32) };
  uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
  zzo_philox4x32_10(ctr, key, out);
}

void zzo_uniform4(uint64_t seed, uint32_t chain, uint64_t step, uint32_t blk, uint64_t gene, double out[4]) {
  uint32_t o[4];
  zzo_philox_block(seed, chain, step, blk, gene, o);
  for (int i = 0; i < 4; ++i) out[i] = u32_open(o[i]);
}

void zzo_box_muller(double ua, double ub, double out[2]) {
  double r = sqrt(-2.0 * log(ua));
  out[0] = r * cos(2.0 * PI_ * ub);
  out[1] = r * sin(2.0 * PI_ * ub);
}

void zzo_sweep_draws(uint64_t seed, uint32_t chain, uint64_t step, uint64_t gene0, int64_t G, double *d) {
  for (int64_t g = 0; g < G; ++g) {
    double a[4], b[4], c[4], n[2];
    zzo_uniform4(seed, chain, step, 0, gene0 + g, a);
    zzo_uniform4(seed, chain, step, 1, gene0 + g, b);
    zzo_uniform4(seed, chain, step, 2, gene0 + g, c);
    zzo_box_muller(a[0], a[1], n);
    d[0 * G + g] = n[0]; d[1 * G + g] = a[2];
    d[2 * G + g] = a[3]; d[3 * G + g] = b[0];
    d[4 * G + g] = b[1]; d[5 * G + g] = b[2];
    zzo_box_muller(c[0], c[1], n);
    d[6 * G + g] = n[0]; d[7 * G + g] = n[1]; d[8 * G + g] = c[2];
  }
}

/* Constructor-side initial state, R/zigzagInitialize.R:416-504, with the draw map of include/zigzag_gpu.h:zz_init_state (Philox
   keyed by the global gene index gene0 + g): block 0 = {u of allocation_active_inactive (I:367), u of allocation_within_active
   (I:372), pair (z of rwm for an all-zero gene, z of rwv for a gene with < 2 detected libraries)}, block 1 = pair (z of Yg, z of
   variance_g), block 2 + j = pair of re-draw j.  rv_m, rv_s: log(mean), sd of the row variances of the fully detected genes (I:446).
   The reference re-draws every gene whose likelihood is still -Inf, round by round (I:480-502): gene by gene that is a loop of
   its own per gene.  Tuning parameters 0.5, windows c(rep(0,77), rep(1,23)), both caches refreshed (I:473, I:504). */
void zzo_init_state(const zzo_hyper *h, zzo_state *s, uint64_t seed, uint32_t chain, uint64_t step, uint64_t gene0, double rv_m, double rv_s,
                    int32_t *no_detect_spike_out) {
  const int R = h->num_libraries, K = h->num_active_components;
  const int64_t G = s->G;
  for (int64_t g = 0; g < G; ++g) {
    int n_det = 0; long double sum = 0;
    for (int r = 0; r < R; ++r) { const double x = s->Xg[(int64_t)r * G + g]; if (x != -INFINITY) { ++n_det; sum += x; } }
    const int allzero = n_det == 0;
    if (no_detect_spike_out) no_detect_spike_out[g] = allzero;        /* I:415 */
    s->inactive_spike_allocation[g] = allzero;                        /* I:418 */
    double a[4], b[4], n0[2], n1[2];
    zzo_uniform4(seed, chain, step, 0, gene0 + g, a);
    zzo_box_muller(a[2], a[3], n0);
    const int pinned = s->pinned_active && s->pinned_active[g];
    s->alloc_active_inactive[g] = ((pinned || a[0] < h->weight_active) && !allzero) ? 1 : 0;   /* I:367-369, I:422 */
    int k = 1; double cum = 0;
    for (int j = 0; j < K - 1; ++j) { cum += h->weight_within_active[j]; k += (cum <= a[1]); }   /* I:372 */
    s->alloc_within_active[g] = k;
    zzo_uniform4(seed, chain, step, 1, gene0 + g, b);
    zzo_box_muller(b[0], b[1], n1);
    const double sd_i = sqrt(h->inactive_variances);
    double rwm, rwv;
    if (n_det > 0) rwm = (double)(sum / n_det); else rwm = h->inactive_means + sd_i * n0[0];      /* I:431-438 */
    if (n_det >= 2) {                                                                            /* I:441-448: var() */
      long double ss = 0;
      for (int r = 0; r < R; ++r) { const double x = s->Xg[(int64_t)r * G + g]; if (x != -INFINITY) ss += ((long double)x - rwm) * ((long double)x - rwm); }
      rwv = (double)(ss / (n_det - 1));
    } else rwv = exp(rv_m + rv_s * n0[1]);
    double y = rwm + sqrt(rwv) * n1[0];                               /* I:449 */
    double v = exp(0.5 + 0.2 * n1[1]);                                /* I:465 */
    if (v > h->variance_g_upper_bound) v = h->variance_g_upper_bound; /* I:468 */
    double px[ZZO_MAX_LIBS];
    if (!allzero)
      for (uint32_t j = 0; j < 1000; ++j) {                           /* I:480-502 */
        zzo_get_px(h, y, s->gene_lengths[g], px);
        if (zzo_xg_lik(h, s->Xg + g, G, y, v, px, 0, 0) != -INFINITY) break;
        double c[4], n2[2];
        zzo_uniform4(seed, chain, step, 2 + j, gene0 + g, c);
        zzo_box_muller(c[0], c[1], n2);
        y = h->inactive_means + sd_i * n2[0];
        v = exp(log(zzo_get_sigmax(h->s0, h->s1, y)) + h->tau + sqrt(h->tau) * n2[1]);
        if (v > h->variance_g_upper_bound) v = h->variance_g_upper_bound;
      }
    s->Yg[g] = y; s->variance_g[g] = v; s->tuningParam_yg[g] = 0.5; s->tuningParam_variance_g[g] = 0.5;
  }
  zzo_window_init(s->Yg_trace, G); zzo_window_init(s->variance_g_trace, G);
  zzo_refresh_caches(h, s);
}

/* ------------------------------------------------------------------ schedule-faithful chain */

struct zzo_chain {
  int64_t G; int R, K;
  zzo_hyper h; zzo_priors pr; zzo_state s;
  double *Xg_own, *gl_own;
  int32_t *nd;
  double active_means_dif[ZZO_MAX_COMP];
  /* scalar tuning parameters (I:185-199) and their 100-wide windows */
  double t_im, t_av, t_am[ZZO_MAX_COMP], t_s0, t_s1, t_tau, t_s0tau, t_alpha[ZZO_MAX_LIBS];
  uint8_t w_im[100], w_av[100], w_am[ZZO_MAX_COMP][100], w_s0[100], w_s1[100], w_tau[100], w_s0tau[100], w_alpha[ZZO_MAX_LIBS][100];
  double proposal_probs[15];
  uint64_t seed, hyper_ctr, step;
  uint32_t chain_index; /* chain field of every Philox counter (0 unless zzo_chain_set_stream) */
  double hyper_buf[4];
  int64_t gen;
  uint32_t *alloc_trace; /* [K+1][G] */
  int64_t n_samples;
  double *draws; /* [9][G] draw table of the current step */
  /* hyper-level draws of the fused schedule: one counter-based sub-stream per move (see sub_begin) */
  int sub_on; uint32_t sub_slot, sub_blk; int sub_pos; double sub_buf[4]; uint64_t sub_step;
};

/* Sub-streams of the fused schedule (zzo_chain_run_fused == the device-resident loop zz_run_generations): the scalar draws of
   hyper move `slot` in a generation are the uniforms of the Philox blocks (gene index 0xFFFFFF00 + slot, blk = 0, 1, 2, ...,
   step = the step of that generation's per-gene sweep), consumed in order.  Counter-based, so every move's draws are independent
   of how many uniforms any other move consumed: the device evaluates independent moves side by side (zz_chain.cuh). */
enum { SUB_MIXW = 0x00 /* + component 0..K */, SUB_SP0 = 0x10, SUB_SP1 = 0x11, SUB_IM = 0x20, SUB_AMD = 0x21, SUB_SV = 0x22,
       SUB_TAU = 0x23, SUB_SG = 0x24, SUB_S0TAU = 0x25, SUB_PX = 0x26 };
static void sub_begin(zzo_chain *c, uint32_t slot) { c->sub_slot = slot; c->sub_blk = 0; c->sub_pos = 4; }

/* hyper-level draws: a private Philox stream (gene index 0xFFFFFFFF, blk 255) */
static double c_unif(zzo_chain *c) {
  double o[4];
  if (c->sub_on) {
    if (c->sub_pos == 4) { zzo_uniform4(c->seed, c->chain_index, c->sub_step, c->sub_blk++ & 255u, 0xFFFFFF00ull + c->sub_slot, c->sub_buf); c->sub_pos = 0; }
    return c->sub_buf[c->sub_pos++];
  }
  if ((c->hyper_ctr & 3) == 0) { zzo_uniform4(c->seed, c->chain_index, c->hyper_ctr >> 2, 255, 0xFFFFFFFFull, o); memcpy(c->hyper_buf, o, sizeof o); }
  return c->hyper_buf[c->hyper_ctr++ & 3];
}
static double c_norm(zzo_chain *c) { double n[2]; double a = c_unif(c), b = c_unif(c); zzo_box_muller(a, b, n); return n[0]; }
static double c_gamma(zzo_chain *c, double shape, double rate) { /* Marsaglia & Tsang 2000; distributionally = rgamma */
  if (shape < 1) { double u = c_unif(c); return c_gamma(c, shape + 1, rate) * pow(u, 1.0 / shape); }
  double d = shape - 1.0 / 3, cc = 1 / sqrt(9 * d);
  for (;;) {
    double x = c_norm(c), v = 1 + cc * x;
    if (v <= 0) continue;
    v = v * v * v;
    double u = c_unif(c);
    if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) return d * v / rate;
  }
}
/* the same sampler on a sub-stream of its own: attempt j uses block j of `slot` (w0, w1 -> normal, w2 -> accept), the boost
   uniform of shape < 1 is w3 of block 0 */
static double c_gamma_slot(zzo_chain *c, uint32_t slot, double shape, double rate) {
  double a[4], n[2], boost = 1;
  if (shape < 1) { zzo_uniform4(c->seed, c->chain_index, c->sub_step, 0, 0xFFFFFF00ull + slot, a); boost = pow(a[3], 1.0 / shape); shape += 1; }
  double d = shape - 1.0 / 3, cc = 1 / sqrt(9 * d);
  for (uint32_t j = 0; j < 255; ++j) {
    zzo_uniform4(c->seed, c->chain_index, c->sub_step, j, 0xFFFFFF00ull + slot, a);
    zzo_box_muller(a[0], a[1], n);
    double x = n[0], v = 1 + cc * x;
    if (v <= 0) continue;
    v = v * v * v;
    if (log(a[2]) < 0.5 * x * x + d - d * v + d * log(v)) return d * v / rate * boost;
  }
  return d / rate * boost;
}
static double c_gamma_at(zzo_chain *c, uint32_t slot, double shape, double rate) { return c->sub_on ? c_gamma_slot(c, slot, shape, rate) : c_gamma(c, shape, rate); }
static double c_beta(zzo_chain *c, double a, double b) { double x = c_gamma(c, a, 1), y = c_gamma(c, b, 1); return x / (x + y); }
static int c_sample_w(zzo_chain *c, const double *w, int n) { /* sample(n, 1, prob = w) */
  double tot = 0; for (int i = 0; i < n; ++i) tot += w[i];
  double u = c_unif(c) * tot, cum = 0;
  for (int i = 0; i < n; ++i) { cum += w[i]; if (u < cum) return i; }
  return n - 1;
}
static void win_init(uint8_t *w) { for (int i = 0; i < 100; ++i) w[i] = (i >= 77); }
static void win_push0(uint8_t *w) { memmove(w, w + 1, 99); w[99] = 0; } /* c(trace[-1], 0) */
static int win_sum(const uint8_t *w) { int s = 0; for (int i = 0; i < 100; ++i) s += w[i]; return s; }

static double dgamma_log(double x, double shape, double rate) {
  if (x < 0) return -INFINITY;
  return shape * log(rate) + (shape - 1) * log(x) - rate * x - lgamma(shape);
}
static double dunif_log(double x, double lo, double hi) { return (x >= lo && x <= hi) ? -log(hi - lo) : -INFINITY; }

/* MP:3-13 */
static double scale_move(zzo_chain *c, double param, double t, double *sf) { *sf = exp(t * (c_unif(c) - 0.5)); return param * *sf; }
/* MP:29-46 */
static double two_boundary_slide(zzo_chain *c, double p, double lo, double hi, double t) {
  double np = (p - t) + ((p + t) - (p - t)) * c_unif(c);
  if (np > hi) np = 2 * hi - np;
  return lo + fabs(np - lo);
}

/* draws of one move for every gene, taken from the sweep draw map at the chain's current step; only the Philox blocks the
   move needs are generated (mask bit b = block b) */
static void fill_sweep_draws(zzo_chain *c, int blocks) {
  int64_t G = c->G;
  if (!c->draws) c->draws = (double *)malloc(sizeof(double) * 9 * G);
  double *d = c->draws;
#pragma omp parallel for schedule(static)
  for (int64_t g = 0; g < G; ++g) {
    double a[4], n[2];
    if (blocks & 1) {
      zzo_uniform4(c->seed, c->chain_index, c->step, 0, (uint64_t)g, a);
      zzo_box_muller(a[0], a[1], n);
      d[0 * G + g] = n[0]; d[1 * G + g] = a[2]; d[2 * G + g] = a[3];
    }
    if (blocks & 2) {
      zzo_uniform4(c->seed, c->chain_index, c->step, 1, (uint64_t)g, a);
      d[3 * G + g] = a[0]; d[4 * G + g] = a[1]; d[5 * G + g] = a[2];
    }
    if (blocks & 4) {
      zzo_uniform4(c->seed, c->chain_index, c->step, 2, (uint64_t)g, a);
      zzo_box_muller(a[0], a[1], n);
      d[6 * G + g] = n[0]; d[7 * G + g] = n[1]; d[8 * G + g] = a[2];
    }
  }
}
#define DRAW(c, j) ((c)->draws + (int64_t)(j) * (c)->G)

/* ---- the 15 proposal_list entries (I:207-223), numbered as in the reference (1-based) ---- */
static void mv_gibbsMixtureWeights(zzo_chain *c) { /* Pr:412-431 */
  int K = c->K; double cnt[ZZO_MAX_COMP + 1] = {0}, g_[ZZO_MAX_COMP + 1] = {0}, tot = 0;
  for (int64_t g = 0; g < c->G; ++g) cnt[c->s.alloc_active_inactive[g] ? c->s.alloc_within_active[g] : 0] += 1;
  for (int k = 0; k <= K; ++k) {
    double a = (k == 0 ? c->pr.weight_active_shape_2 : c->pr.weight_active_shape_1 / K) + cnt[k] * c->h.beta;
    g_[k] = c_gamma_at(c, SUB_MIXW + k, a, 1); tot += g_[k];
  }
  double nw0 = g_[0] / tot;
  if (!isnan(nw0)) {
    c->h.weight_active = 1 - nw0;
    double s = 0; for (int k = 1; k <= K; ++k) s += g_[k] / tot;
    for (int k = 1; k <= K; ++k) c->h.weight_within_active[k - 1] = (g_[k] / tot) / s;
  }
}
static void mv_allocAI(zzo_chain *c) {
  fill_sweep_draws(c, 2);
  zzo_gibbs_alloc_active_inactive(&c->h, &c->s, DRAW(c, 4), DRAW(c, 5), NULL, NULL);
}
static void mv_allocWithin(zzo_chain *c) {
  fill_sweep_draws(c, 2);
  zzo_gibbs_alloc_within_active(&c->h, &c->s, DRAW(c, 5), NULL);
}
static void mv_mhInactiveMeans(zzo_chain *c, int tune) { /* Pr:433-471 */
  double thr = c->pr.threshold_i, im = c->h.inactive_means;
  double imp = thr - fabs(thr - im - c->t_im * c_norm(c));
  double pp = dgamma_log(thr - imp, c->pr.inactive_means_prior_shape, c->pr.inactive_means_prior_rate);
  double pc = dgamma_log(thr - im, c->pr.inactive_means_prior_shape, c->pr.inactive_means_prior_rate);
  double Lp[2], Lc[2];
  zzo_level2_sums(&c->h, &c->s, imp, c->h.inactive_variances, c->h.active_means, c->h.active_variances, Lp);
  zzo_level2_sums(&c->h, &c->s, im, c->h.inactive_variances, c->h.active_means, c->h.active_variances, Lc);
  double Rr = c->h.temperature * (Lp[0] - Lc[0] + pp - pc);
  if (tune) win_push0(c->w_im);
  if (log(c_unif(c)) < Rr) { c->h.inactive_means = imp; if (tune) c->w_im[99] = 1; }
}
static void mv_mhActiveMeansDif(zzo_chain *c, int tune) { /* Pr:640-705, multi_thresh_a = TRUE (U:106-118) */
  int K = c->K;
  for (int i = 0; i < K; ++i) {
    int k = (int)(c_unif(c) * K); if (k >= K) k = K - 1;
    double difp = fabs(c->active_means_dif[k] + c->t_am[k] * c_norm(c));
    double mp[ZZO_MAX_COMP];
    for (int j = 0; j < K; ++j) mp[j] = (j == k ? difp : c->active_means_dif[j]) + c->pr.threshold_a[j];
    double Lp[2], Lc[2];
    zzo_level2_sums(&c->h, &c->s, c->h.inactive_means, c->h.inactive_variances, mp, c->h.active_variances, Lp);
    zzo_level2_sums(&c->h, &c->s, c->h.inactive_means, c->h.inactive_variances, c->h.active_means, c->h.active_variances, Lc);
    double pp = dgamma_log(difp, c->pr.active_means_dif_prior_shape, c->pr.active_means_dif_prior_rate);
    double pc = dgamma_log(c->active_means_dif[k], c->pr.active_means_dif_prior_shape, c->pr.active_means_dif_prior_rate);
    double Rr = c->h.temperature * (Lp[1] + pp - Lc[1] - pc);
    if (tune) win_push0(c->w_am[k]);
    if (log(c_unif(c)) < Rr) {
      c->active_means_dif[k] = difp; if (tune) c->w_am[k][99] = 1;
      for (int j = 0; j < K; ++j) c->h.active_means[j] = mp[j];
    }
  }
}
static void mv_mhSharedVariance(zzo_chain *c, int tune) { /* Pr:768-803 */
  int K = c->K; double lo = c->pr.variances_prior_log_min, hi = c->pr.variances_prior_log_max;
  double avp1 = pow(10.0, two_boundary_slide(c, log10(c->h.active_variances[0]), lo, hi, c->t_av));
  double avp[ZZO_MAX_COMP]; for (int k = 0; k < K; ++k) avp[k] = avp1;
  if (tune) win_push0(c->w_av);
  double Lp[2], Lc[2];
  zzo_level2_sums(&c->h, &c->s, c->h.inactive_means, avp1, c->h.active_means, avp, Lp);
  zzo_level2_sums(&c->h, &c->s, c->h.inactive_means, c->h.inactive_variances, c->h.active_means, c->h.active_variances, Lc);
  double app = dunif_log(log10(avp1), lo, hi), apc = dunif_log(log10(c->h.active_variances[0]), lo, hi);
  double Rr = c->h.temperature * (Lp[1] + Lp[0] + app - Lc[1] - Lc[0] - apc);
  if (log(c_unif(c)) < Rr) {
    for (int k = 0; k < K; ++k) c->h.active_variances[k] = avp1;
    c->h.inactive_variances = avp1; if (tune) c->w_av[99] = 1;
  }
}
static void mv_gibbsSpikeProb(zzo_chain *c) { /* Pr:511-526 */
  double nin = 0, ninact = 0;
  for (int64_t g = 0; g < c->G; ++g) { nin += c->s.inactive_spike_allocation[g]; ninact += (c->s.alloc_active_inactive[g] == 0); }
  double gx = c_gamma_at(c, SUB_SP0, nin + c->pr.spike_prior_shape_1, 1), gy = c_gamma_at(c, SUB_SP1, (ninact - nin) + c->pr.spike_prior_shape_2, 1);
  double ns = gx / (gx + gy);                                        /* rbeta as a ratio of gammas, as c_beta */
  if (!isnan(ns)) c->h.spike_probability = ns;
}
static void mv_mhSpikeAllocation(zzo_chain *c) {
  fill_sweep_draws(c, 4);
  zzo_mh_spike_alloc(&c->h, &c->s, DRAW(c, 6), DRAW(c, 7), DRAW(c, 8), NULL, NULL);
}
static void mv_mhYg(zzo_chain *c, int tune) {
  fill_sweep_draws(c, 1);
  zzo_mh_yg(&c->h, &c->s, DRAW(c, 0), DRAW(c, 1), tune, NULL, NULL);
}
static void mv_mhVariance_g(zzo_chain *c, int tune) {
  fill_sweep_draws(c, 3);
  zzo_mh_variance_g(&c->h, &c->s, DRAW(c, 2), DRAW(c, 3), tune, NULL, NULL);
}
static double sum_vgprob_out(const zzo_chain *c) {
  acc_t a = 0; for (int64_t g = 0; g < c->G; ++g) if (!c->s.inactive_spike_allocation[g]) a += c->s.variance_g_probability[g];
  return (double)a;
}
static void commit_vgprob(zzo_chain *c) { /* Pr:113,147,187: overwrite the cache for out-of-spike genes */
  for (int64_t g = 0; g < c->G; ++g) if (!c->s.inactive_spike_allocation[g])
    c->s.variance_g_probability[g] = zzo_varg_prior(&c->h, c->s.variance_g[g], zzo_get_sigmax(c->h.s0, c->h.s1, c->s.Yg[g]), c->h.tau);
}
static double s0_prior(const zzo_chain *c, double x) { return zzo_dnorm_log(x, c->pr.s0_mu, c->pr.s0_sigma); }       /* P:33-39 */
static double s1_prior(const zzo_chain *c, double x) { return dgamma_log(fabs(x), c->pr.s1_shape, c->pr.s1_rate); }    /* P:41-47 */
static double tau_prior(const zzo_chain *c, double x) { return dgamma_log(x, c->pr.tau_shape, c->pr.tau_rate); }       /* P:3-9 */
static void mv_mhTau(zzo_chain *c, int tune) { /* Pr:119-152 */
  double sf, taup = scale_move(c, c->h.tau, c->t_tau, &sf), HR = log(sf);
  double Lp = zzo_varg_hyper_sum(&c->h, &c->s, c->h.s0, c->h.s1, taup), Lc = sum_vgprob_out(c);
  double Rr = c->h.temperature * (Lp - Lc + tau_prior(c, taup) - tau_prior(c, c->h.tau)) + HR;
  if (tune) win_push0(c->w_tau);
  if (log(c_unif(c)) < Rr) { c->h.tau = taup; if (tune) c->w_tau[99] = 1; commit_vgprob(c); }
}
static void mv_mhSg(zzo_chain *c, int tune) { /* Pr:58-117 */
  double HR = 0, s0p = c->h.s0, s1p = c->h.s1; int sel = (c_unif(c) < 0.5) ? 1 : 2;
  if (sel == 1) { s0p = c->h.s0 + c->t_s0 * c_norm(c); if (tune) win_push0(c->w_s0); }
  else { double sf; s1p = scale_move(c, c->h.s1, c->t_s1, &sf); HR = log(sf); if (tune) win_push0(c->w_s1); }
  double Lp = zzo_varg_hyper_sum(&c->h, &c->s, s0p, s1p, c->h.tau), Lc = sum_vgprob_out(c);
  double pp = s0_prior(c, s0p) + s1_prior(c, s1p), pc = s0_prior(c, c->h.s0) + s1_prior(c, c->h.s1);
  double Rr = c->h.temperature * (Lp - Lc + pp - pc) + HR;
  if (log(c_unif(c)) < Rr) {
    c->h.s0 = s0p; c->h.s1 = s1p;
    if (tune) { if (sel == 1) c->w_s0[99] = 1; else c->w_s1[99] = 1; }
    commit_vgprob(c);
  }
}
static void mv_mhS0Tau(zzo_chain *c, int tune) { /* Pr:154-191 */
  double sf, taup = scale_move(c, c->h.tau, c->t_s0tau, &sf), s0p = c->h.s0 - log(sf), HR = log(sf);
  if (tune) win_push0(c->w_s0tau);
  double Lp = zzo_varg_hyper_sum(&c->h, &c->s, s0p, c->h.s1, taup), Lc = sum_vgprob_out(c);
  double pp = s0_prior(c, s0p) + tau_prior(c, taup), pc = s0_prior(c, c->h.s0) + tau_prior(c, c->h.tau);
  double Rr = c->h.temperature * (Lp - Lc + pp - pc) + HR;
  if (log(c_unif(c)) < Rr) { c->h.s0 = s0p; c->h.tau = taup; if (tune) c->w_s0tau[99] = 1; commit_vgprob(c); }
}
static void mv_mhP_x(zzo_chain *c, int tune) { /* Pr:193-243 */
  int R = c->R, lib = (int)(c_unif(c) * R); if (lib >= R) lib = R - 1;
  double sf, ap = scale_move(c, c->h.alpha_r[lib], c->t_alpha[lib], &sf), HR = log(sf);
  if (tune) win_push0(c->w_alpha[lib]);
  double prop[ZZO_MAX_LIBS], delta[ZZO_MAX_LIBS]; uint8_t mask[ZZO_MAX_LIBS] = {0};
  for (int r = 0; r < R; ++r) prop[r] = c->h.alpha_r[r];
  prop[lib] = ap; mask[lib] = 1;
  zzo_px_delta(&c->h, &c->s, prop, mask, delta);                   /* Lp - Lc, Pr:223-224 */
  double pp = 0, pc = 0;
  for (int r = 0; r < R; ++r) { pp += dgamma_log(prop[r], c->pr.alpha_r_shape, c->pr.alpha_r_rate); pc += dgamma_log(c->h.alpha_r[r], c->pr.alpha_r_shape, c->pr.alpha_r_rate); }
  double Rr = c->h.temperature * (delta[lib] + pp - pc) + HR;
  if (log(c_unif(c)) < Rr && Rr != INFINITY) { zzo_px_commit(&c->h, &c->s, lib, ap); if (tune) c->w_alpha[lib][99] = 1; }
}

static void run_slot(zzo_chain *c, int p, int tune) {
  switch (p) {
    case 1: mv_gibbsMixtureWeights(c); break;
    case 2: mv_allocAI(c); break;
    case 3: mv_allocWithin(c); break;
    case 4: mv_mhInactiveMeans(c, tune); break;
    case 5: break; /* mhInactiveVariances: weight 7*(1-shared_variance) = 0 under the default shared_variance=TRUE */
    case 6: mv_mhActiveMeansDif(c, tune); break;
    case 7: mv_mhSharedVariance(c, tune); break;
    case 8: mv_gibbsSpikeProb(c); break;
    case 9: mv_mhSpikeAllocation(c); break;
    case 10: mv_mhYg(c, tune); break;
    case 11: mv_mhVariance_g(c, tune); break;
    case 12: mv_mhTau(c, tune); break;
    case 13: mv_mhSg(c, tune); break;
    case 14: mv_mhS0Tau(c, tune); break;
    case 15: mv_mhP_x(c, tune); break;
  }
  c->step++;
}

static void tune_all(zzo_chain *c, double target) { /* U:183-240 */
  for (int r = 0; r < c->R; ++r) c->t_alpha[r] = zzo_x_tune(win_sum(c->w_alpha[r]), c->t_alpha[r], target, 0.01, 2);
  c->t_s0 = zzo_x_tune(win_sum(c->w_s0), c->t_s0, target, 0.0001, 5);
  c->t_s1 = zzo_x_tune(win_sum(c->w_s1), c->t_s1, target, 0.0001, 5);
  c->t_tau = zzo_x_tune(win_sum(c->w_tau), c->t_tau, target, 0.0001, 5);
  c->t_s0tau = zzo_x_tune(win_sum(c->w_s0tau), c->t_s0tau, target, 0.0001, 5);
  zzo_tune_per_gene(&c->s, target);
  c->t_im = zzo_x_tune(win_sum(c->w_im), c->t_im, target, 0.001, 10);
  for (int k = 0; k < c->K; ++k) c->t_am[k] = zzo_x_tune(win_sum(c->w_am[k]), c->t_am[k], target, 0.01, 10);
  c->t_av = zzo_x_tune(win_sum(c->w_av), c->t_av, target, 0.01, c->pr.variances_prior_log_max - c->pr.variances_prior_log_min);
}

zzo_chain *zzo_chain_new(int64_t G, int R, int K, const double *Xg, const double *gl, const zzo_priors *pr, uint64_t seed) {
  zzo_chain *c = (zzo_chain *)calloc(1, sizeof(zzo_chain));
  c->G = G; c->R = R; c->K = K; c->pr = *pr; c->seed = seed;
  c->Xg_own = (double *)malloc(sizeof(double) * G * R); memcpy(c->Xg_own, Xg, sizeof(double) * G * R);
  c->gl_own = (double *)malloc(sizeof(double) * G); memcpy(c->gl_own, gl, sizeof(double) * G);
  c->nd = (int32_t *)calloc(G, sizeof(int32_t));
  zzo_state *s = &c->s; s->G = G; s->Xg = c->Xg_own; s->gene_lengths = c->gl_own; s->no_detect_spike = c->nd; s->pinned_active = NULL;
  s->Yg = (double *)calloc(G, 8); s->variance_g = (double *)calloc(G, 8);
  s->alloc_active_inactive = (int32_t *)calloc(G, 4); s->alloc_within_active = (int32_t *)calloc(G, 4);
  s->inactive_spike_allocation = (int32_t *)calloc(G, 4);
  s->XgLikelihood = (double *)calloc(G, 8); s->variance_g_probability = (double *)calloc(G, 8);
  s->tuningParam_yg = (double *)calloc(G, 8); s->tuningParam_variance_g = (double *)calloc(G, 8);
  s->Yg_trace = (uint64_t *)calloc(2 * G, 8); s->variance_g_trace = (uint64_t *)calloc(2 * G, 8);
  c->alloc_trace = (uint32_t *)calloc((size_t)(K + 1) * G, 4);
  zzo_hyper *h = &c->h; h->num_libraries = R; h->num_active_components = K;
  h->beta = 1; h->temperature = 1; h->variance_g_upper_bound = INFINITY;
  /* I:185-199 */
  c->t_im = 0.5; c->t_av = 0.5; c->t_s0 = c->t_s1 = c->t_tau = c->t_s0tau = 0.05;
  for (int k = 0; k < K; ++k) c->t_am[k] = 0.5;
  for (int r = 0; r < R; ++r) c->t_alpha[r] = 0.5;
  for (int64_t g = 0; g < G; ++g) { s->tuningParam_yg[g] = 0.5; s->tuningParam_variance_g[g] = 0.5; }
  win_init(c->w_im); win_init(c->w_av); win_init(c->w_s0); win_init(c->w_s1); win_init(c->w_tau); win_init(c->w_s0tau);
  for (int k = 0; k < K; ++k) win_init(c->w_am[k]);
  for (int r = 0; r < R; ++r) win_init(c->w_alpha[r]);
  zzo_window_init(s->Yg_trace, G); zzo_window_init(s->variance_g_trace, G);
  /* I:225-237 (shared_variance = shared_active_variance = TRUE) */
  double yw = 1 + (R == 2) * 0.5 + 1 * (G < 15000);
  double pp[15] = {4, 30, 5, 6, 0, 5, 5, 2, 5, yw, yw, 8, 10, 8, R * 1.1};
  memcpy(c->proposal_probs, pp, sizeof pp);
  /* I:261-413 initial hyper-parameters drawn from (mostly) the priors */
  h->weight_active = c_beta(c, pr->weight_active_shape_1 + 1, pr->weight_active_shape_2 + 1);
  h->spike_probability = c_beta(c, pr->spike_prior_shape_1, pr->spike_prior_shape_2);
  for (int k = 0; k < K; ++k) {
    c->active_means_dif[k] = c_gamma(c, pr->active_means_dif_prior_rate, pr->active_means_dif_prior_rate); /* I:282 (shape = rate, sic) */
    h->active_means[k] = c->active_means_dif[k] + pr->threshold_a[k];
  }
  double av = pow(10.0, pr->variances_prior_log_min + (pr->variances_prior_log_max - pr->variances_prior_log_min) * c_unif(c));
  for (int k = 0; k < K; ++k) h->active_variances[k] = av;
  h->inactive_means = pr->threshold_i - c_gamma(c, 1, 1);
  h->inactive_variances = av;
  { double gs[ZZO_MAX_COMP], t = 0; for (int k = 0; k < K; ++k) { gs[k] = c_gamma(c, pr->weight_active_shape_1 / K, 1); t += gs[k]; }
    for (int k = 0; k < K; ++k) h->weight_within_active[k] = gs[k] / t; }
  for (int64_t g = 0; g < G; ++g) {
    s->alloc_active_inactive[g] = c_unif(c) < h->weight_active;
    s->alloc_within_active[g] = 1 + c_sample_w(c, h->weight_within_active, K);
  }
  for (int r = 0; r < R; ++r) h->alpha_r[r] = c_gamma(c, 1, 1);
  h->s0 = 0.1 * c_norm(c); h->s1 = -c_gamma(c, 1, 5); h->tau = c_gamma(c, 2, 5);
  /* I:416-424 */
  for (int64_t g = 0; g < G; ++g) {
    int nz = 0; for (int r = 0; r < R; ++r) nz += (Xg[(int64_t)r * G + g] == -INFINITY);
    c->nd[g] = (nz == R); s->inactive_spike_allocation[g] = c->nd[g];
    if (c->nd[g]) s->alloc_active_inactive[g] = 0;
  }
  /* I:428-468: Yg ~ N(row mean, sqrt(row var)) of detected values */
  double mv = 0, vv = 0; int64_t nfin = 0; /* mean / var of rowVars over fully detected rows (na.rm) */
  double *rv = (double *)malloc(8 * G);
  for (int64_t g = 0; g < G; ++g) {
    int nz = 0; double m = 0; for (int r = 0; r < R; ++r) { double x = Xg[(int64_t)r * G + g]; if (x == -INFINITY) nz++; else m += x; }
    rv[g] = NAN;
    if (nz == 0) { m /= R; double q = 0; for (int r = 0; r < R; ++r) { double d = Xg[(int64_t)r * G + g] - m; q += d * d; } rv[g] = q / (R - 1); mv += rv[g]; nfin++; }
  }
  mv /= (nfin > 0 ? nfin : 1);
  for (int64_t g = 0; g < G; ++g) if (!isnan(rv[g])) vv += (rv[g] - mv) * (rv[g] - mv);
  vv /= (nfin > 1 ? nfin - 1 : 1);
  for (int64_t g = 0; g < G; ++g) {
    int nd = 0; double m = 0; for (int r = 0; r < R; ++r) { double x = Xg[(int64_t)r * G + g]; if (x != -INFINITY) { nd++; m += x; } }
    double rwm, rwv;
    if (nd > 0) rwm = m / nd; else rwm = h->inactive_means + sqrt(h->inactive_variances) * c_norm(c);
    if (nd >= 2) { double q = 0; for (int r = 0; r < R; ++r) { double x = Xg[(int64_t)r * G + g]; if (x != -INFINITY) q += (x - rwm) * (x - rwm); } rwv = q / (nd - 1); }
    else rwv = exp(log(mv) + sqrt(vv) * c_norm(c));
    s->Yg[g] = rwm + sqrt(rwv) * c_norm(c);
    s->variance_g[g] = exp(0.5 + 0.2 * c_norm(c));
  }
  free(rv);
  zzo_eval_xg_loglik(h, s, s->XgLikelihood);
  for (int it = 0; it < 1000; ++it) { /* I:480-502 */
    int64_t bad = 0;
    for (int64_t g = 0; g < G; ++g) if (s->XgLikelihood[g] == -INFINITY) {
      bad++;
      s->Yg[g] = h->inactive_means + sqrt(h->inactive_variances) * c_norm(c);
      s->variance_g[g] = exp(log(zzo_get_sigmax(h->s0, h->s1, s->Yg[g])) + h->tau + sqrt(h->tau) * c_norm(c));
    }
    if (!bad) break;
    zzo_eval_xg_loglik(h, s, s->XgLikelihood);
  }
  zzo_eval_varg_logprior(h, s, s->variance_g_probability);
  return c;
}

void zzo_chain_free(zzo_chain *c) {
  if (!c) return;
  zzo_state *s = &c->s;
  free(c->Xg_own); free(c->gl_own); free(c->nd); free(s->Yg); free(s->variance_g); free(s->alloc_active_inactive);
  free(s->alloc_within_active); free(s->inactive_spike_allocation); free(s->XgLikelihood); free(s->variance_g_probability);
  free(s->tuningParam_yg); free(s->tuningParam_variance_g); free(s->Yg_trace); free(s->variance_g_trace);
  free(c->alloc_trace); free(c->draws); free(c);
}

static void one_generation_ref(zzo_chain *c, int tune) { /* B:66-67 / M:172-173 */
  int order[15];
  for (int i = 0; i < 15; ++i) order[i] = 1 + c_sample_w(c, c->proposal_probs, 15);
  for (int i = 0; i < 15; ++i) run_slot(c, order[i], tune);
}

static void accumulate_trace(zzo_chain *c) { /* M:194-199 */
  for (int64_t g = 0; g < c->G; ++g) {
    int comp = c->s.alloc_active_inactive[g] ? c->s.alloc_within_active[g] : 0;
    c->alloc_trace[(int64_t)comp * c->G + g] += 1;
  }
  c->n_samples++;
}

void zzo_chain_burnin(zzo_chain *c, int ngen, double target) {
  for (int i = 0; i < ngen; ++i) {
    one_generation_ref(c, 1);
    c->gen++;
    if (c->gen % 400 == 0) tune_all(c, target);                    /* B:146 */
  }
}

void zzo_chain_mcmc(zzo_chain *c, int ngen, int sample_frequency) {
  if (c->n_samples == 0) accumulate_trace(c);                      /* M:142-145: trace starts at the current allocation */
  for (int i = 0; i < ngen; ++i) {
    one_generation_ref(c, 0);
    if (c->gen % sample_frequency == 0) accumulate_trace(c);       /* M:188-199 */
    c->gen++;
  }
}

void zzo_chain_run_fused(zzo_chain *c, int ngen, int sample_frequency, int tune, double target) {
  double *draws = (double *)malloc(sizeof(double) * 9 * c->G);
  if (!tune && c->n_samples == 0) accumulate_trace(c);
  for (int i = 0; i < ngen; ++i) {
    zzo_sweep_draws(c->seed, c->chain_index, c->step, 0, c->G, draws);
    zzo_sweep(&c->h, &c->s, draws, tune, NULL);
    c->sub_step = c->step; c->sub_on = 1;                          /* hyper draws: sub-streams of this generation's sweep step */
    c->step++;
    static const int hyper_slots[] = {1, 4, 6, 7, 8, 12, 13, 14, 15};
    static const uint32_t sub_of[] = {SUB_MIXW, SUB_IM, SUB_AMD, SUB_SV, SUB_SP0, SUB_TAU, SUB_SG, SUB_S0TAU, SUB_PX};
    for (unsigned j = 0; j < sizeof hyper_slots / sizeof hyper_slots[0]; ++j) { sub_begin(c, sub_of[j]); run_slot(c, hyper_slots[j], tune); }
    c->sub_on = 0;
    if (!tune && c->gen % sample_frequency == 0) accumulate_trace(c);
    c->gen++;
    if (tune && c->gen % 400 == 0) tune_all(c, target);
  }
  free(draws);
}

void zzo_chain_prob_active(const zzo_chain *c, double *out) { /* FileWriting:116-118: 1 - trace[,1]/n */
  for (int64_t g = 0; g < c->G; ++g) out[g] = 1.0 - (double)c->alloc_trace[g] / (double)c->n_samples;
}
const zzo_hyper *zzo_chain_hyper(const zzo_chain *c) { return &c->h; }
void zzo_chain_export_scalars(const zzo_chain *c, zzo_chain_scalars *o) {
  memset(o, 0, sizeof *o);
  memcpy(o->active_means_dif, c->active_means_dif, sizeof o->active_means_dif);
  o->t_im = c->t_im; o->t_av = c->t_av; memcpy(o->t_am, c->t_am, sizeof o->t_am);
  o->t_s0 = c->t_s0; o->t_s1 = c->t_s1; o->t_tau = c->t_tau; o->t_s0tau = c->t_s0tau; memcpy(o->t_alpha, c->t_alpha, sizeof o->t_alpha);
  memcpy(o->w_im, c->w_im, 100); memcpy(o->w_av, c->w_av, 100); memcpy(o->w_am, c->w_am, sizeof o->w_am);
  memcpy(o->w_s0, c->w_s0, 100); memcpy(o->w_s1, c->w_s1, 100); memcpy(o->w_tau, c->w_tau, 100); memcpy(o->w_s0tau, c->w_s0tau, 100);
  memcpy(o->w_alpha, c->w_alpha, sizeof o->w_alpha);
  o->hyper_ctr = c->hyper_ctr; memcpy(o->hyper_buf, c->hyper_buf, sizeof o->hyper_buf);
  o->step = c->step; o->gen = c->gen; o->n_samples = c->n_samples;
}
const uint32_t *zzo_chain_alloc_trace(const zzo_chain *c) { return c->alloc_trace; }
void zzo_chain_set_stream(zzo_chain *c, uint64_t seed, uint32_t chain_index) { c->seed = seed; c->chain_index = chain_index; }
zzo_state *zzo_chain_state(zzo_chain *c) { return &c->s; }
double zzo_chain_lnl(const zzo_chain *c) { return zzo_lnl(&c->h, &c->s); }

/* ------------------------------------------------------------------ posterior-predictive discrepancies
   R/zigzagPostPredictiveSimulation.R:3-156 restated with explicit Philox draws (DESIGN.md RNG): for gene g and library r,
   block 16 + r gives z = Box-Muller(w0, w1) (cos branch) of rnorm(Yg, sqrt(variance_g)) and u = w2 of the detection draw;
   block 15 gives the z of the simulated Yg.  `bins` = seq(min(ranges) - 2, max(ranges) + 2, by = 0.1), built by the caller
   (PostPred:9-17).  out = { W_L, W_U, B_rumsfeld[1], W_L_r[R], B_rumsfeld[2..R+1] } (the rows of the two log files). */
static int upper_idx(const double *bins, int nb, double x) { /* number of bins <= x, so that (x < bins[b]) <=> b >= idx */
  int lo = 0, hi = nb;
  while (lo < hi) { int mid = (lo + hi) >> 1; if (bins[mid] <= x) lo = mid + 1; else hi = mid; }
  return lo;
}
static double trapz(const double *x, const double *y, int n) { /* U:3-9 */
  acc_t a = 0; for (int i = 1; i < n; ++i) a += (x[i] - x[i - 1]) * (y[i] + y[i - 1]);
  return 0.5 * (double)a;
}
/* get_lowerLevel_wasserstein (PostPred:120-144) from per-library histograms: cnt[r][i] = #{detected x with upper_idx == i},
   ndet[r], model[r][i] = sum of p_x[g, r] over out-of-spike genes with upper_idx(Yg) == i, for the libraries lo..hi-1 */
static double lower_wass(const double *bins, int nb, int R, int lo, int hi, const double *cnt, const double *ndet, const double *model, double n_out) {
  double *diff = (double *)calloc(nb, sizeof(double)), *mx = (double *)calloc(nb, sizeof(double)), *mm = (double *)calloc(nb, sizeof(double));
  (void)R;
  for (int r = lo; r < hi; ++r) {
    acc_t cx = 0, cm = 0; double maxm = -INFINITY;
    double *mc = (double *)malloc(sizeof(double) * nb);
    for (int b = 0; b < nb; ++b) { cm += model[(size_t)r * (nb + 1) + b]; mc[b] = (double)cm / n_out; if (mc[b] > maxm) maxm = mc[b]; }
    for (int b = 0; b < nb; ++b) {
      cx += cnt[(size_t)r * (nb + 1) + b];
      mx[b] += ((double)cx / ndet[r]) / (hi - lo);
      mm[b] += (mc[b] / maxm) / (hi - lo);
    }
    free(mc);
  }
  for (int b = 0; b < nb; ++b) diff[b] = fabs(mx[b] - mm[b]);
  double w = trapz(bins, diff, nb);
  free(diff); free(mx); free(mm);
  return w;
}
void zzo_post_predictive(const zzo_hyper *h, const zzo_state *s, uint64_t seed, uint64_t step, const double *bins, int nb, double *out) {
  const int R = h->num_libraries, K = h->num_active_components; const int64_t G = s->G; const size_t H = (size_t)(nb + 1);
  double *cd = (double *)calloc(R * H, 8), *cs = (double *)calloc(R * H, 8), *cds = (double *)calloc(R * H, 8), *css = (double *)calloc(R * H, 8);
  double *md = (double *)calloc(R * H, 8), *hy = (double *)calloc(H, 8), *hsy = (double *)calloc(H, 8);
  double ndd[ZZO_MAX_LIBS] = {0}, nds[ZZO_MAX_LIBS] = {0}, zd[ZZO_MAX_LIBS] = {0}, zs[ZZO_MAX_LIBS] = {0}, n_out = 0;
  acc_t sxd[ZZO_MAX_LIBS] = {0}, sxs[ZZO_MAX_LIBS] = {0}, s1mp[ZZO_MAX_LIBS] = {0};
  double *sim = (double *)malloc(sizeof(double) * G * R);
  for (int64_t g = 0; g < G; ++g) {
    const int sp = s->inactive_spike_allocation[g]; const double y = s->Yg[g], sd = sqrt(s->variance_g[g]);
    double px[ZZO_MAX_LIBS]; zzo_get_px(h, y, s->gene_lengths[g], px);
    if (!sp) { n_out += 1; hy[upper_idx(bins, nb, y)] += 1; }
    { double a[4], n[2]; zzo_uniform4(seed, 0, step, 15, (uint64_t)g, a); zzo_box_muller(a[0], a[1], n);   /* PostPred:79-84 */
      double sy;
      if (s->alloc_active_inactive[g]) { int k = s->alloc_within_active[g] - 1; sy = h->active_means[k] + sqrt(h->active_variances[k]) * n[0]; }
      else sy = h->inactive_means + sqrt(h->inactive_variances) * n[0];
      if (!sp) hsy[upper_idx(bins, nb, sy)] += 1; }
    for (int r = 0; r < R; ++r) {
      const double x = s->Xg[(int64_t)r * G + g];
      if (x == -INFINITY) zd[r] += 1; else { cd[r * H + upper_idx(bins, nb, x)] += 1; ndd[r] += 1; sxd[r] += x; }
      if (!sp) { md[r * H + upper_idx(bins, nb, y)] += px[r]; s1mp[r] += 1 - px[r]; }
      double a[4], n[2]; zzo_uniform4(seed, 0, step, 16 + r, (uint64_t)g, a); zzo_box_muller(a[0], a[1], n);
      double sx = y + sd * n[0];                                         /* PostPred:24 */
      if (a[2] < 1 - px[r]) sx = -INFINITY;                              /* PostPred:30-31 */
      if (sp) sx = -INFINITY;                                            /* PostPred:38 */
      sim[(int64_t)r * G + g] = sx;
      if (sx == -INFINITY) zs[r] += 1; else { cs[r * H + upper_idx(bins, nb, sx)] += 1; nds[r] += 1; sxs[r] += sx; }
    }
  }
  /* scaled copies: x - lib mean + grand mean (PostPred:52-58) */
  double lmd[ZZO_MAX_LIBS], lms[ZZO_MAX_LIBS], gmd = 0, gms = 0;
  for (int r = 0; r < R; ++r) { lmd[r] = (double)sxd[r] / ndd[r]; lms[r] = (double)sxs[r] / nds[r]; gmd += lmd[r] / R; gms += lms[r] / R; }
  for (int64_t g = 0; g < G; ++g) for (int r = 0; r < R; ++r) {
    const double x = s->Xg[(int64_t)r * G + g], sx = sim[(int64_t)r * G + g];
    if (x != -INFINITY) cds[r * H + upper_idx(bins, nb, x - lmd[r] + gmd)] += 1;
    if (sx != -INFINITY) css[r * H + upper_idx(bins, nb, sx - lms[r] + gms)] += 1;
  }
  out[0] = lower_wass(bins, nb, R, 0, R, cd, ndd, md, n_out) - lower_wass(bins, nb, R, 0, R, cs, nds, md, n_out);     /* W_L, PostPred:43-44 */
  for (int r = 0; r < R; ++r)                                                                                         /* W_L_r, PostPred:61-66 */
    out[3 + r] = lower_wass(bins, nb, R, r, r + 1, cds, ndd, md, n_out) - lower_wass(bins, nb, R, r, r + 1, css, nds, md, n_out);
  /* get_rumsfeld (PostPred:107-118) */
  { const double c = h->spike_probability * (1 - h->weight_active); double md_ = 0, ms_ = 0, dd[ZZO_MAX_LIBS], ds[ZZO_MAX_LIBS];
    for (int r = 0; r < R; ++r) {
      const double model = (1 - c) * ((double)s1mp[r] / n_out) + c;
      dd[r] = fabs(zd[r] / (double)G - model); ds[r] = fabs(zs[r] / (double)G - model); md_ += dd[r] / R; ms_ += ds[r] / R;
    }
    out[2] = md_ - ms_;
    for (int r = 0; r < R; ++r) out[3 + R + r] = dd[r] - ds[r]; }
  /* get_upperLevel_wasserstein (PostPred:146-154) with get_model_cumulative (U:20-40) */
  { double *dens = (double *)malloc(8 * nb), *mc = (double *)malloc(8 * nb), *dy = (double *)malloc(8 * nb), *dsy = (double *)malloc(8 * nb);
    const double wi = (1 - h->weight_active) * (1 - h->spike_probability), wa = h->weight_active;
    for (int b = 0; b < nb; ++b) {
      acc_t a = 0;
      for (int k = 0; k < K; ++k) a += h->weight_within_active[k] * exp(zzo_dnorm_log(bins[b], h->active_means[k], sqrt(h->active_variances[k])));
      dens[b] = (wi * exp(zzo_dnorm_log(bins[b], h->inactive_means, sqrt(h->inactive_variances))) + wa * (double)a) / (wi + wa);
    }
    acc_t cum = 0, cy = hy[0], csy = hsy[0];
    for (int b = 1; b < nb; ++b) {
      cum += (bins[b] - bins[b - 1]) * (dens[b] + dens[b - 1]); mc[b] = 0.5 * (double)cum;
      cy += hy[b]; csy += hsy[b];
      dy[b] = fabs((double)cy / n_out - mc[b]); dsy[b] = fabs((double)csy / n_out - mc[b]);
    }
    out[1] = trapz(bins + 1, dy + 1, nb - 1) - trapz(bins + 1, dsy + 1, nb - 1);
    free(dens); free(mc); free(dy); free(dsy); }
  free(cd); free(cs); free(cds); free(css); free(md); free(hy); free(hsy); free(sim);
}
