/*
 * zz_shim.c — the `.Call` shim an R maintainer adds to the zigzag package (as src/zz_shim.c) to run the per-gene moves on
 * the GPU through the C ABI of include/zigzag_gpu.h: the reference-side binding described in INTEGRATION.md.  Link with
 * -lzigzag_gpu.  R is not installed in this image; tests/test_r_overlay.py compiles this file against minimal stand-ins of the R
 * headers (tests/r_stub/) and checks every entry of the registration table against its definition and against every .Call in
 * r_overlay/R/zigzagGPU.R (names and argument counts).
 *
 * Conventions: R numeric = double, R integer/logical = int32, matrices column-major (Xg goes through unchanged), the
 * handle lives in an external pointer whose finalizer calls zz_destroy.  Errors are raised with Rf_error only after every
 * C resource of the call has been released (Rf_error longjmps).
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <string.h>
#include "zigzag_gpu.h"

static zz_handle *H(SEXP ptr) {
  zz_handle *h = (zz_handle *)R_ExternalPtrAddr(ptr);
  if (!h) Rf_error("zigzag GPU handle is NULL (object was copied or already released)");
  return h;
}
static void chk(zz_handle *h, int rc) { if (rc != ZZ_OK) Rf_error("zigzag_gpu: %s (code %d)", zz_last_error(h), rc); }
static void fin(SEXP ptr) { zz_handle *h = (zz_handle *)R_ExternalPtrAddr(ptr); if (h) { zz_destroy(h); R_ClearExternalPtr(ptr); } }
static const double *dbl(SEXP x) { return Rf_isNull(x) ? NULL : REAL(x); }
static const int *intg(SEXP x) { return Rf_isNull(x) ? NULL : INTEGER(x); }

/* zz_create + zz_upload_data: called once at the end of zigzag$new() (R/zigzagInitialize.R:564) */
SEXP zzR_create(SEXP Xg, SEXP gene_lengths, SEXP K, SEXP seed, SEXP device) {
  SEXP dim = Rf_getAttrib(Xg, R_DimSymbol);
  zz_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.num_genes = INTEGER(dim)[0]; cfg.num_libraries = INTEGER(dim)[1]; cfg.num_active_components = Rf_asInteger(K);
  cfg.num_chains = 1; cfg.device = Rf_asInteger(device); cfg.seed = (uint64_t)Rf_asReal(seed); cfg.gene_offset = 0;
  zz_handle *h = NULL;
  int rc = zz_create(&cfg, &h);
  if (rc != ZZ_OK) { char msg[512]; strncpy(msg, h ? zz_last_error(h) : "invalid configuration", 511); msg[511] = 0; if (h) zz_destroy(h); Rf_error("zz_create: %s", msg); }
  rc = zz_upload_data(h, REAL(Xg), REAL(gene_lengths));
  if (rc != ZZ_OK) { char msg[512]; strncpy(msg, zz_last_error(h), 511); msg[511] = 0; zz_destroy(h); Rf_error("zz_upload_data: %s", msg); }
  SEXP ptr = PROTECT(R_MakeExternalPtr(h, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(ptr, fin, TRUE);
  UNPROTECT(1);
  return ptr;
}

/* fields -> device; any argument may be NULL */
SEXP zzR_set_state(SEXP ptr, SEXP Yg, SEXP vg, SEXP ai, SEXP within, SEXP spike, SEXP xglik, SEXP vgprob, SEXP ty, SEXP tv, SEXP pinned) {
  zz_handle *h = H(ptr);
  chk(h, zz_set_state(h, 0, dbl(Yg), dbl(vg), intg(ai), intg(within), intg(spike), dbl(xglik), dbl(vgprob), dbl(ty), dbl(tv), intg(pinned)));
  return R_NilValue;
}

/* device -> a named list of the per-gene fields (lazy pull at sample / write time) */
SEXP zzR_get_state(SEXP ptr, SEXP G_) {
  zz_handle *h = H(ptr);
  const int G = Rf_asInteger(G_);
  const char *names[] = {"Yg", "variance_g", "allocation_active_inactive", "allocation_within_active", "inactive_spike_allocation",
                         "XgLikelihood", "variance_g_probability", "tuningParam_yg", "tuningParam_variance_g", ""};
  SEXP out = PROTECT(Rf_mkNamed(VECSXP, names));
  for (int i = 0; i < 9; ++i) SET_VECTOR_ELT(out, i, Rf_allocVector((i >= 2 && i <= 4) ? INTSXP : REALSXP, G));
  int rc = zz_get_state(h, 0, REAL(VECTOR_ELT(out, 0)), REAL(VECTOR_ELT(out, 1)), INTEGER(VECTOR_ELT(out, 2)), INTEGER(VECTOR_ELT(out, 3)),
                        INTEGER(VECTOR_ELT(out, 4)), REAL(VECTOR_ELT(out, 5)), REAL(VECTOR_ELT(out, 6)), REAL(VECTOR_ELT(out, 7)),
                        REAL(VECTOR_ELT(out, 8)));
  UNPROTECT(1);
  chk(h, rc);
  return out;
}

/* scalar fields -> zz_hyper; called by every accepted hyper-parameter move */
SEXP zzR_set_hyper(SEXP ptr, SEXP s0, SEXP s1, SEXP tau, SEXP alpha_r, SEXP spike_p, SEXP im, SEXP iv, SEXP am, SEXP av, SEXP wa,
                   SEXP ww, SEXP beta, SEXP temperature, SEXP vub) {
  zz_handle *h = H(ptr);
  zz_hyper hy;
  memset(&hy, 0, sizeof hy);
  hy.num_libraries = LENGTH(alpha_r); hy.num_active_components = LENGTH(am);
  hy.s0 = Rf_asReal(s0); hy.s1 = Rf_asReal(s1); hy.tau = Rf_asReal(tau);
  memcpy(hy.alpha_r, REAL(alpha_r), sizeof(double) * hy.num_libraries);
  hy.spike_probability = Rf_asReal(spike_p); hy.inactive_means = Rf_asReal(im); hy.inactive_variances = Rf_asReal(iv);
  memcpy(hy.active_means, REAL(am), sizeof(double) * hy.num_active_components);
  memcpy(hy.active_variances, REAL(av), sizeof(double) * hy.num_active_components);
  memcpy(hy.weight_within_active, REAL(ww), sizeof(double) * hy.num_active_components);
  hy.weight_active = Rf_asReal(wa); hy.beta = Rf_asReal(beta); hy.temperature = Rf_asReal(temperature);
  hy.variance_g_upper_bound = Rf_asReal(vub);
  chk(h, zz_set_hyper(h, 0, &hy));
  return R_NilValue;
}

/* the five per-gene moves: proposal_list[[10]], [[11]], [[2]], [[3]], [[9]] (R/zigzagInitialize.R:207-221) */
SEXP zzR_mh_yg(SEXP ptr, SEXP step, SEXP tune) { zz_handle *h = H(ptr); chk(h, zz_mh_yg(h, (uint64_t)Rf_asReal(step), Rf_asLogical(tune), NULL, NULL)); return R_NilValue; }
SEXP zzR_mh_variance_g(SEXP ptr, SEXP step, SEXP tune) { zz_handle *h = H(ptr); chk(h, zz_mh_variance_g(h, (uint64_t)Rf_asReal(step), Rf_asLogical(tune), NULL, NULL)); return R_NilValue; }
SEXP zzR_gibbs_alloc(SEXP ptr, SEXP step) { zz_handle *h = H(ptr); chk(h, zz_gibbs_alloc(h, (uint64_t)Rf_asReal(step), NULL, NULL)); return R_NilValue; }
SEXP zzR_gibbs_alloc_within(SEXP ptr, SEXP step) { zz_handle *h = H(ptr); chk(h, zz_gibbs_alloc_within(h, (uint64_t)Rf_asReal(step), NULL, NULL)); return R_NilValue; }
SEXP zzR_mh_spike_alloc(SEXP ptr, SEXP step) { zz_handle *h = H(ptr); chk(h, zz_mh_spike_alloc(h, (uint64_t)Rf_asReal(step), NULL, NULL)); return R_NilValue; }
SEXP zzR_sweep(SEXP ptr, SEXP step, SEXP tune) { zz_handle *h = H(ptr); chk(h, zz_sweep(h, (uint64_t)Rf_asReal(step), Rf_asLogical(tune), 0, NULL, NULL)); return R_NilValue; }

/* reductions consumed by the hyper-parameter moves */
SEXP zzR_suffstats(SEXP ptr, SEXP K) {
  zz_handle *h = H(ptr);
  SEXP out = PROTECT(Rf_allocVector(REALSXP, zz_suffstats_len(Rf_asInteger(K))));
  int rc = zz_suffstats(h, REAL(out));
  UNPROTECT(1); chk(h, rc);
  return out;
}
SEXP zzR_varg_hyper_sum(SEXP ptr, SEXP s0, SEXP s1, SEXP tau) {
  zz_handle *h = H(ptr);
  SEXP out = PROTECT(Rf_allocVector(REALSXP, 2));
  int rc = zz_varg_hyper_sum(h, 0, Rf_asReal(s0), Rf_asReal(s1), Rf_asReal(tau), REAL(out));
  UNPROTECT(1); chk(h, rc);
  return out;
}
SEXP zzR_commit_varg_hyper(SEXP ptr) { zz_handle *h = H(ptr); chk(h, zz_commit_varg_hyper(h, 0)); return R_NilValue; }
SEXP zzR_level2_sums(SEXP ptr, SEXP im, SEXP iv, SEXP am, SEXP av) {
  zz_handle *h = H(ptr);
  SEXP out = PROTECT(Rf_allocVector(REALSXP, 2));
  int rc = zz_level2_sums(h, 0, Rf_asReal(im), Rf_asReal(iv), REAL(am), REAL(av), REAL(out));
  UNPROTECT(1); chk(h, rc);
  return out;
}
SEXP zzR_px_delta(SEXP ptr, SEXP alpha_prop, SEXP lib) {   /* lib is 1-based */
  zz_handle *h = H(ptr);
  const int R = LENGTH(alpha_prop);
  unsigned char mask[ZZ_MAX_LIBS];
  memset(mask, 0, sizeof mask); mask[Rf_asInteger(lib) - 1] = 1;
  SEXP out = PROTECT(Rf_allocVector(REALSXP, R));
  int rc = zz_px_delta(h, 0, REAL(alpha_prop), mask, REAL(out));
  UNPROTECT(1); chk(h, rc);
  return out;
}
SEXP zzR_px_commit(SEXP ptr, SEXP lib, SEXP alpha_new) { zz_handle *h = H(ptr); chk(h, zz_px_commit(h, 0, Rf_asInteger(lib) - 1, Rf_asReal(alpha_new))); return R_NilValue; }
SEXP zzR_lnl(SEXP ptr) { zz_handle *h = H(ptr); double v = 0; chk(h, zz_lnl(h, 0, &v)); return Rf_ScalarReal(v); }
SEXP zzR_tune(SEXP ptr, SEXP target) { zz_handle *h = H(ptr); chk(h, zz_tune(h, Rf_asReal(target))); return R_NilValue; }
SEXP zzR_nan_flag(SEXP ptr) { zz_handle *h = H(ptr); int f = 0; chk(h, zz_nan_flag(h, &f)); return Rf_ScalarLogical(f); }
SEXP zzR_accumulate_alloc_trace(SEXP ptr) { zz_handle *h = H(ptr); chk(h, zz_accumulate_alloc_trace(h)); return R_NilValue; }
SEXP zzR_reset_alloc_trace(SEXP ptr) { zz_handle *h = H(ptr); chk(h, zz_reset_alloc_trace(h)); return R_NilValue; }
SEXP zzR_get_alloc_trace(SEXP ptr, SEXP G_, SEXP K_) {   /* G x (K+1) numeric matrix, like the field allocation_trace */
  zz_handle *h = H(ptr);
  const int G = Rf_asInteger(G_), K1 = Rf_asInteger(K_) + 1;
  uint32_t *buf = (uint32_t *)R_alloc((size_t)G * K1, sizeof(uint32_t));
  chk(h, zz_get_alloc_trace(h, 0, buf));
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, G, K1));
  for (size_t i = 0; i < (size_t)G * K1; ++i) REAL(out)[i] = buf[i];   /* [K+1][G] == column-major G x (K+1) */
  UNPROTECT(1);
  return out;
}

/* ---- device-resident generation loop (zz_chain_begin / zz_run_generations / zz_chain_read) ------------------------------
   R side (r_overlay/R/zigzagGPU.R: gpu_run_generations): priors = numeric(19 + K) in the field order of zz_priors; scalars = a list
   with the fields of zz_chain_scalars (numeric vectors and 100-wide 0/1 integer windows).  Returns the list updated. */
static void fill_win(uint8_t *dst, SEXP v) { for (int i = 0; i < 100; ++i) dst[i] = (uint8_t)(INTEGER(v)[i] != 0); }
static SEXP get_el(SEXP list, const char *name) {
  SEXP names = Rf_getAttrib(list, R_NamesSymbol);
  for (int i = 0; i < LENGTH(list); ++i) if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) return VECTOR_ELT(list, i);
  Rf_error("zzR_chain_begin: scalar state has no field '%s'", name);
  return R_NilValue;
}
SEXP zzR_chain_begin(SEXP ptr, SEXP priors, SEXP sc, SEXP K_, SEXP R_) {
  zz_handle *h = H(ptr);
  const int K = Rf_asInteger(K_), R = Rf_asInteger(R_);
  zz_priors pr; zz_chain_scalars s;
  memset(&pr, 0, sizeof pr); memset(&s, 0, sizeof s);
  memcpy(&pr, REAL(priors), sizeof(double) * 19);                      /* 18 scalars + threshold_i, then threshold_a[K] */
  memcpy(pr.threshold_a, REAL(priors) + 19, sizeof(double) * K);
  memcpy(s.active_means_dif, REAL(get_el(sc, "active_means_dif")), sizeof(double) * K);
  s.t_im = Rf_asReal(get_el(sc, "t_im")); s.t_av = Rf_asReal(get_el(sc, "t_av"));
  memcpy(s.t_am, REAL(get_el(sc, "t_am")), sizeof(double) * K);
  s.t_s0 = Rf_asReal(get_el(sc, "t_s0")); s.t_s1 = Rf_asReal(get_el(sc, "t_s1")); s.t_tau = Rf_asReal(get_el(sc, "t_tau"));
  s.t_s0tau = Rf_asReal(get_el(sc, "t_s0tau"));
  memcpy(s.t_alpha, REAL(get_el(sc, "t_alpha")), sizeof(double) * R);
  fill_win(s.w_im, get_el(sc, "w_im")); fill_win(s.w_av, get_el(sc, "w_av")); fill_win(s.w_s0, get_el(sc, "w_s0"));
  fill_win(s.w_s1, get_el(sc, "w_s1")); fill_win(s.w_tau, get_el(sc, "w_tau")); fill_win(s.w_s0tau, get_el(sc, "w_s0tau"));
  for (int k = 0; k < K; ++k) fill_win(s.w_am[k], VECTOR_ELT(get_el(sc, "w_am"), k));
  for (int r = 0; r < R; ++r) fill_win(s.w_alpha[r], VECTOR_ELT(get_el(sc, "w_alpha"), r));
  s.hyper_ctr = (uint64_t)Rf_asReal(get_el(sc, "hyper_ctr"));
  memcpy(s.hyper_buf, REAL(get_el(sc, "hyper_buf")), sizeof s.hyper_buf);
  s.step = (uint64_t)Rf_asReal(get_el(sc, "step")); s.gen = (int64_t)Rf_asReal(get_el(sc, "gen"));
  s.n_samples = (int64_t)Rf_asReal(get_el(sc, "n_samples"));
  chk(h, zz_chain_begin(h, &pr, &s));
  return R_NilValue;
}
SEXP zzR_run_generations(SEXP ptr, SEXP ngen, SEXP tune, SEXP sample_frequency, SEXP target) {
  zz_handle *h = H(ptr);
  chk(h, zz_run_generations(h, Rf_asInteger(ngen), Rf_asLogical(tune), Rf_asInteger(sample_frequency), Rf_asReal(target)));
  return R_NilValue;
}
/* returns c(s0, s1, tau, spike_probability, inactive_means, inactive_variances, weight_active, gen, step, n_samples, hyper_ctr,
   alpha_r[R], active_means[K], active_variances[K], weight_within_active[K], active_means_dif[K], t_im, t_av, t_s0, t_s1, t_tau,
   t_s0tau, t_am[K], t_alpha[R]); the windows come back through zzR_chain_windows */
SEXP zzR_chain_read(SEXP ptr, SEXP K_, SEXP R_) {
  zz_handle *h = H(ptr);
  const int K = Rf_asInteger(K_), R = Rf_asInteger(R_);
  zz_chain_scalars s; zz_hyper hy;
  chk(h, zz_chain_read(h, &s, &hy));
  const int n = 11 + R + 4 * K + 6 + K + R;
  SEXP out = PROTECT(Rf_allocVector(REALSXP, n));
  double *o = REAL(out); int j = 0;
  o[j++] = hy.s0; o[j++] = hy.s1; o[j++] = hy.tau; o[j++] = hy.spike_probability; o[j++] = hy.inactive_means;
  o[j++] = hy.inactive_variances; o[j++] = hy.weight_active; o[j++] = (double)s.gen; o[j++] = (double)s.step;
  o[j++] = (double)s.n_samples; o[j++] = (double)s.hyper_ctr;
  for (int r = 0; r < R; ++r) o[j++] = hy.alpha_r[r];
  for (int k = 0; k < K; ++k) o[j++] = hy.active_means[k];
  for (int k = 0; k < K; ++k) o[j++] = hy.active_variances[k];
  for (int k = 0; k < K; ++k) o[j++] = hy.weight_within_active[k];
  for (int k = 0; k < K; ++k) o[j++] = s.active_means_dif[k];
  o[j++] = s.t_im; o[j++] = s.t_av; o[j++] = s.t_s0; o[j++] = s.t_s1; o[j++] = s.t_tau; o[j++] = s.t_s0tau;
  for (int k = 0; k < K; ++k) o[j++] = s.t_am[k];
  for (int r = 0; r < R; ++r) o[j++] = s.t_alpha[r];
  UNPROTECT(1);
  return out;
}


/* the 100-wide accept windows of the scalar moves after a device-resident segment: a named list of integer vectors (w_im, w_av,
   w_s0, w_s1, w_tau, w_s0tau), a K x 100 matrix w_am and an R x 100 matrix w_alpha — the shapes of the *_trace[[1]][[1]] fields */
static SEXP win_vec(const uint8_t *w) { SEXP v = PROTECT(Rf_allocVector(INTSXP, 100)); for (int i = 0; i < 100; ++i) INTEGER(v)[i] = w[i]; UNPROTECT(1); return v; }
SEXP zzR_chain_windows(SEXP ptr, SEXP K_, SEXP R_) {
  zz_handle *h = H(ptr);
  const int K = Rf_asInteger(K_), R = Rf_asInteger(R_);
  zz_chain_scalars s; zz_hyper hy;
  chk(h, zz_chain_read(h, &s, &hy));
  const char *names[] = {"w_im", "w_av", "w_s0", "w_s1", "w_tau", "w_s0tau", "w_am", "w_alpha", ""};
  SEXP out = PROTECT(Rf_mkNamed(VECSXP, names));
  const uint8_t *single[6] = {s.w_im, s.w_av, s.w_s0, s.w_s1, s.w_tau, s.w_s0tau};
  for (int i = 0; i < 6; ++i) SET_VECTOR_ELT(out, i, win_vec(single[i]));
  SEXP am = PROTECT(Rf_allocMatrix(INTSXP, K, 100)), al = PROTECT(Rf_allocMatrix(INTSXP, R, 100));
  for (int k = 0; k < K; ++k) for (int i = 0; i < 100; ++i) INTEGER(am)[k + (size_t)K * i] = s.w_am[k][i];
  for (int r = 0; r < R; ++r) for (int i = 0; i < 100; ++i) INTEGER(al)[r + (size_t)R * i] = s.w_alpha[r][i];
  SET_VECTOR_ELT(out, 6, am); SET_VECTOR_ELT(out, 7, al);
  UNPROTECT(3);
  return out;
}
SEXP zzR_run_sweeps(SEXP ptr, SEXP step, SEXP nsweeps, SEXP tune) {
  zz_handle *h = H(ptr);
  chk(h, zz_run_sweeps(h, (uint64_t)Rf_asReal(step), Rf_asInteger(nsweeps), Rf_asLogical(tune)));
  return R_NilValue;
}

/* ---- the remaining entry points of include/zigzag_gpu.h ------------------------------------------------------------------ */
/* constructor-side initial state on the device (R/zigzagInitialize.R:416-504): rowvar = c(log(mean(v)), sd(v)) of the row variances of
   the fully detected genes; the fields come back with gpu_pull() */
SEXP zzR_init_state(SEXP ptr, SEXP step, SEXP rowvar) {
  zz_handle *h = H(ptr);
  chk(h, zz_init_state(h, (uint64_t)Rf_asReal(step), REAL(rowvar)[0], REAL(rowvar)[1]));
  return R_NilValue;
}
/* chain-axis bookkeeping of tempered chains (handles created with more than one chain; 0-based chain slots) */
SEXP zzR_accumulate_alloc_trace_from(SEXP ptr, SEXP src, SEXP dst) { zz_handle *h = H(ptr); chk(h, zz_accumulate_alloc_trace_from(h, Rf_asInteger(src), Rf_asInteger(dst))); return R_NilValue; }
SEXP zzR_copy_chain_tuning(SEXP ptr, SEXP src, SEXP dst) { zz_handle *h = H(ptr); chk(h, zz_copy_chain_tuning(h, Rf_asInteger(src), Rf_asInteger(dst))); return R_NilValue; }
SEXP zzR_sync(SEXP ptr) { zz_handle *h = H(ptr); chk(h, zz_sync(h)); return R_NilValue; }
SEXP zzR_launch_count(SEXP ptr) { return Rf_ScalarReal((double)zz_launch_count(H(ptr))); }
SEXP zzR_refresh_caches(SEXP ptr) { zz_handle *h = H(ptr); chk(h, zz_refresh_caches(h)); return R_NilValue; }
SEXP zzR_reset_windows(SEXP ptr) { zz_handle *h = H(ptr); chk(h, zz_reset_windows(h)); return R_NilValue; }
/* list(yg = rowSums(Yg_trace), vg = rowSums(variance_g_trace)) of the device-side windows */
SEXP zzR_window_sums(SEXP ptr, SEXP G_) {
  zz_handle *h = H(ptr);
  const int G = Rf_asInteger(G_);
  const char *names[] = {"yg", "vg", ""};
  SEXP out = PROTECT(Rf_mkNamed(VECSXP, names));
  SET_VECTOR_ELT(out, 0, Rf_allocVector(INTSXP, G)); SET_VECTOR_ELT(out, 1, Rf_allocVector(INTSXP, G));
  int rc = zz_get_window_sums(h, 0, INTEGER(VECTOR_ELT(out, 0)), INTEGER(VECTOR_ELT(out, 1)));
  UNPROTECT(1); chk(h, rc);
  return out;
}
/* c(s0, s1, tau, spike_probability, inactive_means, inactive_variances, weight_active, beta, temperature, variance_g_upper_bound,
   alpha_r[R], active_means[K], active_variances[K], weight_within_active[K]) as the device holds them */
SEXP zzR_get_hyper(SEXP ptr) {
  zz_handle *h = H(ptr);
  zz_hyper hy;
  chk(h, zz_get_hyper(h, 0, &hy));
  const int R = hy.num_libraries, K = hy.num_active_components;
  SEXP out = PROTECT(Rf_allocVector(REALSXP, 10 + R + 3 * K));
  double *o = REAL(out); int j = 0;
  o[j++] = hy.s0; o[j++] = hy.s1; o[j++] = hy.tau; o[j++] = hy.spike_probability; o[j++] = hy.inactive_means; o[j++] = hy.inactive_variances;
  o[j++] = hy.weight_active; o[j++] = hy.beta; o[j++] = hy.temperature; o[j++] = hy.variance_g_upper_bound;
  for (int r = 0; r < R; ++r) o[j++] = hy.alpha_r[r];
  for (int k = 0; k < K; ++k) o[j++] = hy.active_means[k];
  for (int k = 0; k < K; ++k) o[j++] = hy.active_variances[k];
  for (int k = 0; k < K; ++k) o[j++] = hy.weight_within_active[k];
  UNPROTECT(1);
  return out;
}
/* the three per-gene evaluators (which = 1: computeXgLikelihood, 2: computeVarianceGPriorProbability, 3: level-2 likelihood) */
SEXP zzR_eval(SEXP ptr, SEXP which, SEXP G_) {
  zz_handle *h = H(ptr);
  SEXP out = PROTECT(Rf_allocVector(REALSXP, Rf_asInteger(G_)));
  const int w = Rf_asInteger(which);
  int rc = w == 1 ? zz_eval_xg_loglik(h, 0, REAL(out)) : w == 2 ? zz_eval_varg_logprior(h, 0, REAL(out)) : zz_eval_level2_loglik(h, 0, REAL(out));
  UNPROTECT(1); chk(h, rc);
  return out;
}
/* the uniforms / normals move `slot` consumes at `step` (north-star check 2: identical draws on both sides) */
SEXP zzR_philox_draws(SEXP ptr, SEXP step, SEXP slot, SEXP n_, SEXP G_) {
  zz_handle *h = H(ptr);
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, Rf_asInteger(G_), Rf_asInteger(n_)));
  int rc = zz_philox_draws(h, (uint64_t)Rf_asReal(step), Rf_asInteger(slot), REAL(out));
  UNPROTECT(1); chk(h, rc);
  return out;
}
/* zz_sweep_host: the fused sweep on state that lives in R vectors (updated IN PLACE: the caller must own unshared copies) */
SEXP zzR_sweep_host(SEXP ptr, SEXP step, SEXP tune, SEXP Yg, SEXP vg, SEXP ai, SEXP within, SEXP spike, SEXP xglik, SEXP vgprob, SEXP ty, SEXP tv) {
  zz_handle *h = H(ptr);
  int64_t up = 0, down = 0;
  chk(h, zz_sweep_host(h, (uint64_t)Rf_asReal(step), Rf_asLogical(tune), REAL(Yg), REAL(vg), INTEGER(ai), INTEGER(within), INTEGER(spike),
                       REAL(xglik), REAL(vgprob), REAL(ty), REAL(tv), &up, &down));
  SEXP out = PROTECT(Rf_allocVector(REALSXP, 2));
  REAL(out)[0] = (double)up; REAL(out)[1] = (double)down;
  UNPROTECT(1);
  return out;
}
/* posterior-predictive simulation on the device (R/zigzagPostPred.R:3-156): c(W_L, W_U, B, W_L_r[R], B_r[R]) */
SEXP zzR_post_predictive(SEXP ptr, SEXP step, SEXP bins, SEXP R_) {
  zz_handle *h = H(ptr);
  SEXP out = PROTECT(Rf_allocVector(REALSXP, 3 + 2 * Rf_asInteger(R_)));
  int rc = zz_post_predictive(h, (uint64_t)Rf_asReal(step), REAL(bins), LENGTH(bins), REAL(out));
  UNPROTECT(1); chk(h, rc);
  return out;
}
/* sample sink: the device appends the log rows of every sampling point of a device-resident segment to pinned host rings.
   The rings belong to the shim (one sink per process is enough for the drop-in: one zigzag object runs at a time). */
static struct { zz_sample_sink sk; int32_t *cand; int plen, ncand; } g_sink;
static void sink_free(zz_handle *h) {
  if (h) zz_set_sample_sink(h, NULL);
  zz_pinned_free(g_sink.sk.param_rows); zz_pinned_free(g_sink.sk.yg_rows); zz_pinned_free(g_sink.sk.vg_rows);
  memset(&g_sink, 0, sizeof g_sink);
}
SEXP zzR_set_sample_sink(SEXP ptr, SEXP param_capacity, SEXP candidates /* 1-based gene indices or NULL */, SEXP yv_capacity, SEXP yv_every, SEXP R_, SEXP K_) {
  zz_handle *h = H(ptr);
  sink_free(h);
  const int pc = Rf_asInteger(param_capacity), yc = Rf_isNull(candidates) ? 0 : Rf_asInteger(yv_capacity);
  g_sink.plen = zz_param_row_len(Rf_asInteger(R_), Rf_asInteger(K_));
  g_sink.sk.param_rows = (double *)zz_pinned_alloc(sizeof(double) * (size_t)(pc > 0 ? pc : 1) * g_sink.plen);
  g_sink.sk.param_capacity = pc;
  if (!g_sink.sk.param_rows) { sink_free(NULL); Rf_error("zz_pinned_alloc failed"); }
  if (yc > 0) {
    g_sink.ncand = LENGTH(candidates);
    g_sink.cand = (int32_t *)R_alloc((size_t)g_sink.ncand, sizeof(int32_t));
    for (int i = 0; i < g_sink.ncand; ++i) g_sink.cand[i] = INTEGER(candidates)[i] - 1;
    g_sink.sk.yg_rows = (double *)zz_pinned_alloc(sizeof(double) * (size_t)yc * (1 + g_sink.ncand));
    g_sink.sk.vg_rows = (double *)zz_pinned_alloc(sizeof(double) * (size_t)yc * (1 + g_sink.ncand));
    if (!g_sink.sk.yg_rows || !g_sink.sk.vg_rows) { sink_free(NULL); Rf_error("zz_pinned_alloc failed"); }
    g_sink.sk.yv_capacity = yc; g_sink.sk.candidates = g_sink.cand; g_sink.sk.n_candidates = g_sink.ncand; g_sink.sk.yv_every = Rf_asInteger(yv_every);
  }
  int rc = zz_set_sample_sink(h, &g_sink.sk);          /* copies the candidate list to the device */
  g_sink.sk.candidates = NULL; g_sink.cand = NULL;   /* R_alloc memory: gone when this call returns */
  if (rc != ZZ_OK) { sink_free(NULL); chk(h, rc); }
  return R_NilValue;
}
/* list(param = rows x plen, yg = rows x (1 + ncand), vg = ...) written so far, then detaches and frees the rings */
SEXP zzR_sample_rows(SEXP ptr) {
  zz_handle *h = H(ptr);
  int64_t np = 0, ny = 0;
  chk(h, zz_sync(h));
  chk(h, zz_sample_rows_written(h, &np, &ny));
  const char *names[] = {"param", "yg", "vg", ""};
  SEXP out = PROTECT(Rf_mkNamed(VECSXP, names));
  SEXP p = PROTECT(Rf_allocMatrix(REALSXP, (int)np, g_sink.plen));
  for (int64_t r = 0; r < np; ++r) for (int c = 0; c < g_sink.plen; ++c) REAL(p)[r + np * c] = g_sink.sk.param_rows[r * g_sink.plen + c];
  SET_VECTOR_ELT(out, 0, p);
  if (g_sink.sk.yg_rows) {
    const int w = 1 + g_sink.ncand;
    SEXP y = PROTECT(Rf_allocMatrix(REALSXP, (int)ny, w)), v = PROTECT(Rf_allocMatrix(REALSXP, (int)ny, w));
    for (int64_t r = 0; r < ny; ++r) for (int c = 0; c < w; ++c) { REAL(y)[r + ny * c] = g_sink.sk.yg_rows[r * w + c]; REAL(v)[r + ny * c] = g_sink.sk.vg_rows[r * w + c]; }
    SET_VECTOR_ELT(out, 1, y); SET_VECTOR_ELT(out, 2, v);
    UNPROTECT(2);
  }
  sink_free(h);
  UNPROTECT(2);
  return out;
}
/* gene shards over several R processes (one per GPU): the CUDA IPC blob of this rank's inbox as a raw vector; connect takes the
   blobs of all ranks concatenated in rank order (exchanged by the caller, e.g. through Rmpi or files) */
SEXP zzR_comm_init(SEXP ptr, SEXP rank, SEXP world) {
  zz_handle *h = H(ptr);
  SEXP out = PROTECT(Rf_allocVector(RAWSXP, zz_comm_handle_bytes()));
  int rc = zz_comm_init(h, Rf_asInteger(rank), Rf_asInteger(world), RAW(out));
  UNPROTECT(1); chk(h, rc);
  return out;
}
SEXP zzR_comm_connect(SEXP ptr, SEXP all_handles) { zz_handle *h = H(ptr); chk(h, zz_comm_connect(h, RAW(all_handles))); return R_NilValue; }

static const R_CallMethodDef calls[] = {
  {"zzR_chain_begin", (DL_FUNC)&zzR_chain_begin, 5}, {"zzR_run_generations", (DL_FUNC)&zzR_run_generations, 5},
  {"zzR_chain_read", (DL_FUNC)&zzR_chain_read, 3}, {"zzR_chain_windows", (DL_FUNC)&zzR_chain_windows, 3}, {"zzR_run_sweeps", (DL_FUNC)&zzR_run_sweeps, 4},
  {"zzR_create", (DL_FUNC)&zzR_create, 5}, {"zzR_set_state", (DL_FUNC)&zzR_set_state, 11}, {"zzR_get_state", (DL_FUNC)&zzR_get_state, 2},
  {"zzR_set_hyper", (DL_FUNC)&zzR_set_hyper, 15}, {"zzR_mh_yg", (DL_FUNC)&zzR_mh_yg, 3}, {"zzR_mh_variance_g", (DL_FUNC)&zzR_mh_variance_g, 3},
  {"zzR_gibbs_alloc", (DL_FUNC)&zzR_gibbs_alloc, 2}, {"zzR_gibbs_alloc_within", (DL_FUNC)&zzR_gibbs_alloc_within, 2},
  {"zzR_mh_spike_alloc", (DL_FUNC)&zzR_mh_spike_alloc, 2}, {"zzR_sweep", (DL_FUNC)&zzR_sweep, 3}, {"zzR_suffstats", (DL_FUNC)&zzR_suffstats, 2},
  {"zzR_varg_hyper_sum", (DL_FUNC)&zzR_varg_hyper_sum, 4}, {"zzR_commit_varg_hyper", (DL_FUNC)&zzR_commit_varg_hyper, 1},
  {"zzR_level2_sums", (DL_FUNC)&zzR_level2_sums, 5}, {"zzR_px_delta", (DL_FUNC)&zzR_px_delta, 3}, {"zzR_px_commit", (DL_FUNC)&zzR_px_commit, 3},
  {"zzR_lnl", (DL_FUNC)&zzR_lnl, 1}, {"zzR_tune", (DL_FUNC)&zzR_tune, 2}, {"zzR_nan_flag", (DL_FUNC)&zzR_nan_flag, 1},
  {"zzR_accumulate_alloc_trace", (DL_FUNC)&zzR_accumulate_alloc_trace, 1}, {"zzR_reset_alloc_trace", (DL_FUNC)&zzR_reset_alloc_trace, 1},
  {"zzR_get_alloc_trace", (DL_FUNC)&zzR_get_alloc_trace, 3},
  {"zzR_init_state", (DL_FUNC)&zzR_init_state, 3}, {"zzR_accumulate_alloc_trace_from", (DL_FUNC)&zzR_accumulate_alloc_trace_from, 3},
  {"zzR_copy_chain_tuning", (DL_FUNC)&zzR_copy_chain_tuning, 3},
  {"zzR_sync", (DL_FUNC)&zzR_sync, 1}, {"zzR_launch_count", (DL_FUNC)&zzR_launch_count, 1}, {"zzR_refresh_caches", (DL_FUNC)&zzR_refresh_caches, 1},
  {"zzR_reset_windows", (DL_FUNC)&zzR_reset_windows, 1}, {"zzR_window_sums", (DL_FUNC)&zzR_window_sums, 2}, {"zzR_get_hyper", (DL_FUNC)&zzR_get_hyper, 1},
  {"zzR_eval", (DL_FUNC)&zzR_eval, 3}, {"zzR_philox_draws", (DL_FUNC)&zzR_philox_draws, 5}, {"zzR_sweep_host", (DL_FUNC)&zzR_sweep_host, 12},
  {"zzR_post_predictive", (DL_FUNC)&zzR_post_predictive, 4}, {"zzR_set_sample_sink", (DL_FUNC)&zzR_set_sample_sink, 7},
  {"zzR_sample_rows", (DL_FUNC)&zzR_sample_rows, 1}, {"zzR_comm_init", (DL_FUNC)&zzR_comm_init, 3}, {"zzR_comm_connect", (DL_FUNC)&zzR_comm_connect, 2},
  {NULL, NULL, 0}};

void R_init_zigzag(DllInfo *dll) { R_registerRoutines(dll, NULL, calls, NULL, NULL); R_useDynamicSymbols(dll, FALSE); }
