/*
 * zigzag_gpu.h — C ABI of the B200 (sm_100a) implementation of zigzag's per-gene MCMC sweep.
 *
 * The reference (ammonthompson/zigzag v1.0.0) is pure R and has no FFI; its seam is the method table
 * `proposal_list` (R/zigzagInitialize.R:207-223) whose entries are called as proposal_list[[p]](tune=...)
 * from burnin() (R/zigzagBurnin.R:66-67) and mcmc() (R/zigzagMCMC.R:172-173).  Every entry point below
 * replaces one of those R methods (or a per-gene reduction inside one) and is what an R `.Call` shim
 * (r_overlay/src/zz_shim.c) or the ctypes host (zigzag_b200/_capi.py) binds.
 *
 * Conventions
 *   - plain pointers and sizes only; every `const double*` / `double*` argument is a HOST pointer unless the
 *     name ends in _dev; the library owns all device and pinned memory behind the opaque handle.
 *   - Xg is G x R column-major (R's matrix layout == library-major struct-of-arrays), -Inf = not detected.
 *   - integer vectors follow R: allocation_active_inactive in {0,1}, allocation_within_active in 1..K,
 *     inactive_spike_allocation in {0,1}.
 *   - per-chain arrays are length G (the local gene shard); multi-chain draw/decision tables are [n][C][G].
 *   - return value: ZZ_OK (0) or a negative error code; the message is kept on the handle (zz_last_error).
 *     No exceptions or longjmp cross this boundary.  CUDA errors are sticky.
 *   - NaN acceptance ratios are treated as "reject" and raise the sticky NaN flag (zz_nan_flag); the reference
 *     would poison the state with NA (R/zigzagProposals.R:302-310, guard at R/zigzagBurnin.R:69-78).
 *   - random draws: `draws == NULL` => counter-based Philox4x32-10 keyed by (seed; chain, step, move slot,
 *     global gene index), `draws != NULL` => "identical uniforms" test mode, the caller supplies every draw.
 */
#ifndef ZIGZAG_GPU_H
#define ZIGZAG_GPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZZ_MAX_LIBS 64
#define ZZ_MAX_COMP 8

#define ZZ_OK 0
#define ZZ_ERR_INVALID (-1)
#define ZZ_ERR_CUDA (-2)
#define ZZ_ERR_STATE (-3)
#define ZZ_ERR_NOMEM (-4)

/* move-slot ids used in the Philox counter (and by zz_philox_draws) */
#define ZZ_SLOT_YG 1
#define ZZ_SLOT_VARIANCE_G 2
#define ZZ_SLOT_ALLOC 3
#define ZZ_SLOT_ALLOC_WITHIN 4
#define ZZ_SLOT_SPIKE 5

/* bits of the `moves` mask of zz_sweep */
#define ZZ_MOVE_YG 1u
#define ZZ_MOVE_VARIANCE_G 2u
#define ZZ_MOVE_ALLOC 4u
#define ZZ_MOVE_ALLOC_WITHIN 8u
#define ZZ_MOVE_SPIKE 16u
#define ZZ_MOVE_REFERENCE_ORDER 32u /* zz_sweep: use the reference-operation-order kernel instead of the production one */
#define ZZ_MOVE_FUSED (ZZ_MOVE_YG | ZZ_MOVE_VARIANCE_G | ZZ_MOVE_ALLOC | ZZ_MOVE_SPIKE)

typedef struct zz_handle zz_handle;

/* Scalar model parameters read by the per-gene moves: the RefClass fields s0, s1, tau, alpha_r, spike_probability,
   inactive_means, inactive_variances, active_means, active_variances, weight_active, weight_within_active, beta,
   temperature, variance_g_upper_bound (R/zigzag.R:117-338). */
typedef struct zz_hyper {
  int32_t num_libraries;
  int32_t num_active_components;
  double s0, s1, tau;
  double alpha_r[ZZ_MAX_LIBS];
  double spike_probability;
  double inactive_means, inactive_variances;
  double active_means[ZZ_MAX_COMP], active_variances[ZZ_MAX_COMP];
  double weight_active, weight_within_active[ZZ_MAX_COMP];
  double beta, temperature, variance_g_upper_bound;
} zz_hyper;

typedef struct zz_config {
  int64_t num_genes;   /* genes held by this handle (the local shard) */
  int64_t gene_offset; /* global index of local gene 0: Philox is keyed by the global index, so results do not depend on sharding */
  int32_t num_libraries;
  int32_t num_active_components;
  int32_t num_chains;  /* independent chains batched over the same Xg (config 5) */
  int32_t device;      /* CUDA device ordinal */
  uint64_t seed;
} zz_config;

/* ---- lifetime ------------------------------------------------------------------------------------------- */
int zz_create(const zz_config *cfg, zz_handle **out);           /* replaces: zigzag$new() state allocation, R/zigzagInitialize.R:7-564 */
void zz_destroy(zz_handle *h);
const char *zz_last_error(const zz_handle *h);
int zz_set_stream(zz_handle *h, void *cuda_stream);              /* run on the caller's stream (e.g. torch's current stream) */
int zz_sync(zz_handle *h);
uint64_t zz_launch_count(const zz_handle *h);                    /* kernels launched so far (bench.py gpu_launches) */

/* ---- data and state ------------------------------------------------------------------------------------- */
int zz_upload_data(zz_handle *h, const double *Xg, const double *gene_lengths);  /* fields Xg (I:52-56), gene_lengths (I:143-149) */
/* Any pointer may be NULL (= leave untouched / do not fetch).  Fields: Yg, variance_g, allocation_active_inactive,
   allocation_within_active[[1]], inactive_spike_allocation, XgLikelihood, variance_g_probability, tuningParam_yg,
   tuningParam_variance_g, active_gene_set membership (R/zigzag.R:117-338).  no_detect_spike is derived from Xg. */
int zz_set_state(zz_handle *h, int chain, const double *Yg, const double *variance_g, const int32_t *alloc_active_inactive,
                 const int32_t *alloc_within_active, const int32_t *inactive_spike_allocation, const double *XgLikelihood,
                 const double *variance_g_probability, const double *tuningParam_yg, const double *tuningParam_variance_g,
                 const int32_t *pinned_active);
/* zz_get_state: any pointer may be NULL (that field is skipped).  Destinations may be pageable (R vectors, NumPy arrays) or pinned:
   pageable ones are filled through a pinned bounce region of the handle by ZZ_HOST_COPY_THREADS host threads (default 4). */
int zz_get_state(zz_handle *h, int chain, double *Yg, double *variance_g, int32_t *alloc_active_inactive,
                 int32_t *alloc_within_active, int32_t *inactive_spike_allocation, double *XgLikelihood,
                 double *variance_g_probability, double *tuningParam_yg, double *tuningParam_variance_g);
int zz_set_hyper(zz_handle *h, int chain, const zz_hyper *hy);
int zz_get_hyper(zz_handle *h, int chain, zz_hyper *hy);
/* XgLikelihood <- computeXgLikelihood(Xg,Yg,variance_g,p_x); variance_g_probability <- computeVarianceGPriorProbability (I:473, I:504) */
int zz_refresh_caches(zz_handle *h);
int zz_debug_refresh_lps(zz_handle *h);                          /* diagnostic: rebuild the production sweep's cached detection term A(Yg) now */
int zz_reset_windows(zz_handle *h);                              /* Yg_trace / variance_g_trace <- 77 zeros + 23 ones (I:460,466) */
int zz_get_window_sums(zz_handle *h, int chain, int32_t *yg_sum, int32_t *vg_sum); /* rowSums(Yg_trace), rowSums(variance_g_trace) */

/* ---- evaluators (north-star check 1): one double per gene ------------------------------------------------ */
int zz_eval_xg_loglik(zz_handle *h, int chain, double *out);      /* computeXgLikelihood, R/zigzagProbability.R:57-127 (+ get_px U:150-159) */
int zz_eval_varg_logprior(zz_handle *h, int chain, double *out);  /* computeVarianceGPriorProbability, P:11-31 (+ get_sigmax U:144-148) */
int zz_eval_level2_loglik(zz_handle *h, int chain, double *out);  /* computeInactiveLikelihood / computeActiveLikelihood, P:210-273 */

/* ---- per-gene moves (north-star check 2) ------------------------------------------------------------------ */
/* draws layouts ([n][C][G], may be NULL => Philox):  yg: {z, u}; variance_g: {u_proposal, u_accept}; alloc: {u_ai, u_within};
   alloc_within: {u_within}; spike: {z_y, z_v, u}.  decisions ([C][G] uint8, may be NULL): accept bit, or for the
   allocation moves the new allocation word (0 = inactive, k = active component k). */
int zz_mh_yg(zz_handle *h, uint64_t step, int tune, const double *draws, uint8_t *decisions);            /* mhYg, R/zigzagProposals.R:250-333 */
int zz_mh_variance_g(zz_handle *h, uint64_t step, int tune, const double *draws, uint8_t *decisions);    /* mhVariance_g, Pr:7-56 (+ scale_move MP:3-13) */
int zz_gibbs_alloc(zz_handle *h, uint64_t step, const double *draws, uint8_t *decisions);                /* gibbsAllocationActiveInactive (+ chained WithinActive), Pr:335-410 */
int zz_gibbs_alloc_within(zz_handle *h, uint64_t step, const double *draws, uint8_t *decisions);         /* gibbsAllocationWithinActive, Pr:363-410 */
int zz_mh_spike_alloc(zz_handle *h, uint64_t step, const double *draws, uint8_t *decisions);             /* mhSpikeAllocation, Pr:528-638 */
/* Fused systematic-scan sweep: for every gene mhYg -> mhVariance_g -> allocation (AI + within) -> mhSpikeAllocation with the
   gene's state held in registers.  `moves` selects a subset (ZZ_MOVE_*; 0 = ZZ_MOVE_FUSED).  By default the production
   kernels run (zz_fast.cuh: XgLikelihood split into a cached detection term and a closed-form normal term, table-driven
   exp/log; results agree with the reference-order moves to O(1e-15) relative and decisions can differ only at floating-
   point ties); ZZ_MOVE_REFERENCE_ORDER selects the kernel that keeps the reference's operation order, which is also what
   the five per-move entry points above use.  draws: [9][C][G] = {z_yg,u_yg,u1_vg,u2_vg,u_ai,u_within,zy_spike,zv_spike,
   u_spike} or NULL; decisions: [4][C][G] or NULL. */
int zz_sweep(zz_handle *h, uint64_t step, int tune, uint32_t moves, const double *draws, uint8_t *decisions);
/* Exports the draws Philox would produce for (step, slot) in the layout of that move's `draws` argument. */
int zz_philox_draws(zz_handle *h, uint64_t step, int slot, double *out);

/* ---- reductions that feed the hyper-parameter moves (R/zigzagProposals.R:58-243, 412-526, 640-803) ------- */
int zz_suffstats_len(int num_active_components);                 /* 3(K+1)+10 */
/* out [C][len]: n_c (K+1) | sum y per component (K+1; inactive: out-of-spike only) | sum y^2 (K+1) | n_inspike, n_out,
   sum z, sum z^2, sum y(out), sum y^2(out), sum y z (z = log variance_g, out-of-spike) | sum XgLikelihood | sum variance_g_probability(out) | nan count */
int zz_suffstats(zz_handle *h, double *out);
int zz_suffstats_dev(zz_handle *h, double *out_dev);             /* same, written to caller-owned device memory on the handle's stream (for an NCCL all-reduce) */
/* ---- multi-GPU: the one collective of the path, fused with the final reduction over NVLink peer memory ----
   One process per GPU.  Every rank calls zz_comm_init (allocates its inbox, returns zz_comm_handle_bytes() bytes of CUDA IPC
   handles), the host framework all-gathers the handle blobs (rank order) and passes the concatenation to zz_comm_connect.
   zz_suffstats_allreduce_dev then leaves on every rank the bit-identical sum over ranks (added in rank order) of the
   statistics vectors, in one kernel launch.  A peer that does not show up within ~2 s raises bit 2 of zz_nan_flag. */
int zz_comm_handle_bytes(void);
int zz_comm_init(zz_handle *h, int rank, int world, void *handle_out);
int zz_comm_connect(zz_handle *h, const void *all_handles);
int zz_suffstats_allreduce_dev(zz_handle *h, double *out_dev);
/* Fused epilogue: from now on every whole-shard production zz_sweep also leaves the statistics ([C][len], all-reduced over
   ranks when `allreduce` != 0) in caller-owned device memory `out_dev`, written by the last CTA of the sweep kernel itself:
   one launch per generation.  out_dev == NULL switches it off. */
int zz_set_stats_output_dev(zz_handle *h, double *out_dev, int allreduce);
/* out2 = { sum over out-of-spike genes of computeVarianceGPriorProbability(variance_g, exp(s0+s1*Yg), tau), sum of the cached variance_g_probability } ; Pr:84-87,126-130,167-170 */
int zz_varg_hyper_sum(zz_handle *h, int chain, double s0, double s1, double tau, double *out2);
int zz_commit_varg_hyper(zz_handle *h, int chain);               /* accepted mhTau/mhSg/mhS0Tau: overwrite variance_g_probability[out_spike_idx] (Pr:113,147,187) from the current hyper */
/* out2 = { sum computeInactiveLikelihood over inactive genes, sum computeActiveLikelihood over active genes } at the given parameters; Pr:444-447,485-488,653-654,729-730,780-784 */
int zz_level2_sums(zz_handle *h, int chain, double inactive_mean, double inactive_variance, const double *active_means,
                   const double *active_variances, double *out2);
/* out[r] = sum_g [ll_r(alpha_prop[r]) - ll_r(alpha_r[r])] for every library with lib_mask[r] != 0 (NULL = all); Pr:209-224 */
int zz_px_delta(zz_handle *h, int chain, const double *alpha_prop, const uint8_t *lib_mask, double *out);
int zz_px_commit(zz_handle *h, int chain, int lib, double alpha_new); /* accepted mhP_x: alpha_r[lib] and XgLikelihood update (Pr:213-215 / :219, :234-239) */
int zz_lnl(zz_handle *h, int chain, double *out);                /* calculate_lnl, R/zigzagUtilities.R:70-75 */

/* ---- sampling-time and tuning helpers ---------------------------------------------------------------------- */
int zz_reset_alloc_trace(zz_handle *h);                          /* allocation_trace <- one-hot(current) is reset+accumulate (R/zigzagMCMC.R:142-145) */
int zz_accumulate_alloc_trace(zz_handle *h);                     /* allocation_trace += one-hot(allocation), R/zigzagMCMC.R:194-199 */
int zz_get_alloc_trace(zz_handle *h, int chain, uint32_t *out);  /* [K+1][G] */
/* tempering on the chain axis (mc3, R/zigzagDevelopment.R:9-139): the chain at temperature 1 changes slots when a swap is accepted —
   its samples go to ONE trace (allocation_trace[dst] += one-hot(allocation of chain src)), and the slot that becomes cold inherits the
   cold chain's per-gene tuning parameters (set_tuningParam_fields, R/zigzagDevelopment.R:118) */
int zz_accumulate_alloc_trace_from(zz_handle *h, int src_chain, int dst_chain);
int zz_copy_chain_tuning(zz_handle *h, int src_chain, int dst_chain);
/* Constructor-side initialisation on the device, R/zigzagInitialize.R:416-504, for every chain from its own hyper-parameters
   (zz_set_hyper first): allocations, rwm / rwv from the data, Yg, variance_g, the re-draw loop of genes with zero likelihood, tuning
   parameters 0.5, the 100-wide windows, then both likelihood caches.  rowvar_log_mean / rowvar_sd = log(mean(v)), sd(v) of the row
   variances v of the fully detected genes of the whole data set (I:446).  Philox-keyed by the global gene index: a sharded
   initialisation equals the single-GPU one. */
int zz_init_state(zz_handle *h, uint64_t step, double rowvar_log_mean, double rowvar_sd);
int zz_tune(zz_handle *h, double target);                        /* x_tune on Yg_trace / variance_g_trace, R/zigzagUtilities.R:161-181, 202-206 */
int zz_nan_flag(zz_handle *h, int *flag);                        /* replaces the NA guard R/zigzagBurnin.R:69-78 */

/* ---- host-buffer round trip (bench.py `e2e`) ---------------------------------------------------------------- */
/* One fused sweep where the mutable per-gene state lives in HOST arrays, as it does in the R object between .Call()s:
   takes Yg, variance_g, tuning parameters and allocations from the host arrays, runs zz_sweep, returns the updated state including
   the two caches (XgLikelihood / variance_g_probability are outputs only: they are functions of the uploaded state).  Xg and
   gene_lengths stay resident from zz_upload_data; the handle's resident state ends up equal to the returned arrays, so resident
   calls can follow.  Chain 0 only.  Two transfer paths, same results bit for bit:
     * all nine arrays in pinned, device-mapped host memory (cudaHostAlloc / cudaHostRegister, torch pin_memory()): ONE kernel
       reads the arrays over PCIe, sweeps and writes them back — no staging copies, both PCIe directions busy at once;
     * otherwise: staged cudaMemcpyAsync in chunks of ~256K genes on three streams (copy in / sweep / copy out overlap).
   Bytes moved are returned in h2d_bytes / d2h_bytes (44 bytes per gene each way). */
int zz_sweep_host(zz_handle *h, uint64_t step, int tune, double *Yg, double *variance_g, int32_t *alloc_active_inactive,
                  int32_t *alloc_within_active, int32_t *inactive_spike_allocation, double *XgLikelihood,
                  double *variance_g_probability, const double *tuningParam_yg, const double *tuningParam_variance_g,
                  int64_t *h2d_bytes, int64_t *d2h_bytes);

/* ---- device-resident generation loop ---------------------------------------------------------------- */
/* burnin() (R/zigzagBurnin.R:64-148) and mcmc() (R/zigzagMCMC.R:169-229) call 15 moves per generation from the host; ten of them
   are scalar Metropolis/Gibbs moves on hyper-parameters whose only O(G) part is a sum over genes.  zz_run_generations keeps
   the whole generation on the device: the fused per-gene sweep, then one pass of every hyper move (gibbsMixtureWeights,
   mhInactiveMeans, mhActiveMeansDif, mhSharedVariance, gibbsSpikeProb, mhTau, mhSg, mhS0Tau, mhP_x — Pr:412-431, 433-471,
   640-705, 768-803, 511-526, 119-152, 58-117, 154-191, 193-243) whose proposal, prior ratio, accept/reject and tuning
   window are computed by a single device thread between the gene-sum kernels; accepted moves gate the cache-commit kernels
   (Pr:113/147/187, Pr:213-215) on the device.  Nothing is copied to the host and the host never waits: it only enqueues.
   Scalar draws come from a private Philox stream (gene index 0xFFFFFFFF, block 255, counter hyper_ctr / 4).
   Default model (shared_variance = TRUE, multi_thresh_a = TRUE).  A handle with several chains advances all of them in the same
   launches (one stage block, one reduction slice, one commit gate per chain; every chain draws with its own chain index);
   gene shards on several GPUs (zz_comm_connect) and the sample sink need a one-chain handle. */
typedef struct zz_priors {      /* constructor arguments of zigzag$new, R/zigzagInitialize.R:7-41 */
  double weight_active_shape_1, weight_active_shape_2;
  double spike_prior_shape_1, spike_prior_shape_2;
  double active_means_dif_prior_shape, active_means_dif_prior_rate;
  double inactive_means_prior_shape, inactive_means_prior_rate;
  double variances_prior_log_min, variances_prior_log_max;
  double tau_shape, tau_rate, s0_mu, s0_sigma, s1_shape, s1_rate, alpha_r_shape, alpha_r_rate;
  double threshold_i, threshold_a[ZZ_MAX_COMP];
} zz_priors;

typedef struct zz_chain_scalars {   /* the scalar chain state that lives on the device between zz_chain_begin and zz_chain_read */
  double active_means_dif[ZZ_MAX_COMP];
  double t_im, t_av, t_am[ZZ_MAX_COMP], t_s0, t_s1, t_tau, t_s0tau, t_alpha[ZZ_MAX_LIBS];  /* scalar tuning parameters, I:185-199 */
  uint8_t w_im[100], w_av[100], w_am[ZZ_MAX_COMP][100], w_s0[100], w_s1[100], w_tau[100], w_s0tau[100],
          w_alpha[ZZ_MAX_LIBS][100];   /* their 100-wide accept windows (the *_trace fields) */
  uint64_t hyper_ctr;           /* uniforms consumed from the hyper-level Philox stream */
  double hyper_buf[4];          /* the current block of that stream */
  uint64_t step;                /* move counter: Philox `step` of the next per-gene sweep (advances by 10 per generation) */
  int64_t gen;                  /* generations done */
  int64_t n_samples;            /* allocation_trace accumulations done */
} zz_chain_scalars;

int zz_chain_begin(zz_handle *h, const zz_priors *priors, const zz_chain_scalars *scalars);   /* scalars: one struct per chain (gen, step, n_samples equal) */
/* ngen generations; tune != 0: burn-in (acceptance windows, tune_all every 400 generations with `target`, U:183-240);
   tune == 0: sampling, allocation_trace accumulated whenever gen % sample_frequency == 0 (M:188-199).  Asynchronous. */
int zz_run_generations(zz_handle *h, int ngen, int tune, int sample_frequency, double target);
/* `nsweeps` fused sweeps with the hyper-parameters held fixed, steps step, step + 1, ..., in ONE cooperative launch of the
   device-resident kernel (no hyper-parameter moves; zz_chain_begin is not needed).  After every sweep the sufficient statistics
   — all-reduced over the ranks once zz_comm_connect is done — are left in the buffer of zz_set_stats_output_dev, if any
   (zz_suffstats layout, except that entries 3(K+1)+7 and +8, the sums of the two likelihood caches, are 0 here).  Asynchronous. */
int zz_run_sweeps(zz_handle *h, uint64_t step, int nsweeps, int tune);
int zz_debug_gen_times(zz_handle *h, double *out8);   /* diagnostic: ns the releasing CTA spent in the last device-resident launch {ticket -> release, scalar stage, 6 phases of the stage} */
/* waits for the device, returns the scalar state and the hyper-parameters of every chain ([num_chains] each; either may be NULL) */
int zz_chain_read(zz_handle *h, zz_chain_scalars *scalars, zz_hyper *hyper);

/* ---- sample-time outputs without stopping the device loop (R/zigzagMCMC.R:188-220, R/zigzagFileWriting.R:60-109) ----
   While zz_run_generations samples (tune == 0, gen % sample_frequency == 0) the device itself appends one row per sample to
   rings of pinned host memory (zero-copy stores over PCIe; the host never waits):
     param_rows[row] = { gen, lnl (calculate_lnl, U:70-75), s0, s1, tau, alpha_r[R], spike_probability, inactive_means,
                         inactive_variances, active_means[K], active_variances[K], weight_active, weight_within_active[K] }
                       — the columns of <prefix>_model_parameters.log, zz_param_row_len(R, K) doubles;
     yg_rows[row] / vg_rows[row] = { gen, value of every candidate gene } as <prefix>_yg_candidate_genes.log /
                       _varianceg_candidate_genes.log print them (in-spike genes: -10 / 0), one row whenever
                       (gen / sample_frequency) % yv_every == 0, until yv_capacity rows exist (the reference stops after 500).
   All row buffers must come from zz_pinned_alloc (or any pinned, device-mapped allocation).  The writer that formats the log
   files stays on the host and reads the rows after zz_sync / zz_chain_read (or concurrently, rows complete in order). */
typedef struct zz_sample_sink {
  double *param_rows; int64_t param_capacity;          /* [param_capacity][zz_param_row_len(R, K)] */
  double *yg_rows, *vg_rows; int64_t yv_capacity;      /* [yv_capacity][1 + n_candidates] each; NULL: no per-gene rows */
  const int32_t *candidates; int64_t n_candidates;     /* caller's (local) gene indices, host array, copied */
  int32_t yv_every;
} zz_sample_sink;
int zz_param_row_len(int num_libraries, int num_active_components);
void *zz_pinned_alloc(size_t bytes);                   /* cudaHostAlloc (portable, mapped); NULL on failure */
void zz_pinned_free(void *p);
int zz_set_sample_sink(zz_handle *h, const zz_sample_sink *sink);   /* NULL detaches; row counters restart at 0 */
int zz_sample_rows_written(zz_handle *h, int64_t *param_rows, int64_t *yv_rows);   /* rows enqueued so far */

/* ---- posterior-predictive discrepancies (R/zigzagPostPredictiveSimulation.R:3-156) ----------------------------
   Simulates one data set from the current state (x_sim ~ N(Yg, variance_g), dropped with probability 1 - p_x, all -Inf in the
   spike; Yg_sim from the gene's mixture component; Philox blocks 16 + r and 15 of `step`) and returns the realised discrepancy
   statistics the reference appends to <prefix>.post_predictive_output.log and .library_specific_post_predictive_output.log:
     out = { W_L, W_U, B_rumsfeld[1], W_L_r[R], B_rumsfeld[2..R+1] }   (3 + 2 R doubles)
   `bins` = seq(min(ranges) - 2, max(ranges) + 2, by = 0.1) as posteriorPredictiveSimulation builds it (PostPred:9-17).
   Chain 0; synchronous (the histograms come back to the host, where the per-bin functionals are formed). */
int zz_post_predictive(zz_handle *h, uint64_t step, const double *bins, int nbins, double *out);
/* The same on gene shards: `allreduce(ctx, buf, n, dtype)` sums a HOST buffer of n elements in place over the shards (dtype: 0 =
   double, 1 = uint32, 2 = uint64); it is called four times per data set (library sums + gene count, detected-value counts, all
   histograms, the fixed-point model weights).  NULL: one shard holds every gene (zz_post_predictive). */
#define ZZ_DTYPE_F64 0
#define ZZ_DTYPE_U32 1
#define ZZ_DTYPE_U64 2
typedef void (*zz_allreduce_fn)(void *ctx, void *buf, int64_t n, int dtype);
int zz_post_predictive_sharded(zz_handle *h, uint64_t step, const double *bins, int nbins, zz_allreduce_fn allreduce, void *ctx, double *out);

#ifdef __cplusplus
}
#endif
#endif
