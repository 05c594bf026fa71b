# zigzagGPU.R — GPU drop-in for the zigzag R package (add this file as R/zigzagGPU.R, r_overlay/src/zz_shim.c as src/zz_shim.c,
# `useDynLib(zigzag, .registration = TRUE)` to NAMESPACE and link -lzigzag_gpu).
#
# `zigzagGPU` is a subclass of the reference class `zigzag` (R/zigzag.R:113): zigzagGPU$new(...) takes the same arguments,
# burnin() / mcmc() / the output files / mcmc_report are the reference's own code.  What changes:
#   * the five per-gene entries of `proposal_list` (R/zigzagInitialize.R:207-221) are kernel launches on device-resident state;
#   * the ten hyper-parameter moves are overridden below: same proposals, priors, accept rule and trace bookkeeping as
#     R/zigzagProposals.R, but every sum over genes comes from the device (the R copies of Yg, variance_g, the allocations and the
#     two likelihood caches are NOT current between sampling points, so nothing may read them inside a move);
#   * at every sampling point of burnin() (R/zigzagBurnin.R:88-131) and mcmc() (R/zigzagMCMC.R:188-217) the per-gene fields are
#     refreshed from the device BEFORE the reference's sample block reads them: each move ends with gpu_after_move(), which pulls
#     when the current generation is a sampling generation and a per-gene move has run since the last pull;
#   * tune_all (R/zigzagUtilities.R:183-240): scalar parameters by the reference's code, per-gene ones on the device (zz_tune).
# Only the default model (shared_variance = TRUE, shared_active_variance = TRUE) is covered: the weight-0 moves
# mhInactiveVariances / mhActiveVariances are not overridden.
# Every .Call below is registered in src/zz_shim.c; tests/test_r_overlay.py checks names and argument counts of both files.

zigzagGPU <- setRefClass(
  "zigzagGPU",
  contains = "zigzag",
  fields = list(gpu = "ANY", gpu_step = "numeric", gpu_dirty = "logical", gpu_sample_frequency = "numeric",
                gpu_in_burnin = "logical", gpu_hyper_ctr = "numeric", gpu_hyper_buf = "numeric", gpu_n_samples = "numeric"),
  methods = list(

  initialize = function(..., gpu_seed = 1, gpu_device = 0){
    callSuper(...)
    if(!shared_variance || !shared_active_variance) stop("zigzagGPU covers the default shared-variance model only")
    .self$gpu_attach(gpu_seed, gpu_device)
  },

  gpu_attach = function(seed = 1, device = 0){
    .self$gpu <- .Call("zzR_create", Xg, gene_lengths, as.integer(num_active_components), as.numeric(seed), as.integer(device))
    .self$gpu_step <- 0; .self$gpu_dirty <- FALSE; .self$gpu_sample_frequency <- Inf; .self$gpu_in_burnin <- FALSE
    .self$gpu_hyper_ctr <- 0; .self$gpu_hyper_buf <- numeric(4); .self$gpu_n_samples <- 0
    pinned <- NULL
    if(length(active_gene_set_idx) > 0){ pinned <- integer(num_transcripts); pinned[active_gene_set_idx] <- 1L }
    .Call("zzR_set_state", gpu, Yg, variance_g, as.integer(allocation_active_inactive), as.integer(allocation_within_active[[1]]),
          as.integer(inactive_spike_allocation), XgLikelihood, variance_g_probability, tuningParam_yg, tuningParam_variance_g, pinned)
    .self$gpu_push_hyper()
    # the per-gene table entries (same signature: function(recover_x = FALSE, tune = FALSE))
    .self$proposal_list[[2]]  <- .self$gpu_gibbsAllocationActiveInactive
    .self$proposal_list[[3]]  <- .self$gpu_gibbsAllocationWithinActive
    .self$proposal_list[[9]]  <- .self$gpu_mhSpikeAllocation
    .self$proposal_list[[10]] <- .self$gpu_mhYg
    .self$proposal_list[[11]] <- .self$gpu_mhVariance_g
  },

  gpu_push_hyper = function(){
    .Call("zzR_set_hyper", gpu, s0, s1, tau, alpha_r, spike_probability, inactive_means, inactive_variances, active_means,
          active_variances, weight_active, weight_within_active, beta, temperature, variance_g_upper_bound)
  },

  gpu_next = function(){ .self$gpu_step <- gpu_step + 1; gpu_step },

  # device -> fields (everything the reference's sample blocks, file writers and reports read)
  gpu_pull = function(){
    st <- .Call("zzR_get_state", gpu, num_transcripts)
    .self$Yg <- st$Yg; .self$variance_g <- st$variance_g
    .self$allocation_active_inactive <- st$allocation_active_inactive
    .self$allocation_within_active[[1]] <- st$allocation_within_active
    .self$inactive_spike_allocation <- as.numeric(st$inactive_spike_allocation)
    .self$XgLikelihood <- st$XgLikelihood; .self$variance_g_probability <- st$variance_g_probability
    .self$tuningParam_yg <- st$tuningParam_yg; .self$tuningParam_variance_g <- st$tuningParam_variance_g
    .self$setActiveInactive_idx(); .self$setInSpike_idx(); .self$set_sigmaX_pX()
    .self$gpu_dirty <- FALSE
  },

  # called at the end of EVERY move: the last move of a sampling generation leaves current fields behind for the sample block
  # (burnin samples when i %% sample_frequency == 0 & i > 2 * sample_frequency with i == gen, B:88; mcmc when gen %% sample_frequency == 0, M:188)
  gpu_after_move = function(){
    if(gpu_dirty && is.finite(gpu_sample_frequency) && gen %% gpu_sample_frequency == 0 &&
       (!gpu_in_burnin || gen > 2 * gpu_sample_frequency)) .self$gpu_pull()
    if(.Call("zzR_nan_flag", gpu)) .self$Yg[1] <- NA          # trips the reference's NA guard (B:69-78)
  },

  burnin = function(sample_frequency = 100, ...){
    .self$gpu_sample_frequency <- sample_frequency; .self$gpu_in_burnin <- TRUE
    on.exit({ .self$gpu_in_burnin <- FALSE; .self$gpu_pull() })
    callSuper(sample_frequency = sample_frequency, ...)
  },

  mcmc = function(sample_frequency = 100, ...){
    .self$gpu_sample_frequency <- sample_frequency; .self$gpu_in_burnin <- FALSE
    .self$gpu_pull()                                            # M:142-145 starts allocation_trace from the current allocation
    on.exit(.self$gpu_pull())
    callSuper(sample_frequency = sample_frequency, ...)
  },

  # ---- the five per-gene moves --------------------------------------------------------------------------------------------
  gpu_mhYg = function(recover_x = FALSE, tune = FALSE){
    .Call("zzR_mh_yg", gpu, gpu_next(), tune); .self$gpu_dirty <- TRUE; .self$gpu_after_move() },
  gpu_mhVariance_g = function(recover_x = FALSE, tune = FALSE){
    .Call("zzR_mh_variance_g", gpu, gpu_next(), tune); .self$gpu_dirty <- TRUE; .self$gpu_after_move() },
  gpu_gibbsAllocationActiveInactive = function(recover_x = FALSE, tune = FALSE){
    .Call("zzR_gibbs_alloc", gpu, gpu_next()); .self$gpu_dirty <- TRUE; .self$gpu_after_move() },
  gpu_gibbsAllocationWithinActive = function(recover_x = FALSE, tune = FALSE){
    .Call("zzR_gibbs_alloc_within", gpu, gpu_next()); .self$gpu_dirty <- TRUE; .self$gpu_after_move() },
  gpu_mhSpikeAllocation = function(recover_x = FALSE, tune = FALSE){
    .Call("zzR_mh_spike_alloc", gpu, gpu_next()); .self$gpu_dirty <- TRUE; .self$gpu_after_move() },

  # ---- hyper-parameter moves: the moves of R/zigzagProposals.R with every gene sum taken from the device ------------------------
  # Shared pieces.  A trace is the 100-wide 0/1 accept window of a move: `*_trace[[1]][[1]]`, a vector, or one row of a matrix.
  gpu_win = function(name, row = NA, hit = FALSE){
    tr <- .self$field(name)
    w <- tr[[1]][[1]]
    if(is.na(row)){ if(hit) w[100] <- 1 else w <- c(w[-1], 0) }
    else          { if(hit) w[row, 100] <- 1 else w[row, ] <- c(w[row, -1], 0) }
    tr[[1]][[1]] <- w
    .self$field(name, tr)
  },
  gpu_mh_accept = function(log_ratio) log(runif(1)) < log_ratio,
  # counts per component and in the spike (zz_suffstats, layout in include/zigzag_gpu.h)
  gpu_counts = function(){
    st <- .Call("zzR_suffstats", gpu, as.integer(num_active_components))
    K1 <- num_active_components + 1
    list(n = st[seq(K1)], n_inspike = st[3 * K1 + 1])
  },
  # c(sum over inactive genes of computeInactiveLikelihood, sum over active genes of computeActiveLikelihood) at the given parameters
  gpu_l2 = function(im = inactive_means, iv = inactive_variances, am = active_means, av = active_variances)
    .Call("zzR_level2_sums", gpu, im, iv, am, av),
  # after an accepted move on (s0, s1, tau): constants to the device, variance_g_probability[out_spike_idx] rewritten there (Pr:113,147,187)
  gpu_commit_varg = function(){ .self$gpu_push_hyper(); .Call("zzR_commit_varg_hyper", gpu); .self$gpu_dirty <- TRUE },

  gibbsMixtureWeights = function(recover_x = FALSE, tune = FALSE){                       # Pr:412-431
    w <- r_dirichlet(c(weight_active_shape_2, weight_within_active_alpha) + gpu_counts()$n * beta)
    if(!is.nan(w[1])){
      .self$weight_active <- 1 - w[1]
      .self$weight_within_active <- w[-1] / sum(w[-1])
      gpu_push_hyper()
    }
    gpu_after_move()
  },

  gibbsSpikeProb = function(recover_x = FALSE, tune = FALSE){                            # Pr:511-526
    cn <- gpu_counts()
    p <- rbeta(1, cn$n_inspike + spike_prior_shape_1, (cn$n[1] - cn$n_inspike) + spike_prior_shape_2)
    if(!is.nan(p)){ .self$spike_probability <- p; gpu_push_hyper() }
    gpu_after_move()
  },

  mhInactiveMeans = function(recover_x = FALSE, tune = FALSE){                           # Pr:433-471
    .self$inactive_means_proposed <- threshold_i - abs(threshold_i - inactive_means - rnorm(1, 0, inactive_mean_tuningParam[[1]]))
    dL <- gpu_l2(im = inactive_means_proposed)[1] - gpu_l2()[1]          # zero when no gene is inactive, as in the reference
    dP <- computeInactiveMeansPriorProbability(inactive_means_proposed) - computeInactiveMeansPriorProbability(inactive_means)
    if(tune) gpu_win("inactive_means_trace")
    if(gpu_mh_accept(temperature * (dL + dP))){
      .self$inactive_means <- inactive_means_proposed
      if(tune) gpu_win("inactive_means_trace", hit = TRUE)
      gpu_push_hyper()
    }
    gpu_after_move()
  },

  mhActiveMeansDif = function(recover_x = FALSE, tune = FALSE){                          # Pr:640-705 (both branches: the sums are 0 without active genes)
    for(k in sample(num_active_components, num_active_components, replace = TRUE)){
      dif <- active_means_dif
      dif[k] <- abs(dif[k] + rnorm(1, 0, active_mean_tuningParam[k]))
      .self$active_means_dif_proposed <- dif
      means_p <- calculate_active_means(dif)
      Lp <- gpu_l2(am = means_p)[2]
      Lc <- gpu_l2()[2]
      pp <- computeActiveMeansDifPriorProbability(dif[k])
      pc <- computeActiveMeansDifPriorProbability(active_means_dif[k])
      if(tune) gpu_win("active_means_trace", row = k)
      if(gpu_mh_accept(temperature * (Lp + pp - Lc - pc))){
        .self$active_means_dif[k] <- dif[k]
        .self$active_means <- means_p
        if(tune) gpu_win("active_means_trace", row = k, hit = TRUE)
        gpu_push_hyper()
      }
    }
    gpu_after_move()
  },

  mhSharedVariance = function(recover_x = FALSE, tune = FALSE){                          # Pr:768-803
    v <- 10^two_boundary_slide_move(log(active_variances[1], 10), active_variances_prior_log_min, active_variances_prior_log_max,
                                    active_variance_tuningParam[1])
    .self$active_variances_proposed <- rep(v, num_active_components)
    if(tune) gpu_win("active_variances_trace", row = 1)
    Lp <- gpu_l2(iv = v, av = active_variances_proposed)                 # c(inactive, active) at the proposed shared variance
    Lc <- gpu_l2()
    dP <- computeActiveVariancesPriorProbability(v) - computeActiveVariancesPriorProbability(active_variances[1])
    if(gpu_mh_accept(temperature * (sum(Lp) - sum(Lc) + dP))){
      .self$active_variances <- active_variances_proposed
      .self$inactive_variances <- v
      if(tune) gpu_win("active_variances_trace", row = 1, hit = TRUE)
      gpu_push_hyper()
    }
    gpu_after_move()
  },

  mhTau = function(recover_x = FALSE, tune = FALSE){                                     # Pr:119-152
    mv <- scale_move(tau, tuningParam_tau)                               # list(proposed value, scaling factor)
    L <- .Call("zzR_varg_hyper_sum", gpu, s0, s1, mv[[1]])               # c(sum at the proposal, sum at the current values) over out_spike_idx
    dP <- computeTauPriorProbability(mv[[1]]) - computeTauPriorProbability(tau)
    if(tune) gpu_win("tau_trace")
    if(gpu_mh_accept(temperature * (L[1] - L[2] + dP) + log(mv[[2]]))){
      .self$tau <- mv[[1]]
      if(tune) gpu_win("tau_trace", hit = TRUE)
      gpu_commit_varg()
    }
    gpu_after_move()
  },

  mhSg = function(recover_x = FALSE, tune = FALSE){                                      # Pr:58-117
    on_s0 <- sample(seq(2), 1, prob = c(1, 1)) == 1
    new_s0 <- s0; new_s1 <- s1; log_hastings <- 0
    if(on_s0) new_s0 <- s0 + rnorm(1, 0, tuningParam_s0)
    else { mv <- scale_move(s1, tuningParam_s1); new_s1 <- mv[[1]]; log_hastings <- log(mv[[2]]) }
    trace <- if(on_s0) "s0_trace" else "s1_trace"
    if(tune) gpu_win(trace)
    L <- .Call("zzR_varg_hyper_sum", gpu, new_s0, new_s1, tau)
    dP <- computeS0PriorProbability(new_s0) + computeS1PriorProbability(new_s1) - computeS0PriorProbability(s0) - computeS1PriorProbability(s1)
    if(gpu_mh_accept(temperature * (L[1] - L[2] + dP) + log_hastings)){
      .self$s0 <- new_s0; .self$s1 <- new_s1
      if(tune) gpu_win(trace, hit = TRUE)
      gpu_commit_varg()
    }
    gpu_after_move()
  },

  mhS0Tau = function(recover_x = FALSE, tune = FALSE){                                   # Pr:154-191: tau scales, s0 compensates
    mv <- scale_move(tau, tuningParam_s0tau)
    log_sf <- log(mv[[2]])
    new_s0 <- s0 - log_sf
    if(tune) gpu_win("s0tau_trace")
    L <- .Call("zzR_varg_hyper_sum", gpu, new_s0, s1, mv[[1]])
    dP <- computeS0PriorProbability(new_s0) + computeTauPriorProbability(mv[[1]]) - computeS0PriorProbability(s0) - computeTauPriorProbability(tau)
    if(gpu_mh_accept(temperature * (L[1] - L[2] + dP) + log_sf)){
      .self$s0 <- new_s0; .self$tau <- mv[[1]]
      if(tune) gpu_win("s0tau_trace", hit = TRUE)
      gpu_commit_varg()
    }
    gpu_after_move()
  },

  mhP_x = function(recover_x = FALSE, tune = FALSE){                                     # Pr:193-243
    lib <- sample(num_libraries, 1)
    mv <- scale_move(alpha_r[lib], tuningParam_alpha_r[lib])
    alpha_p <- alpha_r; alpha_p[lib] <- mv[[1]]
    if(tune) gpu_win("alpha_r_trace", row = lib)
    # Lp - Lc: the sum over genes of the one library's term at alpha' minus the same at alpha (for num_libraries == 2 the reference
    # recomputes the whole likelihood; the other library's terms cancel in the difference)
    dL <- .Call("zzR_px_delta", gpu, alpha_p, as.integer(lib))[lib]
    log_ratio <- temperature * (dL + computeAlphaRPriorProbability(alpha_p) - computeAlphaRPriorProbability(alpha_r)) + log(mv[[2]])
    if(gpu_mh_accept(log_ratio) & log_ratio != Inf){
      .self$alpha_r[lib] <- mv[[1]]
      if(tune) gpu_win("alpha_r_trace", row = lib, hit = TRUE)
      .Call("zzR_px_commit", gpu, as.integer(lib), mv[[1]])             # alpha_r, p_x and XgLikelihood on the device
      .self$gpu_dirty <- TRUE
    }
    gpu_after_move()
  },

  calculate_lnl = function(num_libs) .Call("zzR_lnl", gpu),                               # U:70-75

  tune_all = function(burnin_target_acceptance_rate){                                    # U:183-240
    callSuper(burnin_target_acceptance_rate)        # scalar parameters: the reference's own code on the R-side traces
    .Call("zzR_tune", gpu, burnin_target_acceptance_rate)     # tuningParam_yg / tuningParam_variance_g from the device-side windows (U:202-206)
    .self$gpu_dirty <- TRUE                          # the per-gene tuning parameters the super call derived from stale traces are replaced at the next pull
  },

  # re-draws the per-gene state on the device by the constructor's recipe (R/zigzagInitialize.R:416-504) from the current hyper-parameters
  gpu_init_state = function(){
    rv <- matrixStats::rowVars(Xg); rv <- rv[is.finite(rv)]
    .Call("zzR_init_state", gpu, 0, c(log(mean(rv)), sqrt(var(rv))))
    .self$gpu_dirty <- TRUE; .self$gpu_pull()
  },

  # posterior-predictive discrepancies of one simulated data set, computed on the device (R/zigzagPostPredictiveSimulation.R:3-156)
  gpu_post_predictive = function(bins){
    out <- .Call("zzR_post_predictive", gpu, gpu_next(), as.numeric(bins), as.integer(num_libraries))
    R <- num_libraries
    list(W_L = out[1], W_U = out[2], B = out[3], W_L_r = out[3 + seq(R)], B_r = out[3 + R + seq(R)])
  },

  # ---- device-resident segments: n generations (fused sweep + every hyper move + tune_all + allocation trace) without returning
  # to R between moves (zz_run_generations).  The scalar state travels once per segment; per-gene fields through gpu_pull().
  gpu_run_generations = function(n, tune = FALSE, sample_frequency = 1, target = 0.44){
    pri <- c(weight_active_shape_1, weight_active_shape_2, spike_prior_shape_1, spike_prior_shape_2,
             active_means_dif_prior_shape, active_means_dif_prior_rate, inactive_means_prior_shape, inactive_means_prior_rate,
             active_variances_prior_log_min, active_variances_prior_log_max, tau_shape, tau_rate, s0_mu, s0_sigma, s1_shape, s1_rate,
             alpha_r_shape, alpha_r_rate, threshold_i, threshold_a)
    K <- num_active_components; R <- num_libraries
    sc <- list(active_means_dif = active_means_dif, t_im = inactive_mean_tuningParam[[1]], t_av = active_variance_tuningParam[1],
               t_am = active_mean_tuningParam, t_s0 = tuningParam_s0, t_s1 = tuningParam_s1, t_tau = tuningParam_tau,
               t_s0tau = tuningParam_s0tau, t_alpha = tuningParam_alpha_r,
               w_im = as.integer(inactive_means_trace[[1]][[1]]), w_av = as.integer(active_variances_trace[[1]][[1]][1,]),
               w_am = lapply(seq(K), function(k) as.integer(active_means_trace[[1]][[1]][k,])),
               w_s0 = as.integer(s0_trace[[1]][[1]]), w_s1 = as.integer(s1_trace[[1]][[1]]), w_tau = as.integer(tau_trace[[1]][[1]]),
               w_s0tau = as.integer(s0tau_trace[[1]][[1]]),
               w_alpha = lapply(seq(R), function(r) as.integer(alpha_r_trace[[1]][[1]][r,])),
               hyper_ctr = gpu_hyper_ctr, hyper_buf = gpu_hyper_buf, step = gpu_step + 1, gen = gen, n_samples = gpu_n_samples)
    .Call("zzR_chain_begin", gpu, pri, sc, as.integer(K), as.integer(R))
    .Call("zzR_run_generations", gpu, as.integer(n), tune, as.integer(sample_frequency), target)
    out <- .Call("zzR_chain_read", gpu, as.integer(K), as.integer(R))
    .self$s0 <- out[1]; .self$s1 <- out[2]; .self$tau <- out[3]; .self$spike_probability <- out[4]
    .self$inactive_means <- out[5]; .self$inactive_variances <- out[6]; .self$weight_active <- out[7]
    .self$gen <- out[8]; .self$gpu_step <- out[9] - 1; .self$gpu_n_samples <- out[10]; .self$gpu_hyper_ctr <- out[11]
    j <- 11
    .self$alpha_r <- out[j + seq(R)]; j <- j + R
    .self$active_means <- out[j + seq(K)]; j <- j + K
    .self$active_variances <- out[j + seq(K)]; j <- j + K
    .self$weight_within_active <- out[j + seq(K)]; j <- j + K
    .self$active_means_dif <- out[j + seq(K)]; j <- j + K
    .self$inactive_mean_tuningParam <- out[j + 1]; .self$active_variance_tuningParam <- rep(out[j + 2], K)
    .self$tuningParam_s0 <- out[j + 3]; .self$tuningParam_s1 <- out[j + 4]; .self$tuningParam_tau <- out[j + 5]
    .self$tuningParam_s0tau <- out[j + 6]; j <- j + 6
    .self$active_mean_tuningParam <- out[j + seq(K)]; j <- j + K
    .self$tuningParam_alpha_r <- out[j + seq(R)]
    w <- .Call("zzR_chain_windows", gpu, as.integer(K), as.integer(R))                   # the *_trace windows of the scalar moves
    .self$inactive_means_trace[[1]][[1]] <- w$w_im; .self$active_variances_trace[[1]][[1]][1,] <- w$w_av
    .self$s0_trace[[1]][[1]] <- w$w_s0; .self$s1_trace[[1]][[1]] <- w$w_s1; .self$tau_trace[[1]][[1]] <- w$w_tau
    .self$s0tau_trace[[1]][[1]] <- w$w_s0tau
    .self$active_means_trace[[1]][[1]][seq(K),] <- w$w_am; .self$alpha_r_trace[[1]][[1]][seq(R),] <- w$w_alpha
    .self$gpu_dirty <- TRUE
  },

  # the allocation counts the device accumulated during device-resident sampling segments, in the shape of the field (M:194-199)
  gpu_allocation_trace = function() .Call("zzR_get_alloc_trace", gpu, num_transcripts, as.integer(num_active_components))
))
