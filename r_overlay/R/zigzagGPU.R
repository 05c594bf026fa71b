# zigzagGPU.R — overlay for the zigzag R package: after `gpu_attach()` the five per-gene entries of `proposal_list`
# (R/zigzagInitialize.R:207-221) and the O(G) sums inside the hyper-parameter moves run on the GPU through src/zz_shim.c.
# zigzag$new()/burnin()/mcmc(), the output files and mcmc_report are untouched: they read the same fields, which
# `gpu_pull()` refreshes from the device at sample time.  Needs `useDynLib(zigzag, .registration = TRUE)` in NAMESPACE.

zigzag$methods(

  gpu_attach = function(seed = 1, device = 0){
    .self$gpu <- .Call("zzR_create", Xg, gene_lengths, as.integer(num_active_components), as.numeric(seed), as.integer(device))
    .self$gpu_step <- 0
    pinned <- NULL
    if(!is.null(active_gene_set)){ pinned <- integer(num_transcripts); pinned[active_gene_set_idx] <- 1L }
    .Call("zzR_set_state", gpu, Yg, variance_g, as.integer(allocation_active_inactive), as.integer(allocation_within_active[[1]]),
          as.integer(inactive_spike_allocation), XgLikelihood, variance_g_probability, tuningParam_yg, tuningParam_variance_g, pinned)
    .self$gpu_push_hyper()
    # swap the per-gene table entries (same signature: function(recover_x = FALSE, tune = FALSE))
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

  # device -> fields; called where the drivers read per-gene state (R/zigzagBurnin.R:116-131, R/zigzagMCMC.R:194-217)
  gpu_pull = function(){
    st <- .Call("zzR_get_state", gpu, num_transcripts)
    .self$Yg <- st$Yg; .self$variance_g <- st$variance_g
    .self$allocation_active_inactive <- st$allocation_active_inactive
    .self$allocation_within_active[[1]] <- st$allocation_within_active
    .self$inactive_spike_allocation <- as.numeric(st$inactive_spike_allocation)
    .self$XgLikelihood <- st$XgLikelihood; .self$variance_g_probability <- st$variance_g_probability
    .self$tuningParam_yg <- st$tuningParam_yg; .self$tuningParam_variance_g <- st$tuningParam_variance_g
    .self$setActiveInactive_idx(); .self$setInSpike_idx(); .self$set_sigmaX_pX()
  },

  gpu_mhYg = function(recover_x = FALSE, tune = FALSE) .Call("zzR_mh_yg", gpu, gpu_next(), tune),
  gpu_mhVariance_g = function(recover_x = FALSE, tune = FALSE) .Call("zzR_mh_variance_g", gpu, gpu_next(), tune),
  gpu_gibbsAllocationActiveInactive = function(recover_x = FALSE, tune = FALSE) .Call("zzR_gibbs_alloc", gpu, gpu_next()),
  gpu_gibbsAllocationWithinActive = function(recover_x = FALSE, tune = FALSE) .Call("zzR_gibbs_alloc_within", gpu, gpu_next()),
  gpu_mhSpikeAllocation = function(recover_x = FALSE, tune = FALSE) .Call("zzR_mh_spike_alloc", gpu, gpu_next()),

  # example of a hyper-parameter move whose sums come from the device (R/zigzagProposals.R:119-152); the others follow suit
  mhTau = function(recover_x = FALSE, tune = FALSE){
    proposal <- .self$scale_move(tau, tuningParam_tau)
    proposed_tau <- proposal[[1]]
    HR <- log(proposal[[2]])
    L <- .Call("zzR_varg_hyper_sum", gpu, s0, s1, proposed_tau)          # c(Lp, Lc)
    pp <- .self$computeTauPriorProbability(proposed_tau)
    pc <- .self$computeTauPriorProbability(tau)
    R <- temperature * (L[1] - L[2] + pp - pc) + HR
    if(tune) .self$tau_trace[[1]][[1]] <- c(tau_trace[[1]][[1]][-1], 0)
    if(log(runif(1)) < R){
      .self$tau <- proposed_tau
      if(tune) .self$tau_trace[[1]][[1]][100] <- 1
      .self$gpu_push_hyper()
      .Call("zzR_commit_varg_hyper", gpu)
    }
  },

  calculate_lnl = function(num_libs) .Call("zzR_lnl", gpu),

  # ---- device-resident segments: n generations (fused sweep + every hyper move + tune_all + allocation trace) without returning
  # to R between moves (zz_run_generations).  burnin()/mcmc() call this between two sampling points instead of looping over
  # proposal_list; the fields the hyper moves mutate travel once per segment.
  gpu_run_generations = function(n, tune = FALSE, sample_frequency = 1, target = 0.44){
    pri <- c(weight_active_shape_1, weight_active_shape_2, spike_prior_shape_1, spike_prior_shape_2,
             active_means_dif_prior_shape, active_means_dif_prior_rate, inactive_means_prior_shape, inactive_means_prior_rate,
             shared_variance_prior_log_min, shared_variance_prior_log_max, tau_shape, tau_rate, s0_mu, s0_sigma, s1_shape, s1_rate,
             alpha_r_shape, alpha_r_rate, threshold_i, threshold_a)
    K <- num_active_components; R <- num_libraries
    sc <- list(active_means_dif = active_means_dif, t_im = inactive_mean_tuningParam, t_av = active_variance_tuningParam[1],
               t_am = active_mean_tuningParam, t_s0 = tuningParam_s0, t_s1 = tuningParam_s1, t_tau = tuningParam_tau,
               t_s0tau = tuningParam_s0tau, t_alpha = tuningParam_alpha_r,
               w_im = as.integer(inactive_means_trace[[1]][[1]]), w_av = as.integer(active_variances_trace[[1]][[1]]),
               w_am = lapply(seq(K), function(k) as.integer(active_means_trace[[k]][[1]])),
               w_s0 = as.integer(s0_trace[[1]][[1]]), w_s1 = as.integer(s1_trace[[1]][[1]]), w_tau = as.integer(tau_trace[[1]][[1]]),
               w_s0tau = as.integer(s0tau_trace[[1]][[1]]),
               w_alpha = lapply(seq(R), function(r) as.integer(alpha_r_trace[[r]][[1]])),
               hyper_ctr = gpu_hyper_ctr, hyper_buf = gpu_hyper_buf, step = gpu_step + 1, gen = gen, n_samples = gpu_n_samples)
    .Call("zzR_chain_begin", gpu, pri, sc, K, R)
    .Call("zzR_run_generations", gpu, as.integer(n), tune, as.integer(sample_frequency), target)
    out <- .Call("zzR_chain_read", gpu, K, R)
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
    # (the *_trace windows come back the same way through zzR_chain_windows; per-gene fields through gpu_pull())
  }
)
