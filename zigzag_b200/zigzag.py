"""Host-side mirror of the reference's `zigzag` Reference Class (R/zigzag.R) with the per-gene work on the GPU.

Same object API and move names as the R package: zigzag(data, gene_length, ...), .burnin(), .mcmc(), a 15-entry
`proposal_list` drawn with weights `proposal_probs` (R/zigzagInitialize.R:207-237), the same output files
(R/zigzagFileWriting.R).  The five per-gene moves and every O(G) reduction inside the hyper-parameter moves are calls
into the C ABI (include/zigzag_gpu.h); only scalar accept/reject logic runs here.  There is no CPU fallback.

Two schedules:
  schedule="reference": every generation draws 15 moves with replacement (R/zigzagBurnin.R:66, R/zigzagMCMC.R:172) and
      each per-gene move is one reference-operation-order kernel launch.
  schedule="fused":     every generation is ONE fused sweep (mhYg -> mhVariance_g -> allocation -> spike allocation over all
      genes, production kernels) followed by one pass over the hyper-parameter moves, whose likelihood sums come in closed
      form from the sufficient statistics the sweep emits.  A different but equally valid scan order.
Gene sharding: pass `shard=(rank, world)` and a torch.distributed process group; every reduction is all-reduced, the
hyper-parameter RNG is seeded identically on all ranks, so all replicas take identical decisions.
"""
import math
import os

import numpy as np

from . import _capi, synth

LN_SQRT_2PI = 0.918938533204672741780329736406
SQRT2PI = math.sqrt(2 * math.pi)


def _dgamma_log(x, shape, rate):
    if x < 0:
        return -math.inf
    if x == 0:
        return math.log(rate) if shape == 1 else (-math.inf if shape > 1 else math.inf)
    return shape * math.log(rate) + (shape - 1) * math.log(x) - rate * x - math.lgamma(shape)


def _dnorm_log(x, mu, sd):
    z = (x - mu) / sd
    return -(LN_SQRT_2PI + 0.5 * z * z + math.log(sd))


def _dunif_log(x, lo, hi):
    return -math.log(hi - lo) if lo <= x <= hi else -math.inf


def x_tune(window_sum, tuning, target=0.44, lo=0.0, hi=math.inf):
    """R/zigzagUtilities.R:161-181."""
    pjump = window_sum / 100
    nt = tuning * (math.tan(math.pi * pjump / 2) / math.tan(math.pi * target / 2))
    return min(max(nt, lo), hi)


class _Window:
    """100-wide accept window of a scalar parameter, c(rep(0,77), rep(1,23)) at start (R/zigzagInitialize.R:277)."""

    def __init__(self):
        self.w = [0] * 77 + [1] * 23

    def push0(self):
        self.w = self.w[1:] + [0]

    def accept(self):
        self.w[-1] = 1

    def sum(self):
        return sum(self.w)


class zigzag:
    def __init__(self, data, gene_length=None, candidate_gene_list="all", num_active_components=2,
                 weight_active_shape_1=2, weight_active_shape_2=2, shared_variance_prior_min=0.01, shared_variance_prior_max=5,
                 inactive_means_prior_shape=1, inactive_means_prior_rate=1 / 3, spike_prior_shape_1=1, spike_prior_shape_2=1,
                 active_means_dif_prior_shape=1, active_means_dif_prior_rate=1 / 3, shared_active_variance=True,
                 shared_variance=True, inactive_variances_prior_min=None, inactive_variances_prior_max=None,
                 active_variances_prior_min=None, active_variances_prior_max=None, output_directory="output", threshold_a=None, threshold_i=None, beta=1, temperature=1,
                 tau_shape=1, tau_rate=1, s0_mu=-1, s0_sigma=2, s1_shape=1, s1_rate=2, variance_g_upper_bound=math.inf,
                 alpha_r_shape=1, alpha_r_rate=1 / 10, active_gene_set=None, gene_names=None, seed=0, device=0,
                 shard=None, process_group=None, write_settings=True, init="host"):
        self.shared_variance = bool(shared_variance)
        self.shared_active_variance = True if shared_variance else bool(shared_active_variance)                # I:43-45
        if threshold_a is None or threshold_i is None or isinstance(num_active_components, str):
            raise NotImplementedError('threshold_a / threshold_i / num_active_components = "auto" rely on R\'s density() and loess() '
                                      "(R/zigzagPriorThresholdUtils.R, out of scope): pass them explicitly")
        data = np.asarray(data, dtype=np.float64)
        self.Xg = np.log(data, where=data > 0, out=np.full(data.shape, -np.inf)) if data.min() >= 0 else data.copy()   # I:52-56
        self.Xg = np.asfortranarray(self.Xg)
        self.num_transcripts, self.num_libraries = self.Xg.shape
        if self.num_libraries < 2:
            raise ValueError("Data must have > 1 library.")                                                    # I:135
        self.threshold_i = float(np.atleast_1d(threshold_i)[0])
        self.threshold_a = np.atleast_1d(np.asarray(threshold_a, dtype=np.float64))
        self.num_active_components = K = len(self.threshold_a)                                               # I:103-108
        self.multi_thresh_a = True                                                                             # I:111-113
        self.temperature, self.beta = float(temperature), float(beta)
        self.variance_g_upper_bound = float(variance_g_upper_bound)
        self.gene_names = list(gene_names) if gene_names is not None else [f"gene_{i + 1}" for i in range(self.num_transcripts)]
        gl = np.ones(self.num_transcripts) if gene_length is None else np.asarray(gene_length, dtype=np.float64).reshape(-1)
        self._gl_mean = float(gl.mean())
        self.candidate_gene_list = self.gene_names if isinstance(candidate_gene_list, str) and candidate_gene_list == "all" else list(candidate_gene_list)
        self._cand_idx = np.arange(self.num_transcripts) if self.candidate_gene_list is self.gene_names else \
            np.array([self.gene_names.index(g) for g in self.candidate_gene_list])
        self.output_directory = output_directory
        # ---- gene sharding (each rank keeps rows [g0, g1) of every per-gene array)
        self.rank, self.world = (0, 1) if shard is None else shard
        self.pg = process_group
        G = self.num_transcripts
        self.g0, self.g1 = self.rank * G // self.world, (self.rank + 1) * G // self.world
        self.gene_lengths = (gl / self._gl_mean)[self.g0:self.g1]                                              # I:143-149
        Xl = np.asfortranarray(self.Xg[self.g0:self.g1])
        # ---- priors (I:261-413)
        self.pri = dict(weight_active_shape_1=weight_active_shape_1, weight_active_shape_2=weight_active_shape_2,
                        spike_prior_shape_1=spike_prior_shape_1, spike_prior_shape_2=spike_prior_shape_2,
                        active_means_dif_prior_shape=active_means_dif_prior_shape, active_means_dif_prior_rate=active_means_dif_prior_rate,
                        inactive_means_prior_shape=inactive_means_prior_shape, inactive_means_prior_rate=inactive_means_prior_rate,
                        var_log_min=math.log10(shared_variance_prior_min), var_log_max=math.log10(shared_variance_prior_max),
                        tau_shape=tau_shape, tau_rate=tau_rate, s0_mu=s0_mu, s0_sigma=s0_sigma, s1_shape=s1_shape, s1_rate=s1_rate,
                        alpha_r_shape=alpha_r_shape, alpha_r_rate=alpha_r_rate,
                        shared_variance_prior_min=shared_variance_prior_min, shared_variance_prior_max=shared_variance_prior_max)
        # separate bounds of the two variance priors when the variance is not shared (I:296-349); shared: both are the shared bounds
        _b = lambda v, d: d if (v is None or self.shared_variance) else v
        self.pri.update(inactive_variances_prior_min=_b(inactive_variances_prior_min, shared_variance_prior_min),
                        inactive_variances_prior_max=_b(inactive_variances_prior_max, shared_variance_prior_max),
                        active_variances_prior_min=_b(active_variances_prior_min, shared_variance_prior_min),
                        active_variances_prior_max=_b(active_variances_prior_max, shared_variance_prior_max))
        self.pri.update(ivar_log_min=math.log10(self.pri["inactive_variances_prior_min"]), ivar_log_max=math.log10(self.pri["inactive_variances_prior_max"]),
                        avar_log_min=math.log10(self.pri["active_variances_prior_min"]), avar_log_max=math.log10(self.pri["active_variances_prior_max"]))
        self.weight_within_active_alpha = np.full(K, weight_active_shape_1 / K)                                # I:358
        # ---- tuning parameters (I:185-199) and their windows
        self.inactive_mean_tuningParam = 0.5
        self.inactive_variance_tuningParam = 0.5
        self.active_mean_tuningParam = np.full(K, 0.5)
        self.active_variance_tuningParam = np.full(K, 0.5)
        self.tuningParam_s0 = self.tuningParam_s1 = self.tuningParam_tau = self.tuningParam_s0tau = 0.05
        self.tuningParam_alpha_r = np.full(self.num_libraries, 0.5)
        self._win = dict(im=_Window(), iv=_Window(), avk=[_Window() for _ in range(K)], av=_Window(), s0=_Window(), s1=_Window(), tau=_Window(), s0tau=_Window(),
                         am=[_Window() for _ in range(K)], alpha=[_Window() for _ in range(self.num_libraries)])
        # ---- proposal table (I:207-237)
        R = self.num_libraries
        self.proposal_list = [self.gibbsMixtureWeights, self.gibbsAllocationActiveInactive, self.gibbsAllocationWithinActive,
                              self.mhInactiveMeans, self.mhInactiveVariances, self.mhActiveMeansDif,
                              self.mhSharedVariance if self.shared_variance else self.mhActiveVariances,                # I:223
                              self.gibbsSpikeProb, self.mhSpikeAllocation, self.mhYg, self.mhVariance_g, self.mhTau, self.mhSg,
                              self.mhS0Tau, self.mhP_x]
        yw = 1 + (R == 2) * 0.5 + 1 * (G < 15000)
        self.proposal_probs = np.array([4, 30, 5, 6, 7 * (1 - self.shared_variance), 5, 5 + 2 * (1 - self.shared_active_variance), 2, 5, yw, yw,
                                        8, 10, 8, R * 1.1], dtype=np.float64)                                          # I:228-237
        # ---- initial hyper-parameters: the same draws on every rank (I:263-408)
        self.rng = np.random.Generator(np.random.PCG64(seed))
        r = self.rng
        self.weight_active = r.beta(weight_active_shape_1 + 1, weight_active_shape_2 + 1)
        self.spike_probability = r.beta(spike_prior_shape_1, spike_prior_shape_2)
        self.active_means_dif = r.gamma(active_means_dif_prior_rate, 1 / active_means_dif_prior_rate, K)       # I:282 (shape = rate, sic)
        self.active_means = self.calculate_active_means(self.active_means_dif)
        av = 10 ** r.uniform(self.pri["avar_log_min"], self.pri["avar_log_max"])                               # I:296-323
        self.active_variances = np.full(K, av)
        self.inactive_means = self.threshold_i - r.gamma(1, 1)
        self.inactive_variances = av
        if not self.shared_variance:
            if not self.shared_active_variance:
                self.active_variances = 10 ** r.uniform(self.pri["avar_log_min"], self.pri["avar_log_max"], K)
            self.inactive_variances = 10 ** r.uniform(self.pri["ivar_log_min"], self.pri["ivar_log_max"])      # I:343-349
        self.weight_within_active = r.dirichlet(self.weight_within_active_alpha)
        self.alpha_r = r.gamma(1, 1, R)
        self.s0 = r.normal(0, 0.1)
        self.s1 = -r.gamma(1, 1 / 5)
        self.tau = r.gamma(2, 1 / 5)
        self.gen = 0
        self._step = 0
        self.lnl_trace = []
        self.n_samples = 0
        # ---- per-gene initial state (I:416-504): init="host" draws it per shard with NumPy (synth.initial_state), init="device" runs
        # the same recipe on the GPU (zz_init_state: Philox keyed by the global gene index, so it does not depend on the sharding)
        pinned = None
        if active_gene_set is not None:
            idx = [self.gene_names.index(g) for g in active_gene_set]
            pinned = np.zeros(G, np.int32); pinned[idx] = 1
            pinned = pinned[self.g0:self.g1]
        self._seed, self._device = seed, device
        self.h = _capi.Handle(self.g1 - self.g0, R, K, num_chains=1, device=device, seed=seed, gene_offset=self.g0)
        self.h.upload_data(Xl, self.gene_lengths)
        self._push_hyper()
        if init == "device":
            full = ~np.isneginf(self.Xg).any(1)
            rv = self.Xg[full].var(axis=1, ddof=1)                                                             # I:446, whole data set
            # fewer than two fully detected genes: rwv = 0, i.e. Yg = rwm (the reference's branch for data without any detection, I:452-455)
            rv_m, rv_s = (math.log(rv.mean()), math.sqrt(rv.var(ddof=1))) if len(rv) > 1 else (-math.inf, 0.0)
            if pinned is not None:
                self.h.set_state(0, pinned_active=pinned)
            self.h.init_state(0, rv_m, rv_s)
        else:
            st = synth.initial_state(Xl, self.gene_lengths, self._hyper_dict(), seed + 7919 * (self.rank + 1))
            if pinned is not None:
                st["alloc_active_inactive"][pinned == 1] = 1                                                   # I:369
            self.h.set_state(0, pinned_active=pinned, **st)
            self.h.refresh_caches()                                                                            # I:473, I:504
        self._pinned = pinned
        self._state_cache = None
        # gene shards on several GPUs of one node: peer-memory all-reduce inside the kernels (needed by schedule="device")
        self._p2p = False
        if self.world > 1:
            import torch.distributed as dist
            if dist.get_backend(self.pg) == "nccl":
                blobs = [None] * self.world
                dist.all_gather_object(blobs, self.h.comm_init(self.rank, self.world), group=self.pg)
                self.h.comm_connect(blobs)
                dist.barrier(group=self.pg)
                self._p2p = True
        if write_settings and self.rank == 0:
            self._write_hyperparameter_settings()

    # ------------------------------------------------------------------ plumbing
    def _hyper_dict(self):
        return dict(s0=self.s0, s1=self.s1, tau=self.tau, alpha_r=np.asarray(self.alpha_r), spike_probability=self.spike_probability,
                    inactive_means=self.inactive_means, inactive_variances=self.inactive_variances,
                    active_means=np.asarray(self.active_means), active_variances=np.asarray(self.active_variances),
                    weight_active=self.weight_active, weight_within_active=np.asarray(self.weight_within_active), beta=self.beta,
                    temperature=self.temperature, variance_g_upper_bound=self.variance_g_upper_bound)

    def _push_hyper(self):
        d = self._hyper_dict()
        hy = _capi.Hyper.make(d["alpha_r"], d["active_means"], d["active_variances"], d["weight_within_active"],
                              **{k: float(d[k]) for k in ("s0", "s1", "tau", "spike_probability", "inactive_means", "inactive_variances",
                                                          "weight_active", "beta", "temperature", "variance_g_upper_bound")})
        self.h.set_hyper(hy, 0)

    def _allreduce(self, v, op="sum"):
        """Sum (or maximum) of a small float64 vector over the gene shards (identical on every rank afterwards)."""
        v = np.asarray(v, dtype=np.float64)
        if self.world == 1:
            return v
        import torch
        import torch.distributed as dist
        t = torch.from_numpy(v.copy())
        if dist.get_backend(self.pg) == "nccl":
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM, group=self.pg)
        return t.cpu().numpy()

    def _allreduce_inplace(self, a):
        """In-place sum of a numpy array (float64 / uint32 / uint64: counts travel as float64, exact below 2^53) over the gene shards."""
        if a.dtype == np.uint64:          # exact: the two 32-bit halves are summed separately (each sum stays below 2^53) and recombined
            lo = self._allreduce((a & np.uint64(0xFFFFFFFF)).astype(np.float64)).astype(np.uint64)
            hi = self._allreduce((a >> np.uint64(32)).astype(np.float64)).astype(np.uint64)
            a[...] = (hi << np.uint64(32)) + lo
        else:
            a[...] = self._allreduce(a.astype(np.float64)).astype(a.dtype)

    # ------------------------------------------------------------------ schedule="device": whole generations on the GPU
    def _priors_struct(self):
        p = self.pri
        pr = _capi.Priors(p["weight_active_shape_1"], p["weight_active_shape_2"], p["spike_prior_shape_1"], p["spike_prior_shape_2"],
                          p["active_means_dif_prior_shape"], p["active_means_dif_prior_rate"], p["inactive_means_prior_shape"],
                          p["inactive_means_prior_rate"], p["var_log_min"], p["var_log_max"], p["tau_shape"], p["tau_rate"], p["s0_mu"],
                          p["s0_sigma"], p["s1_shape"], p["s1_rate"], p["alpha_r_shape"], p["alpha_r_rate"], self.threshold_i)
        for k, t in enumerate(self.threshold_a):
            pr.threshold_a[k] = float(t)
        return pr

    def _chain_scalars(self):
        """The scalar chain state (tuning parameters, windows, counters) in the layout of zz_chain_scalars."""
        import ctypes as C
        K, R, w = self.num_active_components, self.num_libraries, self._win
        win = lambda x: (C.c_uint8 * 100)(*[int(b) for b in x.w])
        sc = _capi.ChainScalars()
        sc.t_im, sc.t_av = float(self.inactive_mean_tuningParam), float(self.active_variance_tuningParam[0])
        sc.t_s0, sc.t_s1, sc.t_tau, sc.t_s0tau = (float(self.tuningParam_s0), float(self.tuningParam_s1), float(self.tuningParam_tau),
                                                  float(self.tuningParam_s0tau))
        for name in ("im", "av", "s0", "s1", "tau", "s0tau"):
            setattr(sc, "w_" + name, win(w[name]))
        for k in range(K):
            sc.active_means_dif[k] = float(self.active_means_dif[k]); sc.t_am[k] = float(self.active_mean_tuningParam[k]); sc.w_am[k] = win(w["am"][k])
        for r in range(R):
            sc.t_alpha[r] = float(self.tuningParam_alpha_r[r]); sc.w_alpha[r] = win(w["alpha"][r])
        sc.hyper_ctr = getattr(self, "_hyper_ctr", 0)
        for j, v in enumerate(getattr(self, "_hyper_buf", (0.0,) * 4)):
            sc.hyper_buf[j] = v
        sc.step, sc.gen, sc.n_samples = self._step + 1, self.gen, self.n_samples
        return sc

    def _adopt_chain(self, sc, hy):
        """Inverse of _chain_scalars / _push_hyper after a device-resident segment."""
        K, R = self.num_active_components, self.num_libraries
        self.s0, self.s1, self.tau = hy.s0, hy.s1, hy.tau
        self.alpha_r = np.array(hy.alpha_r[:R]); self.spike_probability = hy.spike_probability
        self.inactive_means, self.inactive_variances = hy.inactive_means, hy.inactive_variances
        self.active_means = np.array(hy.active_means[:K]); self.active_variances = np.array(hy.active_variances[:K])
        self.weight_active = hy.weight_active; self.weight_within_active = np.array(hy.weight_within_active[:K])
        self.active_means_dif = np.array(sc.active_means_dif[:K])
        self.inactive_mean_tuningParam = sc.t_im; self.active_variance_tuningParam = np.full(K, sc.t_av)
        self.active_mean_tuningParam = np.array(sc.t_am[:K]); self.tuningParam_alpha_r = np.array(sc.t_alpha[:R])
        self.tuningParam_s0, self.tuningParam_s1, self.tuningParam_tau, self.tuningParam_s0tau = sc.t_s0, sc.t_s1, sc.t_tau, sc.t_s0tau
        for name in ("im", "av", "s0", "s1", "tau", "s0tau"):
            self._win[name].w = list(getattr(sc, "w_" + name))
        for k in range(K):
            self._win["am"][k].w = list(sc.w_am[k])
        for r in range(R):
            self._win["alpha"][r].w = list(sc.w_alpha[r])
        self._hyper_ctr, self._hyper_buf = sc.hyper_ctr, tuple(sc.hyper_buf)
        self._step, self.gen, self.n_samples = sc.step - 1, sc.gen, sc.n_samples
        self._state_cache = None

    def _device_segment(self, ngen, tune, sample_frequency=1, target=0.44):
        """ngen generations of the fused schedule without leaving the GPU (zz_run_generations): sweep + every hyper move + tuning +
        allocation trace on the device; the host reads the scalar state back once per segment."""
        if self.world != 1 and not self._p2p:
            raise NotImplementedError('schedule="device" on gene shards needs the peer-memory all-reduce (one node, NCCL process group)')
        if not self.shared_variance:
            raise NotImplementedError('schedule="device" runs the default shared-variance model (its scalar stage holds mhSharedVariance); '
                                      'use schedule="fused" or "reference" with shared_variance=False')
        self.h.chain_begin(self._priors_struct(), self._chain_scalars())
        self.h.run_generations(ngen, tune=tune, sample_frequency=sample_frequency, target=target)
        self._adopt_chain(*self.h.chain_read())

    def _next_step(self):
        self._step += 1
        self._state_cache = None
        return self._step

    def _accept(self, log_ratio):
        return math.log(self.rng.random()) < log_ratio

    def state(self):
        """Per-gene fields of this shard pulled back from the device (lazily, like R code reading .self$Yg between moves)."""
        if self._state_cache is None:
            self._state_cache = self.h.get_state(0)
        return self._state_cache

    Yg = property(lambda self: self.state()["Yg"])
    variance_g = property(lambda self: self.state()["variance_g"])
    allocation_active_inactive = property(lambda self: self.state()["alloc_active_inactive"])
    allocation_within_active = property(lambda self: self.state()["alloc_within_active"])
    inactive_spike_allocation = property(lambda self: self.state()["inactive_spike_allocation"])
    XgLikelihood = property(lambda self: self.state()["XgLikelihood"])
    variance_g_probability = property(lambda self: self.state()["variance_g_probability"])
    tuningParam_yg = property(lambda self: self.state()["tuningParam_yg"])
    tuningParam_variance_g = property(lambda self: self.state()["tuningParam_variance_g"])

    def calculate_active_means(self, am=None):                                          # U:106-118 (multi_thresh_a)
        return (self.active_means_dif if am is None else np.asarray(am)) + self.threshold_a

    def get_sigmax(self, ss0=None, ss1=None, yy=None):                                   # U:144-148
        return np.exp((self.s0 if ss0 is None else ss0) + (self.s1 if ss1 is None else ss1) * (self.Yg if yy is None else yy))

    def get_px(self, falpha_r=None, yy=None, gl=None):                                   # U:150-159
        a = self.alpha_r if falpha_r is None else np.atleast_1d(falpha_r)
        return 1 - np.exp(-np.outer((self.gene_lengths if gl is None else gl) * np.exp(self.Yg if yy is None else yy), a))

    # ------------------------------------------------------------------ sufficient statistics (device reductions)
    def _stats(self):
        K1 = self.num_active_components + 1
        s = self._allreduce(self.h.suffstats()[0])
        return dict(n=s[:K1], sy=s[K1:2 * K1], sy2=s[2 * K1:3 * K1], n_inspike=s[3 * K1], n_out=s[3 * K1 + 1], sz=s[3 * K1 + 2],
                    sz2=s[3 * K1 + 3], syo=s[3 * K1 + 4], sy2o=s[3 * K1 + 5], syz=s[3 * K1 + 6], lnl=s[3 * K1 + 7],
                    svg=s[3 * K1 + 8], nan=s[3 * K1 + 9])

    def _level2_sums(self, im, iv, am, av, st=None):
        """(sum computeInactiveLikelihood over inactive genes, sum computeActiveLikelihood over active genes)."""
        if st is None:      # reference schedule: per-gene sums in the reference's own arithmetic
            return self._allreduce(self.h.level2_sums(im, iv, am, av))
        n0out = st["n"][0] - st["n_inspike"]
        li = st["n_inspike"] * math.log(self.spike_probability) + n0out * (math.log(1 - self.spike_probability) - math.log(SQRT2PI * math.sqrt(iv))) \
            - (st["sy2"][0] - 2 * im * st["sy"][0] + n0out * im * im) / (2 * iv)
        la = 0.0
        for k in range(self.num_active_components):
            n, sy, sy2 = st["n"][k + 1], st["sy"][k + 1], st["sy2"][k + 1]
            la += -n * math.log(SQRT2PI * math.sqrt(av[k])) - (sy2 - 2 * am[k] * sy + n * am[k] ** 2) / (2 * av[k])
        return np.array([li, la])

    def _varg_sum(self, s0, s1, tau, st=None):
        """(sum of computeVarianceGPriorProbability over out-of-spike genes at (s0,s1,tau), same at the current parameters)."""
        if st is None or self.variance_g_upper_bound != math.inf:
            return self._allreduce(self.h.varg_hyper_sum(s0, s1, tau))

        def closed(s0, s1, tau):
            c, n = s0 + tau, st["n_out"]
            q = st["sz2"] + n * c * c + s1 * s1 * st["sy2o"] - 2 * c * st["sz"] - 2 * s1 * st["syz"] + 2 * c * s1 * st["syo"]
            return -st["sz"] - n * math.log(math.sqrt(tau) * SQRT2PI) - q / (2 * tau)
        return np.array([closed(s0, s1, tau), closed(self.s0, self.s1, self.tau)])

    # ------------------------------------------------------------------ the 15 proposal_list entries
    def gibbsMixtureWeights(self, tune=False, st=None):                                  # Pr:412-431
        st = st or self._stats()
        alpha = np.concatenate([[self.pri["weight_active_shape_2"]], self.weight_within_active_alpha]) + st["n"] * self.beta
        g = self.rng.gamma(alpha, 1.0)
        new = g / g.sum()
        if not np.isnan(new[0]):
            self.weight_active = 1 - new[0]
            self.weight_within_active = new[1:] / new[1:].sum()
            self._push_hyper()

    def gibbsAllocationActiveInactive(self, tune=False):                                 # Pr:335-361 (+ chained Pr:363-410)
        self.h.gibbs_alloc(self._next_step())

    def gibbsAllocationWithinActive(self, tune=False):                                   # Pr:363-410
        self.h.gibbs_alloc_within(self._next_step())

    def mhInactiveMeans(self, tune=False, st=None):                                      # Pr:433-471
        thr, im = self.threshold_i, self.inactive_means
        imp = thr - abs(thr - im - self.rng.normal(0, self.inactive_mean_tuningParam))
        pp = _dgamma_log(thr - imp, self.pri["inactive_means_prior_shape"], self.pri["inactive_means_prior_rate"])
        pc = _dgamma_log(thr - im, self.pri["inactive_means_prior_shape"], self.pri["inactive_means_prior_rate"])
        Lp = self._level2_sums(imp, self.inactive_variances, self.active_means, self.active_variances, st)[0]
        Lc = self._level2_sums(im, self.inactive_variances, self.active_means, self.active_variances, st)[0]
        if tune:
            self._win["im"].push0()
        if self._accept(self.temperature * (Lp - Lc + pp - pc)):
            self.inactive_means = imp
            if tune:
                self._win["im"].accept()
            self._push_hyper()

    def mhInactiveVariances(self, tune=False, st=None):                                  # Pr:473-509 (weight 7 * (1 - shared_variance))
        if self.shared_variance:
            return                                                                       # weight 0 in the default model
        lo, hi = self.pri["ivar_log_min"], self.pri["ivar_log_max"]
        ivp = 10 ** self._two_boundary_slide_move(math.log10(self.inactive_variances), lo, hi, self.inactive_variance_tuningParam)
        if tune:
            self._win["iv"].push0()
        pp, pc = _dunif_log(math.log10(ivp), lo, hi), _dunif_log(math.log10(self.inactive_variances), lo, hi)
        Lp = self._level2_sums(self.inactive_means, ivp, self.active_means, self.active_variances, st)[0]
        Lc = self._level2_sums(self.inactive_means, self.inactive_variances, self.active_means, self.active_variances, st)[0]
        if self._accept(self.temperature * (Lp - Lc + pp - pc)):
            self.inactive_variances = ivp
            if tune:
                self._win["iv"].accept()
            self._push_hyper()

    def mhActiveVariances(self, tune=False, st=None):                                    # Pr:707-766
        K = self.num_active_components
        lo, hi = self.pri["avar_log_min"], self.pri["avar_log_max"]
        k = int(self.rng.integers(0, K))
        avp = np.array(self.active_variances, dtype=np.float64)
        if self.shared_active_variance:
            avp[:] = 10 ** self._two_boundary_slide_move(math.log10(self.active_variances[0]), lo, hi, self.active_variance_tuningParam[0])
            win = self._win["av"]
        else:
            avp[k] = 10 ** self._two_boundary_slide_move(math.log10(self.active_variances[k]), lo, hi, self.active_variance_tuningParam[k])
            win = self._win["avk"][k]
        if tune:
            win.push0()
        Lp = self._level2_sums(self.inactive_means, self.inactive_variances, self.active_means, avp, st)[1]
        Lc = self._level2_sums(self.inactive_means, self.inactive_variances, self.active_means, self.active_variances, st)[1]
        pp, pc = _dunif_log(math.log10(avp[k]), lo, hi), _dunif_log(math.log10(self.active_variances[k]), lo, hi)
        if self._accept(self.temperature * (Lp + pp - Lc - pc)):
            self.active_variances = avp
            if tune:
                win.accept()
            self._push_hyper()

    def mhActiveMeansDif(self, tune=False, st=None):                                     # Pr:640-705
        K = self.num_active_components
        for k in self.rng.integers(0, K, K):
            difp = self.active_means_dif.copy()
            difp[k] = abs(self.active_means_dif[k] + self.rng.normal(0, self.active_mean_tuningParam[k]))
            mp = self.calculate_active_means(difp)
            Lp = self._level2_sums(self.inactive_means, self.inactive_variances, mp, self.active_variances, st)[1]
            Lc = self._level2_sums(self.inactive_means, self.inactive_variances, self.active_means, self.active_variances, st)[1]
            pp = _dgamma_log(difp[k], self.pri["active_means_dif_prior_shape"], self.pri["active_means_dif_prior_rate"])
            pc = _dgamma_log(self.active_means_dif[k], self.pri["active_means_dif_prior_shape"], self.pri["active_means_dif_prior_rate"])
            if tune:
                self._win["am"][k].push0()
            if self._accept(self.temperature * (Lp + pp - Lc - pc)):
                self.active_means_dif = difp
                self.active_means = mp
                if tune:
                    self._win["am"][k].accept()
                self._push_hyper()

    def _two_boundary_slide_move(self, param, lower, upper, t):                           # MP:29-46
        new = self.rng.uniform(param - t, param + t)
        if new > upper:
            new = 2 * upper - new
        return lower + abs(new - lower)

    def mhSharedVariance(self, tune=False, st=None):                                     # Pr:768-803
        lo, hi, K = self.pri["var_log_min"], self.pri["var_log_max"], self.num_active_components
        avp1 = 10 ** self._two_boundary_slide_move(math.log10(self.active_variances[0]), lo, hi, self.active_variance_tuningParam[0])
        avp = np.full(K, avp1)
        if tune:
            self._win["av"].push0()
        Lp = self._level2_sums(self.inactive_means, avp1, self.active_means, avp, st)
        Lc = self._level2_sums(self.inactive_means, self.inactive_variances, self.active_means, self.active_variances, st)
        app, apc = _dunif_log(math.log10(avp1), lo, hi), _dunif_log(math.log10(self.active_variances[0]), lo, hi)
        if self._accept(self.temperature * (Lp[1] + Lp[0] + app - Lc[1] - Lc[0] - apc)):
            self.active_variances = avp
            self.inactive_variances = avp1
            if tune:
                self._win["av"].accept()
            self._push_hyper()

    def gibbsSpikeProb(self, tune=False, st=None):                                       # Pr:511-526
        st = st or self._stats()
        n_in, n_inactive = st["n_inspike"], st["n"][0]
        new = self.rng.beta(n_in + self.pri["spike_prior_shape_1"], (n_inactive - n_in) + self.pri["spike_prior_shape_2"])
        if not math.isnan(new):
            self.spike_probability = new
            self._push_hyper()

    def mhSpikeAllocation(self, tune=False):                                             # Pr:528-638
        self.h.mh_spike_alloc(self._next_step())

    def mhYg(self, tune=False):                                                          # Pr:250-333
        self.h.mh_yg(self._next_step(), tune=tune)

    def mhVariance_g(self, tune=False):                                                  # Pr:7-56
        self.h.mh_variance_g(self._next_step(), tune=tune)

    def _s0_prior(self, x):
        return _dnorm_log(x, self.pri["s0_mu"], self.pri["s0_sigma"])                    # P:33-39

    def _s1_prior(self, x):
        return _dgamma_log(abs(x), self.pri["s1_shape"], self.pri["s1_rate"])            # P:41-47

    def _tau_prior(self, x):
        return _dgamma_log(x, self.pri["tau_shape"], self.pri["tau_rate"])               # P:3-9

    def _scale_move(self, param, t):                                                     # MP:3-13
        sf = math.exp(t * (self.rng.random() - 0.5))
        return param * sf, sf

    def _commit_varg(self, st):
        self._push_hyper()
        if st is None:          # reference schedule: overwrite variance_g_probability[out_spike_idx] (Pr:113,147,187)
            self.h.commit_varg_hyper(0)
            self._state_cache = None

    def mhTau(self, tune=False, st=None):                                                # Pr:119-152
        taup, sf = self._scale_move(self.tau, self.tuningParam_tau)
        Lp, Lc = self._varg_sum(self.s0, self.s1, taup, st)
        Rr = self.temperature * (Lp - Lc + self._tau_prior(taup) - self._tau_prior(self.tau)) + math.log(sf)
        if tune:
            self._win["tau"].push0()
        if self._accept(Rr):
            self.tau = taup
            if tune:
                self._win["tau"].accept()
            self._commit_varg(st)

    def mhSg(self, tune=False, st=None):                                                 # Pr:58-117
        HR, s0p, s1p = 0.0, self.s0, self.s1
        sel = 1 if self.rng.random() < 0.5 else 2
        if sel == 1:
            s0p = self.s0 + self.rng.normal(0, self.tuningParam_s0)
            if tune:
                self._win["s0"].push0()
        else:
            s1p, sf = self._scale_move(self.s1, self.tuningParam_s1)
            HR = math.log(sf)
            if tune:
                self._win["s1"].push0()
        Lp, Lc = self._varg_sum(s0p, s1p, self.tau, st)
        pp, pc = self._s0_prior(s0p) + self._s1_prior(s1p), self._s0_prior(self.s0) + self._s1_prior(self.s1)
        if self._accept(self.temperature * (Lp - Lc + pp - pc) + HR):
            self.s0, self.s1 = s0p, s1p
            if tune:
                self._win["s0" if sel == 1 else "s1"].accept()
            self._commit_varg(st)

    def mhS0Tau(self, tune=False, st=None):                                              # Pr:154-191
        taup, sf = self._scale_move(self.tau, self.tuningParam_s0tau)
        s0p, HR = self.s0 - math.log(sf), math.log(sf)
        if tune:
            self._win["s0tau"].push0()
        Lp, Lc = self._varg_sum(s0p, self.s1, taup, st)
        pp, pc = self._s0_prior(s0p) + self._tau_prior(taup), self._s0_prior(self.s0) + self._tau_prior(self.tau)
        if self._accept(self.temperature * (Lp - Lc + pp - pc) + HR):
            self.s0, self.tau = s0p, taup
            if tune:
                self._win["s0tau"].accept()
            self._commit_varg(st)

    def mhP_x(self, tune=False, st=None):                                                # Pr:193-243
        R = self.num_libraries
        lib = int(self.rng.integers(0, R))
        ap, sf = self._scale_move(self.alpha_r[lib], self.tuningParam_alpha_r[lib])
        if tune:
            self._win["alpha"][lib].push0()
        prop = np.array(self.alpha_r, dtype=np.float64)
        prop[lib] = ap
        mask = np.zeros(R, np.uint8); mask[lib] = 1
        delta = self._allreduce(self.h.px_delta(prop, mask))[lib]                        # Lp - Lc, Pr:223-224
        pp = sum(_dgamma_log(a, self.pri["alpha_r_shape"], self.pri["alpha_r_rate"]) for a in prop)
        pc = sum(_dgamma_log(a, self.pri["alpha_r_shape"], self.pri["alpha_r_rate"]) for a in self.alpha_r)
        Rr = self.temperature * (delta + pp - pc) + math.log(sf)
        if self._accept(Rr) and Rr != math.inf:                                          # Pr:232
            self.alpha_r = prop
            self.h.px_commit(lib, ap)                                                    # Pr:234-239
            if tune:
                self._win["alpha"][lib].accept()
            self._state_cache = None

    # ------------------------------------------------------------------ tuning, lnl, sampling
    def tune_all(self, target):                                                          # U:183-240
        w = self._win
        self.tuningParam_alpha_r = np.array([x_tune(w["alpha"][r].sum(), self.tuningParam_alpha_r[r], target, 0.01, 2) for r in range(self.num_libraries)])
        self.tuningParam_s0 = x_tune(w["s0"].sum(), self.tuningParam_s0, target, 0.0001, 5)
        self.tuningParam_s1 = x_tune(w["s1"].sum(), self.tuningParam_s1, target, 0.0001, 5)
        self.tuningParam_tau = x_tune(w["tau"].sum(), self.tuningParam_tau, target, 0.0001, 5)
        self.tuningParam_s0tau = x_tune(w["s0tau"].sum(), self.tuningParam_s0tau, target, 0.0001, 5)
        self.h.tune(target)                                                              # U:202-206 on the device
        self._state_cache = None
        self.inactive_mean_tuningParam = x_tune(w["im"].sum(), self.inactive_mean_tuningParam, target, 0.001, 10)
        self.active_mean_tuningParam = np.array([x_tune(w["am"][k].sum(), self.active_mean_tuningParam[k], target, 0.01, 10)
                                                 for k in range(self.num_active_components)])
        self.inactive_variance_tuningParam = x_tune(w["iv"].sum(), self.inactive_variance_tuningParam, target, 0.01,
                                                    self.pri["ivar_log_max"] - self.pri["ivar_log_min"])          # U:215-217
        span = self.pri["avar_log_max"] - self.pri["avar_log_min"]
        if self.shared_active_variance:                                                                        # U:224-237
            self.active_variance_tuningParam = np.full(self.num_active_components, x_tune(w["av"].sum(), self.active_variance_tuningParam[0], target, 0.01, span))
        else:
            self.active_variance_tuningParam = np.array([x_tune(w["avk"][k].sum(), self.active_variance_tuningParam[k], target, 0.01, span)
                                                         for k in range(self.num_active_components)])

    def calculate_lnl(self):                                                             # U:70-75
        return float(self._allreduce([self.h.lnl()])[0])

    def _check_nan(self):
        """Replaces the 9 full-length sums of the reference's NA guard (R/zigzagBurnin.R:69-78) by a device flag."""
        if self._allreduce([float(self.h.nan_flag())])[0] > 0:
            raise FloatingPointError("an NA or NaN somewhere! (device NaN flag raised)")

    def _generation(self, tune, schedule):
        if schedule == "reference":
            for p in self.rng.choice(15, 15, replace=True, p=self.proposal_probs / self.proposal_probs.sum()):   # B:66 / M:172
                self.proposal_list[p](tune=tune)
        else:
            self.h.sweep(self._next_step(), tune=tune)
            st = self._stats()
            var_moves = (self.mhSharedVariance,) if self.shared_variance else (self.mhInactiveVariances, self.mhActiveVariances)
            for mv in (self.gibbsMixtureWeights, self.mhInactiveMeans, self.mhActiveMeansDif, *var_moves, self.gibbsSpikeProb,
                       self.mhTau, self.mhSg, self.mhS0Tau):
                mv(tune=tune, st=st)
            self.mhP_x(tune=tune)

    def burnin(self, sample_frequency=100, ngen=10000, burnin_target_acceptance_rate=0.44, progress_plot=False, write_to_files=True,
               burninprefix="output", append=False, schedule="reference"):
        """R/zigzagBurnin.R:21-152."""
        prefix = burninprefix.split("/")[-1]
        pdir = burninprefix + "_burnin"
        if write_to_files and self.rank == 0 and not append:
            os.makedirs(os.path.join(self.output_directory, pdir), exist_ok=True)
            self.initializeOutputFiles(pdir + "/" + prefix)
        if schedule == "device":
            # segments end at the sampling points of B:116-131 (the device re-tunes by itself every 400 generations, B:146)
            done = 0
            while done <= ngen:
                n = min((1 - self.gen) % sample_frequency or sample_frequency, ngen + 1 - done)   # end right after generation i = k * sf
                self._device_segment(n, True, target=burnin_target_acceptance_rate)
                done += n
                self._check_nan()
                i = self.gen - 1
                if i % sample_frequency == 0 and i > 2 * sample_frequency:
                    self.lnl_trace.append(self.calculate_lnl())
                    if write_to_files and self.temperature == 1:
                        self._write_sample(pdir + "/" + prefix, i, (i // sample_frequency) % 4 == 0)
            self.lnl_trace = self.lnl_trace[-1:]
            return
        i, j = self.gen, 0
        while j <= ngen:
            self._generation(True, schedule)
            self._check_nan()
            j = i
            if i % sample_frequency == 0 and i > 2 * sample_frequency:
                self.lnl_trace.append(self.calculate_lnl())
                if write_to_files and self.temperature == 1:
                    self._write_sample(pdir + "/" + prefix, i, (i // sample_frequency) % 4 == 0)
            i += 1
            self.gen = i
            if self.gen % 400 == 0:
                self.tune_all(burnin_target_acceptance_rate)                              # B:146
        self.lnl_trace = self.lnl_trace[-1:]

    def mcmc(self, sample_frequency=100, ngen=10000, write_to_files=True, mcmcprefix="out", compute_probs=True, append=False,
             schedule="reference", run_posterior_predictive=False):
        """R/zigzagMCMC.R:26-311 (posterior-predictive simulation and the PDF reports are out of scope)."""
        prefix = mcmcprefix.split("/")[-1]
        pdir = mcmcprefix + "_mcmc_output"
        if write_to_files and self.rank == 0 and not append:
            os.makedirs(os.path.join(self.output_directory, pdir), exist_ok=True)
            self.initializeOutputFiles(pdir + "/" + prefix, post_pred=run_posterior_predictive)
        postpred_every = max(int(round(ngen / (500 * sample_frequency))), 1) * sample_frequency       # M:83-84, M:250
        pp_prefix = (pdir + "/" + prefix) if write_to_files else None
        self.h.reset_alloc_trace(); self.h.accumulate_alloc_trace(); self.n_samples = 1       # M:142-145
        self.lnl_trace.append(self.calculate_lnl())
        j, starting_gen = self.gen, self.gen
        yv_freq = max(ngen // (sample_frequency * 500), 1)
        write_yv = write_to_files
        ngen = ngen + self.gen
        if schedule == "device" and self.world == 1 and not run_posterior_predictive:
            # ONE device-resident segment for the whole run: the device accumulates allocation_trace (M:194-199) and appends the rows
            # of the three log files to pinned host rings at every sampling point (zz_set_sample_sink); the host formats them afterwards
            first, last = self.gen, ngen
            samples = [g for g in range(first, last + 1) if g % sample_frequency == 0]
            logging = write_to_files and self.temperature == 1
            self.h.set_sample_sink(max(1, len(samples)), self._cand_idx if logging else None, 501 if logging else 0, yv_freq)
            self._device_segment(last - first + 1, False, sample_frequency=sample_frequency)
            prow, yrow, vrow = self.h.sample_rows()
            self.h.clear_sample_sink()
            self.lnl_trace.extend(float(r[1]) for r in prow)
            if logging:
                with open(self._path(pdir + "/" + prefix, "_model_parameters.log"), "a") as f:
                    for r in prow:
                        f.write("\t".join([_rnum(r[0])] + [_rnum(round(float(x), 4)) for x in r[1:]]) + "\n")
                with open(self._path(pdir + "/" + prefix, "_yg_candidate_genes.log"), "a") as f:
                    for r in yrow:
                        f.write("\t".join([_rnum(r[0])] + [_rnum(round(float(x), 3)) for x in r[1:]]) + "\n")
                with open(self._path(pdir + "/" + prefix, "_varianceg_candidate_genes.log"), "a") as f:
                    for r in vrow:
                        f.write("\t".join([_rnum(r[0])] + [_rnum(round(float(x), 4)) for x in r[1:]]) + "\n")
            self._check_nan()
            if compute_probs and write_to_files:
                self.computeGeneExpressionProbs_writeToFile(pdir + "/" + prefix)
            return
        if schedule == "device":
            # the device accumulates allocation_trace itself whenever gen % sample_frequency == 0 (M:194-199); segments end right
            # after such a generation so that the log rows can be written from the state of that generation
            while self.gen <= ngen:
                n = min((1 - self.gen) % sample_frequency or sample_frequency, ngen + 1 - self.gen)
                self._device_segment(n, False, sample_frequency=sample_frequency)
                g = self.gen - 1
                if g % sample_frequency == 0:
                    self.lnl_trace.append(self.calculate_lnl())
                    if write_to_files and write_yv and self.temperature == 1:
                        self._write_sample(pdir + "/" + prefix, g, (g // sample_frequency) % yv_freq == 0)
                        if (g - starting_gen) / (yv_freq * sample_frequency) > 500:
                            write_yv = False
                if run_posterior_predictive and g % postpred_every == 0:                          # M:250-252
                    self.posteriorPredictiveSimulation(pp_prefix, gen=g)
            self._check_nan()
            if compute_probs and write_to_files:
                self.computeGeneExpressionProbs_writeToFile(pdir + "/" + prefix)
            return
        while j <= ngen:
            self._generation(False, schedule)
            j = self.gen
            if self.gen % sample_frequency == 0:
                if self.temperature == 1:
                    self.h.accumulate_alloc_trace(); self.n_samples += 1                          # M:194-199
                self.lnl_trace.append(self.calculate_lnl())                                       # M:202
                if write_to_files and write_yv and self.temperature == 1:
                    self._write_sample(pdir + "/" + prefix, self.gen, (self.gen // sample_frequency) % yv_freq == 0)
                    if (self.gen - starting_gen) / (yv_freq * sample_frequency) > 500:
                        write_yv = False
            if run_posterior_predictive and self.gen % postpred_every == 0:                       # M:250-252
                self.posteriorPredictiveSimulation(pp_prefix)
            self.gen += 1
        self._check_nan()
        if compute_probs and write_to_files:
            self.computeGeneExpressionProbs_writeToFile(pdir + "/" + prefix)

    # ------------------------------------------------------------------ the chain axis: tempering and stepping stones
    def _batched_handle(self, nchains, temperatures=None, betas=None):
        """A handle with `nchains` chains over the same data, every chain a copy of this object's state and hyper-parameters except
        for its temperature / beta, plus the per-chain scalar state in the layout of zz_chain_begin."""
        if self.world != 1:
            raise NotImplementedError("batched chains live on one GPU (shard chains, not genes, over several GPUs)")
        hb = _capi.Handle(self.num_transcripts, self.num_libraries, self.num_active_components, num_chains=nchains, device=self._device,
                          seed=self._seed + 104729, gene_offset=0)
        hb.upload_data(np.asfortranarray(self.Xg), self.gene_lengths)
        st = self.h.get_state(0)
        d = self._hyper_dict()
        for c in range(nchains):
            dc = dict(d)
            if temperatures is not None:
                dc["temperature"] = float(temperatures[c])
            if betas is not None:
                dc["beta"] = float(betas[c])
            hb.set_hyper(_capi.Hyper.make(dc["alpha_r"], dc["active_means"], dc["active_variances"], dc["weight_within_active"],
                                          **{k: float(dc[k]) for k in ("s0", "s1", "tau", "spike_probability", "inactive_means", "inactive_variances",
                                                                       "weight_active", "beta", "temperature", "variance_g_upper_bound")}), c)
            hb.set_state(c, pinned_active=self._pinned, **{k: st[k] for k in ("Yg", "variance_g", "alloc_active_inactive", "alloc_within_active",
                                                                             "inactive_spike_allocation", "tuningParam_yg", "tuningParam_variance_g")})
        hb.refresh_caches()
        hb.chain_begin(self._priors_struct(), [self._chain_scalars() for _ in range(nchains)])
        return hb

    def _log_prior(self, hy, sc):
        """lnPrior of get_logPosterior (R/zigzagUtilities.R:249-254) without its per-gene term, from zz_hyper + zz_chain_scalars."""
        p, K, R = self.pri, self.num_active_components, self.num_libraries
        lbeta = lambda a, b: math.lgamma(a) + math.lgamma(b) - math.lgamma(a + b)
        dbeta = lambda x, a, b: (a - 1) * math.log(x) + (b - 1) * math.log1p(-x) - lbeta(a, b)
        ww, al = list(hy.weight_within_active[:K]), self.weight_within_active_alpha
        ddir = sum((a - 1) * math.log(x) for a, x in zip(al, ww)) + math.lgamma(sum(al)) - sum(math.lgamma(a) for a in al)
        return (dbeta(hy.spike_probability, p["spike_prior_shape_1"], p["spike_prior_shape_2"])
                + sum(_dgamma_log(a, p["alpha_r_shape"], p["alpha_r_rate"]) for a in hy.alpha_r[:R])
                + dbeta(hy.weight_active, p["weight_active_shape_1"], p["weight_active_shape_2"]) + ddir
                + sum(_dgamma_log(x, p["active_means_dif_prior_shape"], p["active_means_dif_prior_rate"]) for x in sc.active_means_dif[:K])
                + sum(_dunif_log(math.log10(v), p["var_log_min"], p["var_log_max"]) for v in hy.active_variances[:K])
                + _dgamma_log(self.threshold_i - hy.inactive_means, p["inactive_means_prior_shape"], p["inactive_means_prior_rate"])
                + _dunif_log(math.log10(hy.inactive_variances), p["var_log_min"], p["var_log_max"])
                + _dgamma_log(hy.tau, p["tau_shape"], p["tau_rate"]) + _dnorm_log(hy.s0, p["s0_mu"], p["s0_sigma"])
                + _dgamma_log(abs(hy.s1), p["s1_shape"], p["s1_rate"]))

    def _log_posterior_untempered(self, stats, hy, sc):
        """get_logPosterior(t = 1) of one chain (U:241-263): data likelihood + level-2 likelihoods + variance prior (gene sums, from the
        chain's sufficient statistics) + the scalar priors."""
        K1 = self.num_active_components + 1
        st = dict(n=stats[:K1], sy=stats[K1:2 * K1], sy2=stats[2 * K1:3 * K1], n_inspike=stats[3 * K1])
        n0out = st["n"][0] - st["n_inspike"]
        iv, im, sp = hy.inactive_variances, hy.inactive_means, hy.spike_probability
        li = st["n_inspike"] * math.log(sp) + n0out * (math.log(1 - sp) - math.log(SQRT2PI * math.sqrt(iv))) \
            - (st["sy2"][0] - 2 * im * st["sy"][0] + n0out * im * im) / (2 * iv)
        la = 0.0
        for k in range(self.num_active_components):
            n, sy, sy2, am, av = st["n"][k + 1], st["sy"][k + 1], st["sy2"][k + 1], hy.active_means[k], hy.active_variances[k]
            la += -n * math.log(SQRT2PI * math.sqrt(av)) - (sy2 - 2 * am * sy + n * am * am) / (2 * av)
        return float(stats[3 * K1 + 7] + stats[3 * K1 + 8] + li + la) + self._log_prior(hy, sc)

    def mc3(self, nchains=4, burnin_ngen=1000, mcmc_ngen=10000, sample_frequency=20, dT=0.025, swap_every=1, write_to_files=True,
            mcmcprefix="mc3_out", compute_probs=True):
        """Metropolis-coupled chains (R/zigzagDevelopment.R:9-139) on the CHAIN AXIS of one batched handle: chain c runs at temperature
        1 / (1 + dT c); after every `swap_every` generations (the reference: 1) each neighbouring pair proposes to exchange temperatures
        with ratio posterior1(T2) + posterior2(T1) - posterior1(T1) - posterior2(T2) (get_logPosterior, U:241-263).  States stay in
        their slots, temperatures move; the slot that becomes cold inherits the cold chain's tuning parameters (R/zigzagDevelopment.R:
        113-120) and the cold chain's samples accumulate in ONE allocation trace.  Returns (swaps accepted, swaps proposed)."""
        temps = [1.0 / (1.0 + dT * c) for c in range(nchains)]
        hb = self._batched_handle(nchains, temperatures=temps)
        K1 = self.num_active_components + 1
        hb.run_generations(burnin_ngen, tune=True, target=0.23)                                   # R/zigzagDevelopment.R:44-52
        prefix, pdir = mcmcprefix.split("/")[-1], mcmcprefix + "_mcmc_output"
        if write_to_files:
            os.makedirs(os.path.join(self.output_directory, pdir), exist_ok=True)
            self.initializeOutputFiles(pdir + "/" + prefix)
        slot = list(range(nchains))                                                                # slot[i] = chain slot at temperature temps[i]
        acc = prop = 0
        n_samples, gen0 = 0, self.gen
        for cycle in range(mcmc_ngen // swap_every):
            hb.run_generations(swap_every, tune=False, sample_frequency=1 << 30)
            sc, hy = hb.chain_read()
            stats = hb.suffstats()
            L = [self._log_posterior_untempered(stats[c], hy[c], sc[c]) for c in range(nchains)]
            changed = False
            for i in range(nchains - 1):                                                           # every neighbouring pair, in order (:75)
                a, b = slot[i], slot[i + 1]
                prop += 1
                if self._accept((temps[i + 1] - temps[i]) * (L[a] - L[b])):                        # t2 L1 + t1 L2 - t1 L1 - t2 L2
                    acc += 1; changed = True
                    slot[i], slot[i + 1] = b, a
                    if i == 0:                                                                     # the cold chain moves to slot b (:107-120)
                        hb.copy_chain_tuning(a, b)
                        for name in ("t_im", "t_av", "t_s0", "t_s1", "t_tau", "t_s0tau"):
                            setattr(sc[b], name, getattr(sc[a], name))
                        sc[b].t_am, sc[b].t_alpha = sc[a].t_am, sc[a].t_alpha
            if changed:
                for i, c in enumerate(slot):
                    hy[c].temperature = temps[i]
                    hb.set_hyper(hy[c], c)
                hb.chain_begin(self._priors_struct(), sc)
            g = gen0 + (cycle + 1) * swap_every
            if g % sample_frequency < swap_every:                                                  # sample the cold chain (M:188-217)
                cold = slot[0]
                hb.accumulate_alloc_trace_from(cold, 0)
                n_samples += 1
                self.lnl_trace.append(float(stats[cold][3 * K1 + 7]))
                if write_to_files:
                    h0 = hy[cold]
                    row = [self.lnl_trace[-1], h0.s0, h0.s1, h0.tau, *h0.alpha_r[:self.num_libraries], h0.spike_probability, h0.inactive_means,
                           h0.inactive_variances, *h0.active_means[:K1 - 1], *h0.active_variances[:K1 - 1], h0.weight_active,
                           *h0.weight_within_active[:K1 - 1]]
                    with open(self._path(pdir + "/" + prefix, "_model_parameters.log"), "a") as f:
                        f.write("\t".join([_rnum(g)] + [_rnum(round(float(x), 4)) for x in row]) + "\n")
        # the cold chain becomes this object's state
        sc, hy = hb.chain_read()
        cold = slot[0]
        st = hb.get_state(cold)
        self._adopt_chain(sc[cold], hy[cold])
        self.temperature = 1.0
        self.h.set_state(0, **st)
        self._push_hyper()
        self.h.refresh_caches()
        tr = hb.get_alloc_trace(0).T.astype(np.float64)
        self.mc3_prob_active = 1 - np.round(tr[:, 0] / max(tr[0].sum(), 1.0), 3)
        self.mc3_temperatures, self.mc3_slots = temps, slot
        if write_to_files and compute_probs:
            with open(self._path(pdir + "/" + prefix, "_probability_active.tab"), "w") as f:
                f.write("gene\tprob_active\n")
                for gname, pa in zip(self.gene_names, self.mc3_prob_active):
                    f.write(f"{gname}\t{_rfmt3(pa)}\n")
        hb.close()
        return acc, prop

    def estimateMarginalLikelihood(self, burnin_gen=1000, mcmc_gen=10000, ss_sample_frequency=10, power_alpha=0.2, stones=10):
        """Stepping-stone estimate of the log marginal likelihood (R/zigzagDevelopment.R:142-208).  The reference runs the stones
        as parallel R processes; here every stone k = 1 .. stones - 1 is one chain of a batched handle at power beta_k =
        (k / stones)^(1 / power_alpha): all of them burn in and sample side by side in the same launches.  Stone k contributes
        (b_{k+1} - b_k) max + log mean exp((b_{k+1} - b_k)(lnl / b_k - max)) over its sampled lnl (:161-165).
        Returns (sum over stones, per-stone terms, betas)."""
        ssbeta = (np.arange(1, stones + 1) / stones) ** (1.0 / power_alpha)
        C = stones - 1
        hb = self._batched_handle(C, betas=ssbeta[:C])
        K1 = self.num_active_components + 1
        hb.run_generations(burnin_gen, tune=True)
        lnl = [[] for _ in range(C)]
        stats = hb.suffstats()
        for c in range(C):
            lnl[c].append(float(stats[c][3 * K1 + 7]))                                            # M:145: the trace starts at the current state
        done = 0
        while done < mcmc_gen:
            n = min(ss_sample_frequency, mcmc_gen - done)
            hb.run_generations(n, tune=False, sample_frequency=1 << 30)
            done += n
            stats = hb.suffstats()
            for c in range(C):
                lnl[c].append(float(stats[c][3 * K1 + 7]))
        terms = []
        for k in range(C):
            ss = np.array(lnl[k]) / ssbeta[k]                                                      # :161 (lnl_trace holds beta * lnl, P:125)
            db, mx = ssbeta[k + 1] - ssbeta[k], ss.max()
            terms.append(float(db * mx + np.log(np.mean(np.exp(db * (ss - mx))))))
        if hb.nan_flag():
            raise FloatingPointError("an NA or NaN somewhere! (device NaN flag raised)")
        hb.close()
        return float(np.sum(terms)), terms, ssbeta

    def allocation_trace(self):
        """G_shard x (K+1) counts, R/zigzagMCMC.R:194-199."""
        return self.h.get_alloc_trace(0).T.astype(np.float64)

    def prob_active(self):
        tr = self.allocation_trace()
        return 1 - np.round(tr[:, 0] / tr[0].sum(), 3)                                   # FileWriting:116-118

    # ------------------------------------------------------------------ output files (R/zigzagFileWriting.R, formats kept)
    def _path(self, prefix, suffix):
        return os.path.join(self.output_directory, prefix + suffix)

    def _write_hyperparameter_settings(self):                                            # I:509-536 (shared_variance branch)
        os.makedirs(self.output_directory, exist_ok=True)
        p = self.pri
        rows = [("s0_mu", p["s0_mu"]), ("s0_sigma", p["s0_sigma"]), ("s1_shape", p["s1_shape"]), ("s1_rate", p["s1_rate"]),
                ("tau_rate", p["tau_rate"]), ("tau_shape", p["tau_shape"]), ("alpha_r_shape", p["alpha_r_shape"]), ("alpha_r_rate", p["alpha_r_rate"]),
                ("weight_active_shape_1", p["weight_active_shape_1"]), ("weight_active_shape_2", p["weight_active_shape_2"]),
                ("weight_within_active_alpha", self.weight_within_active_alpha[0]), ("spike_prior_shape_1", p["spike_prior_shape_1"]),
                ("spike_prior_shape_2", p["spike_prior_shape_2"]), ("active_means_dif_prior_shape", p["active_means_dif_prior_shape"]),
                ("active_meaans_dif_prior_rate", round(p["active_means_dif_prior_rate"], 3)), ("inactive_means_prior_shape", p["inactive_means_prior_shape"]),
                ("inactive_means_prior_rate", round(p["inactive_means_prior_rate"], 3)), ("shared_variance_prior_min", p["shared_variance_prior_min"]),
                ("shared_variance_prior_max", p["shared_variance_prior_max"]), ("threshold_i", round(self.threshold_i, 2)),
                ("threshold_a", ", ".join(_rnum(round(t, 2)) for t in self.threshold_a))]
        if not self.shared_variance:                                                     # I:522-536: the bounds of the two variance priors instead
            i_act = [k for k, _ in rows].index("active_meaans_dif_prior_rate") + 1
            rows[i_act:i_act] = [("active_variances_prior_min", p["active_variances_prior_min"]), ("active_variances_prior_max", p["active_variances_prior_max"])]
            i_sh = [k for k, _ in rows].index("shared_variance_prior_min")
            rows[i_sh:i_sh + 2] = [("inactive_variances_prior_min", p["inactive_variances_prior_min"]), ("inactive_variances_prior_max", p["inactive_variances_prior_max"])]
        with open(os.path.join(self.output_directory, "hyperparameter_settings.txt"), "w") as f:
            for k, v in rows:
                f.write(f"{k}\t{v if isinstance(v, str) else _rnum(v)}\n")

    def initializeOutputFiles(self, prefix, post_pred=False):                            # FileWriting:3-58
        K = self.num_active_components
        if post_pred:                                                                    # FileWriting:44-55
            R = self.num_libraries
            with open(self._path(prefix, ".post_predictive_output.log"), "w") as f:
                f.write("\t".join(["gen", "Wass_Lower", "Wass_Upper", "Rums"]) + "\n")
            with open(self._path(prefix, ".library_specific_post_predictive_output.log"), "w") as f:
                f.write("\t".join(["gen"] + [f"Scaled_Wass_Lower_library_{r + 1}" for r in range(R)] + [f"Rums_library_{r + 1}" for r in range(R)]) + "\n")
        cols = ["gen", "lnl", "s0", "s1", "tau"] + [f"alpha_r_library_{r + 1}" for r in range(self.num_libraries)] + \
            ["spike_probability", "inactive_mean", "inactive_variance"] + [f"active_mean_component{k + 1}" for k in range(K)] + \
            [f"active_variance_component{k + 1}" for k in range(K)] + ["weight_active"] + [f"weight_within_active_component_{k + 1}" for k in range(K)]
        with open(self._path(prefix, "_model_parameters.log"), "w") as f:
            f.write("\t".join(cols) + "\n")
        for suf in ("_yg_candidate_genes.log", "_varianceg_candidate_genes.log"):
            with open(self._path(prefix, suf), "w") as f:
                f.write("\t".join(["gen"] + list(self.candidate_gene_list)) + "\n")

    def posteriorPredictiveSimulation(self, prefix=None, gen=None):                      # PostPred:3-90
        """Simulates one data set from the current state on the device and returns / logs the realised discrepancy statistics
        (lower- and upper-level Wasserstein distances, Rumsfeld zero-fraction discrepancy, and their per-library versions)."""
        from statistics import NormalDist
        K = self.num_active_components
        st = self.state()
        yo = st["Yg"][st["inactive_spike_allocation"] == 0]
        xf = self.Xg[np.isfinite(self.Xg)]
        y_lo, y_hi = (yo.min(), yo.max()) if len(yo) else (math.inf, -math.inf)
        if self.world > 1:                                                               # the range of Yg over ALL shards
            y_lo, y_hi = -self._allreduce([-y_lo], "max")[0], self._allreduce([y_hi], "max")[0]
        rng_ = [NormalDist(self.inactive_means, math.sqrt(self.inactive_variances)).inv_cdf(0.000001),
                NormalDist(self.active_means[K - 1], math.sqrt(self.active_variances[K - 1])).inv_cdf(0.999999),
                y_lo, y_hi, xf.min(), xf.max()]                                          # PostPred:9-15
        bins = np.arange(min(rng_) - 2, max(rng_) + 2, 0.1)                              # PostPred:17 (binwidth 0.1)
        # gene shards: the histograms are summed over the shards inside the call (zz_post_predictive_sharded), every rank gets the statistics
        out = self.h.post_predictive(self._next_step(), bins, allreduce=self._allreduce_inplace if self.world > 1 else None)
        if prefix is not None and self.rank == 0:
            g = self.gen if gen is None else gen
            with open(self._path(prefix, ".post_predictive_output.log"), "a") as f:      # PostPred:82-84
                f.write("\t".join(_rnum(x) for x in (g, out["W_L"], out["W_U"], out["B"])) + "\n")
            with open(self._path(prefix, ".library_specific_post_predictive_output.log"), "a") as f:   # PostPred:86-88
                f.write("\t".join(_rnum(x) for x in (g, *out["W_L_r"], *out["B_r"])) + "\n")
        return out

    def _write_sample(self, prefix, gen, yv_row):
        """The rows of one sampling point.  Called on EVERY rank with identical arguments: the candidate-gene rows are gathered
        over the gene shards (a collective); only rank 0 touches the files."""
        if self.rank == 0:
            self.writeToOutputFiles(prefix, gen)
        if yv_row:
            self.writeToYgVariancegOutputFiles(prefix, gen)

    def writeToOutputFiles(self, prefix, gen):                                           # FileWriting:60-86
        row = [self.lnl_trace[-1], self.s0, self.s1, self.tau, *self.alpha_r, self.spike_probability, self.inactive_means,
               self.inactive_variances, *self.active_means, *self.active_variances, self.weight_active, *self.weight_within_active]
        with open(self._path(prefix, "_model_parameters.log"), "a") as f:
            f.write("\t".join([_rnum(gen)] + [_rnum(round(float(x), 4)) for x in row]) + "\n")

    def _gather_genes(self, v):
        if self.world == 1:
            return v
        full = np.zeros(self.num_transcripts)
        full[self.g0:self.g1] = v
        return self._allreduce(full)

    def writeToYgVariancegOutputFiles(self, prefix, gen):                                # FileWriting:88-109
        s = self.state()
        sp = s["inactive_spike_allocation"]
        ge = self._gather_genes(s["Yg"] * (1 - sp) + -10.0 * sp)[self._cand_idx]
        gv = self._gather_genes(s["variance_g"] * (1 - sp))[self._cand_idx]
        if self.rank != 0:
            return
        with open(self._path(prefix, "_yg_candidate_genes.log"), "a") as f:
            f.write("\t".join([_rnum(gen)] + [_rnum(round(float(x), 3)) for x in ge]) + "\n")
        with open(self._path(prefix, "_varianceg_candidate_genes.log"), "a") as f:
            f.write("\t".join([_rnum(gen)] + [_rnum(round(float(x), 4)) for x in gv]) + "\n")

    def computeGeneExpressionProbs_writeToFile(self, prefix):                            # FileWriting:111-128
        tr = self.allocation_trace()
        if self.world > 1:
            full = np.zeros((self.num_transcripts, tr.shape[1])); full[self.g0:self.g1] = tr
            tr = self._allreduce(full.ravel()).reshape(full.shape)
        if self.rank != 0:
            return
        probs = np.round(tr / tr[0].sum(), 3)
        with open(self._path(prefix, "_probability_active.tab"), "w") as f:
            f.write("gene\tprob_active\n")
            for g, p in zip(self.gene_names, 1 - probs[:, 0]):
                f.write(f"{g}\t{_rfmt3(p)}\n")
        with open(self._path(prefix, "_probability_in_component.tab"), "w") as f:
            f.write("gene\t" + "\t".join(f"prob_{k}" for k in range(tr.shape[1])) + "\n")
            for g, row in zip(self.gene_names, probs):
                f.write(g + "\t" + "\t".join(_rfmt3(p) for p in row) + "\n")


def _rnum(x):
    """R's default number -> text conversion (15 significant digits, integers without a decimal point)."""
    x = float(x)
    if x == int(x) and abs(x) < 1e15:
        return str(int(x))
    return repr(float(f"{x:.15g}"))


def _rfmt3(p):
    """format(x, digits = 3) on an already rounded probability."""
    return f"{float(p):.3f}"
