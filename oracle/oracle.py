"""ctypes loader for the C oracle (oracle/zz_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; the product package zigzag_b200/ never does.  See oracle/zz_oracle.h for the pinning statement.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
MAX_LIBS, MAX_COMP = 64, 8


class Hyper(C.Structure):
    """Mirror of zzo_hyper (same field order as include/zigzag_gpu.h:zz_hyper)."""
    _fields_ = [("num_libraries", C.c_int32), ("num_active_components", C.c_int32),
                ("s0", C.c_double), ("s1", C.c_double), ("tau", C.c_double),
                ("alpha_r", C.c_double * MAX_LIBS),
                ("spike_probability", C.c_double),
                ("inactive_means", C.c_double), ("inactive_variances", C.c_double),
                ("active_means", C.c_double * MAX_COMP), ("active_variances", C.c_double * MAX_COMP),
                ("weight_active", C.c_double), ("weight_within_active", C.c_double * MAX_COMP),
                ("beta", C.c_double), ("temperature", C.c_double), ("variance_g_upper_bound", C.c_double)]

    @classmethod
    def make(cls, alpha_r, active_means, active_variances, weight_within_active, **kw):
        h = cls()
        h.num_libraries = len(alpha_r)
        h.num_active_components = len(active_means)
        for i, a in enumerate(alpha_r):
            h.alpha_r[i] = a
        for i in range(len(active_means)):
            h.active_means[i] = active_means[i]
            h.active_variances[i] = active_variances[i]
            h.weight_within_active[i] = weight_within_active[i]
        h.beta, h.temperature, h.variance_g_upper_bound = 1.0, 1.0, float("inf")
        for k, v in kw.items():
            setattr(h, k, v)
        return h

    def copy(self):
        h = Hyper()
        C.memmove(C.byref(h), C.byref(self), C.sizeof(Hyper))
        return h


class _State(C.Structure):
    _fields_ = [("G", C.c_int64), ("Xg", C.c_void_p), ("gene_lengths", C.c_void_p),
                ("Yg", C.c_void_p), ("variance_g", C.c_void_p),
                ("alloc_active_inactive", C.c_void_p), ("alloc_within_active", C.c_void_p),
                ("inactive_spike_allocation", C.c_void_p), ("no_detect_spike", C.c_void_p),
                ("pinned_active", C.c_void_p),
                ("XgLikelihood", C.c_void_p), ("variance_g_probability", C.c_void_p),
                ("tuningParam_yg", C.c_void_p), ("tuningParam_variance_g", C.c_void_p),
                ("Yg_trace", C.c_void_p), ("variance_g_trace", C.c_void_p)]


class Priors(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "weight_active_shape_1", "weight_active_shape_2", "spike_prior_shape_1", "spike_prior_shape_2",
        "active_means_dif_prior_shape", "active_means_dif_prior_rate",
        "inactive_means_prior_shape", "inactive_means_prior_rate",
        "variances_prior_log_min", "variances_prior_log_max",
        "tau_shape", "tau_rate", "s0_mu", "s0_sigma", "s1_shape", "s1_rate", "alpha_r_shape", "alpha_r_rate",
        "threshold_i")] + [("threshold_a", C.c_double * MAX_COMP)]

    @classmethod
    def defaults(cls, threshold_i, threshold_a):
        """Constructor defaults, R/zigzagInitialize.R:7-41."""
        p = cls(2, 2, 1, 1, 1, 1 / 3, 1, 1 / 3, np.log10(0.01), np.log10(5), 1, 1, -1, 2, 1, 2, 1, 1 / 10, threshold_i)
        for i, t in enumerate(threshold_a):
            p.threshold_a[i] = t
        return p


def build(force=False):
    so = os.path.join(_HERE, "libzzoracle.so")
    src = os.path.join(_HERE, "zz_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


_libs = {}


def lib(omp=False):
    key = bool(omp)
    if key not in _libs:
        build()
        L = C.CDLL(os.path.join(_HERE, "libzzoracle_omp.so" if omp else "libzzoracle.so"))
        L.zzo_varg_hyper_sum.restype = C.c_double
        L.zzo_lnl.restype = C.c_double
        L.zzo_x_tune.restype = C.c_double
        L.zzo_x_tune.argtypes = [C.c_double] * 5
        L.zzo_varg_hyper_sum.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double]
        L.zzo_level2_sums.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        L.zzo_chain_new.restype = C.c_void_p
        L.zzo_chain_new.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        for f in ("zzo_chain_free", "zzo_chain_burnin", "zzo_chain_mcmc", "zzo_chain_run_fused", "zzo_chain_prob_active"):
            getattr(L, f).restype = None
        L.zzo_chain_free.argtypes = [C.c_void_p]
        L.zzo_chain_burnin.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.zzo_chain_mcmc.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.zzo_chain_run_fused.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double]
        L.zzo_chain_prob_active.argtypes = [C.c_void_p, C.c_void_p]
        L.zzo_chain_hyper.restype = C.POINTER(Hyper)
        L.zzo_chain_hyper.argtypes = [C.c_void_p]
        L.zzo_chain_lnl.restype = C.c_double
        L.zzo_chain_lnl.argtypes = [C.c_void_p]
        L.zzo_uniform4.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint64, C.c_void_p]
        L.zzo_sweep_draws.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, C.c_int64, C.c_void_p]
        _libs[key] = L
    return _libs[key]


def set_threads(n, omp=True):
    """Thread count of the OpenMP build (returns the count in effect)."""
    return int(lib(omp).zzo_set_threads(int(n)))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class State:
    """Per-gene chain state as numpy arrays + the zzo_state view the C functions read/write in place."""
    F64 = ("Yg", "variance_g", "XgLikelihood", "variance_g_probability", "tuningParam_yg", "tuningParam_variance_g")
    I32 = ("alloc_active_inactive", "alloc_within_active", "inactive_spike_allocation")

    def __init__(self, Xg, gene_lengths, Yg, variance_g, alloc_active_inactive, alloc_within_active,
                 inactive_spike_allocation, XgLikelihood=None, variance_g_probability=None,
                 tuningParam_yg=None, tuningParam_variance_g=None, pinned_active=None):
        Xg = np.asarray(Xg, dtype=np.float64)
        self.G, self.R = Xg.shape  # Xg given as G x R (like R's matrix)
        self.Xg = np.asfortranarray(Xg)  # column-major == library-major SoA
        self.gene_lengths = np.ascontiguousarray(gene_lengths, dtype=np.float64)
        G = self.G
        self.Yg = np.array(Yg, dtype=np.float64)
        self.variance_g = np.array(variance_g, dtype=np.float64)
        self.alloc_active_inactive = np.array(alloc_active_inactive, dtype=np.int32)
        self.alloc_within_active = np.array(alloc_within_active, dtype=np.int32)
        self.inactive_spike_allocation = np.array(inactive_spike_allocation, dtype=np.int32)
        self.no_detect_spike = np.ascontiguousarray((np.isneginf(self.Xg).sum(axis=1) == self.R).astype(np.int32))
        self.pinned_active = None if pinned_active is None else np.array(pinned_active, dtype=np.int32)
        self.XgLikelihood = np.zeros(G) if XgLikelihood is None else np.array(XgLikelihood, dtype=np.float64)
        self.variance_g_probability = np.zeros(G) if variance_g_probability is None else np.array(variance_g_probability, dtype=np.float64)
        self.tuningParam_yg = np.full(G, 0.5) if tuningParam_yg is None else np.array(tuningParam_yg, dtype=np.float64)
        self.tuningParam_variance_g = np.full(G, 0.5) if tuningParam_variance_g is None else np.array(tuningParam_variance_g, dtype=np.float64)
        self.Yg_trace = np.zeros(2 * G, dtype=np.uint64)
        self.variance_g_trace = np.zeros(2 * G, dtype=np.uint64)
        lib().zzo_window_init(_p(self.Yg_trace), C.c_int64(G))
        lib().zzo_window_init(_p(self.variance_g_trace), C.c_int64(G))

    def c(self):
        s = _State()
        s.G = self.G
        for n, _ in _State._fields_[1:]:
            a = getattr(self, n)
            setattr(s, n, None if a is None else a.ctypes.data)
        self._keep = s
        return C.byref(s)

    def copy(self):
        o = State.__new__(State)
        o.__dict__.update({k: (v.copy(order="K") if isinstance(v, np.ndarray) else v) for k, v in self.__dict__.items() if k != "_keep"})
        return o

    def window_sums(self, which="Yg_trace"):
        w = getattr(self, which).reshape(-1, 2)
        pc = np.vectorize(lambda x: bin(int(x)).count("1"))
        return pc(w[:, 0]) + pc(w[:, 1] & np.uint64((1 << 36) - 1))


# ---- thin functional wrappers (all operate in place on State) ----

def refresh_caches(h, s, omp=False):
    lib(omp).zzo_refresh_caches(C.byref(h), s.c())


def eval_xg_loglik(h, s, omp=False):
    out = np.empty(s.G); lib(omp).zzo_eval_xg_loglik(C.byref(h), s.c(), _p(out)); return out


def eval_varg_logprior(h, s):
    out = np.empty(s.G); lib().zzo_eval_varg_logprior(C.byref(h), s.c(), _p(out)); return out


def eval_level2_loglik(h, s):
    out = np.empty(s.G); lib().zzo_eval_level2_loglik(C.byref(h), s.c(), _p(out)); return out


def mh_yg(h, s, z, u, tune=False):
    dec = np.zeros(s.G, np.uint8); lr = np.empty(s.G)
    lib().zzo_mh_yg(C.byref(h), s.c(), _p(np.ascontiguousarray(z)), _p(np.ascontiguousarray(u)), int(tune), _p(dec), _p(lr))
    return dec, lr


def mh_variance_g(h, s, u1, u2, tune=False):
    dec = np.zeros(s.G, np.uint8); lr = np.empty(s.G)
    lib().zzo_mh_variance_g(C.byref(h), s.c(), _p(np.ascontiguousarray(u1)), _p(np.ascontiguousarray(u2)), int(tune), _p(dec), _p(lr))
    return dec, lr


def gibbs_alloc_active_inactive(h, s, u_ai, u_within):
    dec = np.zeros(s.G, np.uint8); pa = np.empty(s.G)
    lib().zzo_gibbs_alloc_active_inactive(C.byref(h), s.c(), _p(np.ascontiguousarray(u_ai)), _p(np.ascontiguousarray(u_within)), _p(dec), _p(pa))
    return dec, pa


def gibbs_alloc_within_active(h, s, u_within):
    dec = np.zeros(s.G, np.uint8)
    lib().zzo_gibbs_alloc_within_active(C.byref(h), s.c(), _p(np.ascontiguousarray(u_within)), _p(dec))
    return dec


def mh_spike_alloc(h, s, z_y, z_v, u):
    dec = np.zeros(s.G, np.uint8); lr = np.empty(s.G)
    lib().zzo_mh_spike_alloc(C.byref(h), s.c(), _p(np.ascontiguousarray(z_y)), _p(np.ascontiguousarray(z_v)), _p(np.ascontiguousarray(u)), _p(dec), _p(lr))
    return dec, lr


def sweep(h, s, draws, tune=False, omp=False, want_decisions=True):
    draws = np.ascontiguousarray(draws, dtype=np.float64)
    assert draws.shape == (9, s.G)
    dec = np.zeros((4, s.G), np.uint8) if want_decisions else None
    lib(omp).zzo_sweep(C.byref(h), s.c(), _p(draws), int(tune), _p(dec))
    return dec


def suffstats(h, s):
    n = 3 * (h.num_active_components + 1) + 10
    out = np.empty(n); lib().zzo_suffstats(C.byref(h), s.c(), _p(out)); return out


def varg_hyper_sum(h, s, s0, s1, tau):
    return lib().zzo_varg_hyper_sum(C.byref(h), s.c(), s0, s1, tau)


def level2_sums(h, s, im, iv, am, av):
    out = np.empty(2)
    am = np.ascontiguousarray(am, dtype=np.float64); av = np.ascontiguousarray(av, dtype=np.float64)
    lib().zzo_level2_sums(C.byref(h), s.c(), im, iv, _p(am), _p(av), _p(out)); return out


def px_delta(h, s, alpha_prop, lib_mask=None):
    out = np.zeros(h.num_libraries)
    ap = np.ascontiguousarray(alpha_prop, dtype=np.float64)
    m = None if lib_mask is None else np.ascontiguousarray(lib_mask, dtype=np.uint8)
    lib().zzo_px_delta(C.byref(h), s.c(), _p(ap), _p(m), _p(out)); return out


def px_commit(h, s, lib_idx, alpha_new):
    lib().zzo_px_commit(C.byref(h), s.c(), int(lib_idx), C.c_double(alpha_new))


def lnl(h, s):
    return lib().zzo_lnl(C.byref(h), s.c())


def tune_per_gene(s, target=0.44):
    lib().zzo_tune_per_gene(s.c(), C.c_double(target))


def init_state(h, s, seed, chain, step, gene0, rv_m, rv_s):
    """zzo_init_state: the constructor-side initial state (I:416-504) written into State `s` in place."""
    L = lib()
    L.zzo_init_state.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, C.c_double, C.c_double, C.c_void_p]
    L.zzo_init_state.restype = None
    L.zzo_init_state(C.byref(h), s.c(), seed, chain, step, gene0, rv_m, rv_s, None)


def rowvar_moments(X):
    """(log(mean(v)), sd(v)) of the row variances v of the fully detected genes (I:446)."""
    X = np.asarray(X)
    full = ~np.isneginf(X).any(1)
    v = X[full].var(axis=1, ddof=1)
    return float(np.log(v.mean())), float(np.sqrt(v.var(ddof=1)))


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
    lib().zzo_philox4x32_10(c, k, o); return list(o)


def uniform4(seed, chain, step, blk, gene):
    o = (C.c_double * 4)(); lib().zzo_uniform4(seed, chain, step, blk, gene, o); return tuple(o)


def sweep_draws(seed, chain, step, gene0, G):
    d = np.empty((9, G)); lib().zzo_sweep_draws(seed, chain, step, gene0, G, _p(d)); return d


def post_predictive(h, s, seed, step, bins):
    """zzo_post_predictive: (W_L, W_U, B, W_L_r[R], B_r[R]) of one simulated data set."""
    bins = np.ascontiguousarray(bins, dtype=np.float64)
    R = h.num_libraries
    out = np.empty(3 + 2 * R)
    L = lib()
    L.zzo_post_predictive.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_int, C.c_void_p]
    L.zzo_post_predictive.restype = None
    L.zzo_post_predictive(C.byref(h), s.c(), seed, step, _p(bins), len(bins), _p(out))
    return dict(W_L=out[0], W_U=out[1], B=out[2], W_L_r=out[3:3 + R].copy(), B_r=out[3 + R:].copy())


class ChainScalars(C.Structure):
    """zzo_chain_scalars == zz_chain_scalars (include/zigzag_gpu.h): scalar tuning parameters, their windows, stream positions."""
    _fields_ = [("active_means_dif", C.c_double * MAX_COMP),
                ("t_im", C.c_double), ("t_av", C.c_double), ("t_am", C.c_double * MAX_COMP), ("t_s0", C.c_double),
                ("t_s1", C.c_double), ("t_tau", C.c_double), ("t_s0tau", C.c_double), ("t_alpha", C.c_double * MAX_LIBS),
                ("w_im", C.c_uint8 * 100), ("w_av", C.c_uint8 * 100), ("w_am", (C.c_uint8 * 100) * MAX_COMP),
                ("w_s0", C.c_uint8 * 100), ("w_s1", C.c_uint8 * 100), ("w_tau", C.c_uint8 * 100), ("w_s0tau", C.c_uint8 * 100),
                ("w_alpha", (C.c_uint8 * 100) * MAX_LIBS),
                ("hyper_ctr", C.c_uint64), ("hyper_buf", C.c_double * 4), ("step", C.c_uint64), ("gen", C.c_int64),
                ("n_samples", C.c_int64)]


class Chain:
    """Schedule-faithful host chain (zzo_chain_*): restates burnin()/mcmc() with the reference's 15-move schedule."""

    def __init__(self, Xg, gene_lengths, priors, K, seed, omp=False):
        Xg = np.asfortranarray(Xg, dtype=np.float64)
        self.G, self.R = Xg.shape
        self.L = lib(omp)
        gl = np.ascontiguousarray(gene_lengths, dtype=np.float64)
        self.ptr = self.L.zzo_chain_new(self.G, self.R, K, _p(Xg), _p(gl), C.byref(priors), seed)

    def burnin(self, ngen, target=0.44):
        self.L.zzo_chain_burnin(self.ptr, ngen, target)

    def mcmc(self, ngen, sample_frequency):
        self.L.zzo_chain_mcmc(self.ptr, ngen, sample_frequency)

    def run_fused(self, ngen, sample_frequency=1, tune=False, target=0.44):
        self.L.zzo_chain_run_fused(self.ptr, ngen, sample_frequency, int(tune), target)

    def prob_active(self):
        out = np.empty(self.G); self.L.zzo_chain_prob_active(self.ptr, _p(out)); return out

    @property
    def hyper(self):
        return self.L.zzo_chain_hyper(self.ptr).contents

    def lnl(self):
        return self.L.zzo_chain_lnl(self.ptr)

    def set_stream(self, seed, chain_index):
        self.L.zzo_chain_set_stream.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32]
        self.L.zzo_chain_set_stream.restype = None
        self.L.zzo_chain_set_stream(self.ptr, seed, chain_index)

    def scalars(self):
        out = ChainScalars()
        self.L.zzo_chain_export_scalars.argtypes = [C.c_void_p, C.c_void_p]
        self.L.zzo_chain_export_scalars.restype = None
        self.L.zzo_chain_export_scalars(self.ptr, C.byref(out))
        return out

    def state_arrays(self):
        """Copies of the chain's per-gene state (dict of numpy arrays, the names of State)."""
        self.L.zzo_chain_state.restype = C.POINTER(_State)
        self.L.zzo_chain_state.argtypes = [C.c_void_p]
        st = self.L.zzo_chain_state(self.ptr).contents
        G = self.G
        out = {}
        for n in State.F64:
            out[n] = np.ctypeslib.as_array(C.cast(getattr(st, n), C.POINTER(C.c_double)), (G,)).copy()
        for n in State.I32:
            out[n] = np.ctypeslib.as_array(C.cast(getattr(st, n), C.POINTER(C.c_int32)), (G,)).copy()
        for n in ("Yg_trace", "variance_g_trace"):
            out[n] = np.ctypeslib.as_array(C.cast(getattr(st, n), C.POINTER(C.c_uint64)), (2 * G,)).copy()
        return out

    def alloc_trace(self, K):
        self.L.zzo_chain_alloc_trace.restype = C.POINTER(C.c_uint32)
        self.L.zzo_chain_alloc_trace.argtypes = [C.c_void_p]
        return np.ctypeslib.as_array(self.L.zzo_chain_alloc_trace(self.ptr), (K + 1, self.G)).copy()

    def __del__(self):
        if getattr(self, "ptr", None):
            self.L.zzo_chain_free(self.ptr); self.ptr = None
