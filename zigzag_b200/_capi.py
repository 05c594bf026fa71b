"""ctypes binding of include/zigzag_gpu.h — the same C ABI an R `.Call` shim binds.

There is no CPU fallback: importing this module without the built CUDA library raises, and creating a handle
without a CUDA device returns ZZ_ERR_CUDA which is raised as ZigzagGpuError.
"""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ZIGZAG_GPU_LIB", os.path.join(HERE, "libzigzag_gpu.so"))   # override: kernel-tuning builds only
MAX_LIBS, MAX_COMP = 64, 8

SLOT_YG, SLOT_VARIANCE_G, SLOT_ALLOC, SLOT_ALLOC_WITHIN, SLOT_SPIKE = 1, 2, 3, 4, 5
MOVE_YG, MOVE_VARIANCE_G, MOVE_ALLOC, MOVE_ALLOC_WITHIN, MOVE_SPIKE = 1, 2, 4, 8, 16
MOVE_REFERENCE_ORDER = 32
MOVE_FUSED = MOVE_YG | MOVE_VARIANCE_G | MOVE_ALLOC | MOVE_SPIKE


class ZigzagGpuError(RuntimeError):
    pass


class Hyper(C.Structure):
    """zz_hyper: the scalar RefClass fields the per-gene moves read (R/zigzag.R:117-338)."""
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
            h.alpha_r[i] = float(a)
        for i in range(len(active_means)):
            h.active_means[i] = float(active_means[i])
            h.active_variances[i] = float(active_variances[i])
            h.weight_within_active[i] = float(weight_within_active[i])
        h.beta, h.temperature, h.variance_g_upper_bound = 1.0, 1.0, float("inf")
        for k, v in kw.items():
            setattr(h, k, v)
        return h

    def copy(self):
        o = Hyper()
        C.memmove(C.byref(o), C.byref(self), C.sizeof(Hyper))
        return o


class Priors(C.Structure):
    """zz_priors: the prior hyper-parameters of zigzag$new (R/zigzagInitialize.R:7-41) read by the hyper moves."""
    _fields_ = [(n, C.c_double) for n in (
        "weight_active_shape_1", "weight_active_shape_2", "spike_prior_shape_1", "spike_prior_shape_2",
        "active_means_dif_prior_shape", "active_means_dif_prior_rate",
        "inactive_means_prior_shape", "inactive_means_prior_rate",
        "variances_prior_log_min", "variances_prior_log_max",
        "tau_shape", "tau_rate", "s0_mu", "s0_sigma", "s1_shape", "s1_rate", "alpha_r_shape", "alpha_r_rate",
        "threshold_i")] + [("threshold_a", C.c_double * MAX_COMP)]

    @classmethod
    def defaults(cls, threshold_i, threshold_a):
        """Constructor defaults of zigzag$new, R/zigzagInitialize.R:7-41."""
        p = cls(2, 2, 1, 1, 1, 1 / 3, 1, 1 / 3, np.log10(0.01), np.log10(5), 1, 1, -1, 2, 1, 2, 1, 1 / 10, threshold_i)
        for i, t in enumerate(threshold_a):
            p.threshold_a[i] = t
        return p


class ChainScalars(C.Structure):
    """zz_chain_scalars: scalar tuning parameters, their 100-wide windows and the stream positions of a device-resident chain."""
    _fields_ = [("active_means_dif", C.c_double * MAX_COMP),
                ("t_im", C.c_double), ("t_av", C.c_double), ("t_am", C.c_double * MAX_COMP), ("t_s0", C.c_double),
                ("t_s1", C.c_double), ("t_tau", C.c_double), ("t_s0tau", C.c_double), ("t_alpha", C.c_double * MAX_LIBS),
                ("w_im", C.c_uint8 * 100), ("w_av", C.c_uint8 * 100), ("w_am", (C.c_uint8 * 100) * MAX_COMP),
                ("w_s0", C.c_uint8 * 100), ("w_s1", C.c_uint8 * 100), ("w_tau", C.c_uint8 * 100), ("w_s0tau", C.c_uint8 * 100),
                ("w_alpha", (C.c_uint8 * 100) * MAX_LIBS),
                ("hyper_ctr", C.c_uint64), ("hyper_buf", C.c_double * 4), ("step", C.c_uint64), ("gen", C.c_int64),
                ("n_samples", C.c_int64)]

    @classmethod
    def initial(cls, active_means, threshold_a, num_libraries):
        """The scalar state right after zigzag$new: tuning parameters of I:185-199, windows c(rep(0,77), rep(1,23))."""
        s = cls()
        win = (C.c_uint8 * 100)(*([0] * 77 + [1] * 23))
        s.t_im = s.t_av = 0.5
        s.t_s0 = s.t_s1 = s.t_tau = s.t_s0tau = 0.05
        for name in ("w_im", "w_av", "w_s0", "w_s1", "w_tau", "w_s0tau"):
            setattr(s, name, win)
        for k, (m, t) in enumerate(zip(active_means, threshold_a)):
            s.active_means_dif[k] = float(m) - float(t)
            s.t_am[k] = 0.5
            s.w_am[k] = win
        for r in range(num_libraries):
            s.t_alpha[r] = 0.5
            s.w_alpha[r] = win
        return s


class SampleSink(C.Structure):
    """zz_sample_sink: rings of pinned host memory the device appends the sample-time log rows to."""
    _fields_ = [("param_rows", C.c_void_p), ("param_capacity", C.c_int64), ("yg_rows", C.c_void_p), ("vg_rows", C.c_void_p),
                ("yv_capacity", C.c_int64), ("candidates", C.c_void_p), ("n_candidates", C.c_int64), ("yv_every", C.c_int32)]


class Config(C.Structure):
    _fields_ = [("num_genes", C.c_int64), ("gene_offset", C.c_int64), ("num_libraries", C.c_int32),
                ("num_active_components", C.c_int32), ("num_chains", C.c_int32), ("device", C.c_int32),
                ("seed", C.c_uint64)]


# every symbol include/zigzag_gpu.h declares: (name, restype, argtypes)
_P = C.c_void_p
_SIG = {
    "zz_create": (C.c_int, [C.POINTER(Config), C.POINTER(_P)]),
    "zz_destroy": (None, [_P]),
    "zz_last_error": (C.c_char_p, [_P]),
    "zz_set_stream": (C.c_int, [_P, _P]),
    "zz_sync": (C.c_int, [_P]),
    "zz_launch_count": (C.c_uint64, [_P]),
    "zz_upload_data": (C.c_int, [_P, _P, _P]),
    "zz_set_state": (C.c_int, [_P, C.c_int] + [_P] * 10),
    "zz_get_state": (C.c_int, [_P, C.c_int] + [_P] * 9),
    "zz_set_hyper": (C.c_int, [_P, C.c_int, C.POINTER(Hyper)]),
    "zz_get_hyper": (C.c_int, [_P, C.c_int, C.POINTER(Hyper)]),
    "zz_refresh_caches": (C.c_int, [_P]),
    "zz_debug_refresh_lps": (C.c_int, [_P]),
    "zz_reset_windows": (C.c_int, [_P]),
    "zz_get_window_sums": (C.c_int, [_P, C.c_int, _P, _P]),
    "zz_eval_xg_loglik": (C.c_int, [_P, C.c_int, _P]),
    "zz_eval_varg_logprior": (C.c_int, [_P, C.c_int, _P]),
    "zz_eval_level2_loglik": (C.c_int, [_P, C.c_int, _P]),
    "zz_mh_yg": (C.c_int, [_P, C.c_uint64, C.c_int, _P, _P]),
    "zz_mh_variance_g": (C.c_int, [_P, C.c_uint64, C.c_int, _P, _P]),
    "zz_gibbs_alloc": (C.c_int, [_P, C.c_uint64, _P, _P]),
    "zz_gibbs_alloc_within": (C.c_int, [_P, C.c_uint64, _P, _P]),
    "zz_mh_spike_alloc": (C.c_int, [_P, C.c_uint64, _P, _P]),
    "zz_sweep": (C.c_int, [_P, C.c_uint64, C.c_int, C.c_uint32, _P, _P]),
    "zz_philox_draws": (C.c_int, [_P, C.c_uint64, C.c_int, _P]),
    "zz_suffstats_len": (C.c_int, [C.c_int]),
    "zz_suffstats": (C.c_int, [_P, _P]),
    "zz_suffstats_dev": (C.c_int, [_P, _P]),
    "zz_comm_handle_bytes": (C.c_int, []),
    "zz_comm_init": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "zz_comm_connect": (C.c_int, [_P, _P]),
    "zz_suffstats_allreduce_dev": (C.c_int, [_P, _P]),
    "zz_set_stats_output_dev": (C.c_int, [_P, _P, C.c_int]),
    "zz_varg_hyper_sum": (C.c_int, [_P, C.c_int, C.c_double, C.c_double, C.c_double, _P]),
    "zz_commit_varg_hyper": (C.c_int, [_P, C.c_int]),
    "zz_level2_sums": (C.c_int, [_P, C.c_int, C.c_double, C.c_double, _P, _P, _P]),
    "zz_px_delta": (C.c_int, [_P, C.c_int, _P, _P, _P]),
    "zz_px_commit": (C.c_int, [_P, C.c_int, C.c_int, C.c_double]),
    "zz_lnl": (C.c_int, [_P, C.c_int, _P]),
    "zz_reset_alloc_trace": (C.c_int, [_P]),
    "zz_accumulate_alloc_trace": (C.c_int, [_P]),
    "zz_get_alloc_trace": (C.c_int, [_P, C.c_int, _P]),
    "zz_accumulate_alloc_trace_from": (C.c_int, [_P, C.c_int, C.c_int]),
    "zz_copy_chain_tuning": (C.c_int, [_P, C.c_int, C.c_int]),
    "zz_init_state": (C.c_int, [_P, C.c_uint64, C.c_double, C.c_double]),
    "zz_tune": (C.c_int, [_P, C.c_double]),
    "zz_nan_flag": (C.c_int, [_P, _P]),
    "zz_sweep_host": (C.c_int, [_P, C.c_uint64, C.c_int] + [_P] * 11),
    "zz_chain_begin": (C.c_int, [_P, _P, _P]),
    "zz_run_generations": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_double]),
    "zz_run_sweeps": (C.c_int, [_P, C.c_uint64, C.c_int, C.c_int]),
    "zz_debug_gen_times": (C.c_int, [_P, _P]),
    "zz_chain_read": (C.c_int, [_P, _P, _P]),
    "zz_param_row_len": (C.c_int, [C.c_int, C.c_int]),
    "zz_pinned_alloc": (C.c_void_p, [C.c_size_t]),
    "zz_pinned_free": (None, [_P]),
    "zz_set_sample_sink": (C.c_int, [_P, _P]),
    "zz_sample_rows_written": (C.c_int, [_P, _P, _P]),
    "zz_post_predictive": (C.c_int, [_P, C.c_uint64, _P, C.c_int, _P]),
    "zz_post_predictive_sharded": (C.c_int, [_P, C.c_uint64, _P, C.c_int, _P, _P, _P]),
}
EXPORTS = tuple(_SIG)

_lib = None


def load():
    """Loads the CUDA library.  Raises if it has not been built — there is deliberately no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ZigzagGpuError(f"{LIB_PATH} is missing: build it with `python -m zigzag_b200.build` "
                                 "(zigzag_b200 has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIG.items():
            if "ZIGZAG_GPU_LIB" in os.environ and not hasattr(L, name):
                continue                                      # a kernel-tuning build of an older source tree (A/B timings only)
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _lib = L
    return _lib


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return a
    return a.ctypes.data


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


class Handle:
    """RAII wrapper of zz_handle.  All array arguments / results are numpy (host) arrays."""

    def __init__(self, num_genes, num_libraries, num_active_components, num_chains=1, device=0, seed=0, gene_offset=0):
        self.L = load()
        self.G, self.R, self.K, self.C = int(num_genes), int(num_libraries), int(num_active_components), int(num_chains)
        cfg = Config(self.G, int(gene_offset), self.R, self.K, self.C, int(device), int(seed))
        p = _P()
        rc = self.L.zz_create(C.byref(cfg), C.byref(p))
        self.h = p
        if rc != 0:
            msg = self.L.zz_last_error(p).decode() if p else "invalid configuration"
            if p:
                self.L.zz_destroy(p)
            self.h = None
            raise ZigzagGpuError(f"zz_create failed ({rc}): {msg}")

    def close(self):
        if getattr(self, "h", None):
            try:
                self.clear_sample_sink()
            except Exception:
                pass
            self.L.zz_destroy(self.h)
            self.h = None

    __del__ = close

    def _ck(self, rc):
        if rc != 0:
            raise ZigzagGpuError(f"zigzag_gpu error {rc}: {self.L.zz_last_error(self.h).decode()}")

    # ---- lifetime / plumbing
    def set_stream(self, cuda_stream):
        self._ck(self.L.zz_set_stream(self.h, _P(cuda_stream)))

    def sync(self):
        self._ck(self.L.zz_sync(self.h))

    @property
    def launch_count(self):
        return int(self.L.zz_launch_count(self.h))

    # ---- data / state
    def upload_data(self, Xg, gene_lengths):
        """Xg: G x R array (R's matrix); stored library-major on the device."""
        X = np.asfortranarray(Xg, dtype=np.float64)
        assert X.shape == (self.G, self.R), (X.shape, self.G, self.R)
        gl = _f64(gene_lengths)
        self._ck(self.L.zz_upload_data(self.h, _ptr(X), _ptr(gl)))

    def set_state(self, chain=0, Yg=None, variance_g=None, alloc_active_inactive=None, alloc_within_active=None,
                  inactive_spike_allocation=None, XgLikelihood=None, variance_g_probability=None, tuningParam_yg=None,
                  tuningParam_variance_g=None, pinned_active=None):
        a = [_f64(Yg), _f64(variance_g), _i32(alloc_active_inactive), _i32(alloc_within_active), _i32(inactive_spike_allocation),
             _f64(XgLikelihood), _f64(variance_g_probability), _f64(tuningParam_yg), _f64(tuningParam_variance_g), _i32(pinned_active)]
        for x in a:
            assert x is None or x.shape == (self.G,)
        self._ck(self.L.zz_set_state(self.h, chain, *[_ptr(x) for x in a]))

    def get_state(self, chain=0):
        G = self.G
        out = dict(Yg=np.empty(G), variance_g=np.empty(G), alloc_active_inactive=np.empty(G, np.int32),
                   alloc_within_active=np.empty(G, np.int32), inactive_spike_allocation=np.empty(G, np.int32),
                   XgLikelihood=np.empty(G), variance_g_probability=np.empty(G), tuningParam_yg=np.empty(G),
                   tuningParam_variance_g=np.empty(G))
        self._ck(self.L.zz_get_state(self.h, chain, *[_ptr(v) for v in out.values()]))
        return out

    def set_hyper(self, hyper, chain=0):
        self._ck(self.L.zz_set_hyper(self.h, chain, C.byref(hyper)))

    def get_hyper(self, chain=0):
        h = Hyper()
        self._ck(self.L.zz_get_hyper(self.h, chain, C.byref(h)))
        return h

    def refresh_caches(self):
        self._ck(self.L.zz_refresh_caches(self.h))

    def debug_refresh_lps(self):
        self._ck(self.L.zz_debug_refresh_lps(self.h))

    def reset_windows(self):
        self._ck(self.L.zz_reset_windows(self.h))

    def window_sums(self, chain=0):
        a, b = np.empty(self.G, np.int32), np.empty(self.G, np.int32)
        self._ck(self.L.zz_get_window_sums(self.h, chain, _ptr(a), _ptr(b)))
        return a, b

    # ---- evaluators
    def _eval(self, fn, chain):
        out = np.empty(self.G)
        self._ck(fn(self.h, chain, _ptr(out)))
        return out

    def eval_xg_loglik(self, chain=0):
        return self._eval(self.L.zz_eval_xg_loglik, chain)

    def eval_varg_logprior(self, chain=0):
        return self._eval(self.L.zz_eval_varg_logprior, chain)

    def eval_level2_loglik(self, chain=0):
        return self._eval(self.L.zz_eval_level2_loglik, chain)

    # ---- moves
    def _draws(self, draws, n):
        if draws is None:
            return None
        d = np.ascontiguousarray(draws, dtype=np.float64).reshape(n, self.C, self.G)
        return d

    def _move(self, fn, n, step, tune, draws, want_decisions, ndec=1):
        d = self._draws(draws, n)
        dec = np.zeros((ndec, self.C, self.G), np.uint8) if want_decisions else None
        if tune is None:
            self._ck(fn(self.h, step, _ptr(d), _ptr(dec)))
        else:
            self._ck(fn(self.h, step, int(tune), _ptr(d), _ptr(dec)))
        if dec is None:
            return None
        return dec[0, 0] if (ndec == 1 and self.C == 1) else (dec[0] if ndec == 1 else dec)

    def mh_yg(self, step, tune=False, draws=None, decisions=False):
        return self._move(self.L.zz_mh_yg, 2, step, tune, draws, decisions)

    def mh_variance_g(self, step, tune=False, draws=None, decisions=False):
        return self._move(self.L.zz_mh_variance_g, 2, step, tune, draws, decisions)

    def gibbs_alloc(self, step, draws=None, decisions=False):
        return self._move(self.L.zz_gibbs_alloc, 2, step, None, draws, decisions)

    def gibbs_alloc_within(self, step, draws=None, decisions=False):
        return self._move(self.L.zz_gibbs_alloc_within, 1, step, None, draws, decisions)

    def mh_spike_alloc(self, step, draws=None, decisions=False):
        return self._move(self.L.zz_mh_spike_alloc, 3, step, None, draws, decisions)

    def sweep(self, step, tune=False, moves=MOVE_FUSED, draws=None, decisions=False):
        d = self._draws(draws, 9)
        dec = np.zeros((4, self.C, self.G), np.uint8) if decisions else None
        self._ck(self.L.zz_sweep(self.h, step, int(tune), moves, _ptr(d), _ptr(dec)))
        return dec

    def philox_draws(self, step, slot):
        n = {SLOT_YG: 2, SLOT_VARIANCE_G: 2, SLOT_ALLOC: 2, SLOT_ALLOC_WITHIN: 1, SLOT_SPIKE: 3}[slot]
        out = np.empty((n, self.C, self.G))
        self._ck(self.L.zz_philox_draws(self.h, step, slot, _ptr(out)))
        return out

    # ---- reductions
    @property
    def nstat(self):
        return int(self.L.zz_suffstats_len(self.K))

    def suffstats(self):
        out = np.empty((self.C, self.nstat))
        self._ck(self.L.zz_suffstats(self.h, _ptr(out)))
        return out

    def suffstats_dev(self, dev_ptr):
        self._ck(self.L.zz_suffstats_dev(self.h, _P(dev_ptr)))

    def comm_init(self, rank, world):
        """Allocates this rank's peer-memory inbox; returns the CUDA IPC handle blob to all-gather."""
        buf = (C.c_ubyte * int(self.L.zz_comm_handle_bytes()))()
        self._ck(self.L.zz_comm_init(self.h, int(rank), int(world), C.addressof(buf)))
        return bytes(buf)

    def comm_connect(self, blobs):
        """blobs: the handle blobs of all ranks, in rank order."""
        cat = b"".join(blobs)
        buf = (C.c_ubyte * len(cat)).from_buffer_copy(cat)
        self._ck(self.L.zz_comm_connect(self.h, C.addressof(buf)))

    def set_stats_output_dev(self, dev_ptr, allreduce=False):
        self._ck(self.L.zz_set_stats_output_dev(self.h, _P(dev_ptr) if dev_ptr else None, int(allreduce)))

    def suffstats_allreduce_dev(self, dev_ptr):
        self._ck(self.L.zz_suffstats_allreduce_dev(self.h, _P(dev_ptr)))

    def varg_hyper_sum(self, s0, s1, tau, chain=0):
        out = np.empty(2)
        self._ck(self.L.zz_varg_hyper_sum(self.h, chain, s0, s1, tau, _ptr(out)))
        return out

    def commit_varg_hyper(self, chain=0):
        self._ck(self.L.zz_commit_varg_hyper(self.h, chain))

    def level2_sums(self, im, iv, am, av, chain=0):
        out = np.empty(2)
        am, av = _f64(am), _f64(av)
        self._ck(self.L.zz_level2_sums(self.h, chain, im, iv, _ptr(am), _ptr(av), _ptr(out)))
        return out

    def px_delta(self, alpha_prop, lib_mask=None, chain=0):
        out = np.zeros(self.R)
        ap = _f64(alpha_prop)
        m = None if lib_mask is None else np.ascontiguousarray(lib_mask, dtype=np.uint8)
        self._ck(self.L.zz_px_delta(self.h, chain, _ptr(ap), _ptr(m), _ptr(out)))
        return out

    def px_commit(self, lib, alpha_new, chain=0):
        self._ck(self.L.zz_px_commit(self.h, chain, int(lib), float(alpha_new)))

    def lnl(self, chain=0):
        out = np.empty(1)
        self._ck(self.L.zz_lnl(self.h, chain, _ptr(out)))
        return float(out[0])

    # ---- sampling / tuning
    def reset_alloc_trace(self):
        self._ck(self.L.zz_reset_alloc_trace(self.h))

    def accumulate_alloc_trace(self):
        self._ck(self.L.zz_accumulate_alloc_trace(self.h))

    def get_alloc_trace(self, chain=0):
        out = np.empty((self.K + 1, self.G), np.uint32)
        self._ck(self.L.zz_get_alloc_trace(self.h, chain, _ptr(out)))
        return out

    def accumulate_alloc_trace_from(self, src_chain, dst_chain=0):
        self._ck(self.L.zz_accumulate_alloc_trace_from(self.h, int(src_chain), int(dst_chain)))

    def copy_chain_tuning(self, src_chain, dst_chain):
        self._ck(self.L.zz_copy_chain_tuning(self.h, int(src_chain), int(dst_chain)))

    def init_state(self, step, rowvar_log_mean, rowvar_sd):
        """Constructor-side initial state on the device (R/zigzagInitialize.R:416-504), every chain from its own hyper-parameters."""
        self._ck(self.L.zz_init_state(self.h, int(step), float(rowvar_log_mean), float(rowvar_sd)))

    def tune(self, target=0.44):
        self._ck(self.L.zz_tune(self.h, float(target)))

    def nan_flag(self):
        f = C.c_int(0)
        self._ck(self.L.zz_nan_flag(self.h, C.byref(f)))
        return int(f.value)

    # ---- device-resident generation loop
    def chain_begin(self, priors, scalars):
        """priors: ctypes structure with the layout of zz_priors; scalars: one structure with the layout of zz_chain_scalars per
        chain of the handle (a single structure when the handle has one chain)."""
        if not isinstance(scalars, (list, tuple)):
            scalars = [scalars]
        assert len(scalars) == self.C, "one zz_chain_scalars per chain"
        arr = (ChainScalars * self.C)()
        for c, sc in enumerate(scalars):
            C.memmove(C.byref(arr, c * C.sizeof(ChainScalars)), C.addressof(sc), C.sizeof(ChainScalars))
        self._ck(self.L.zz_chain_begin(self.h, C.addressof(priors), C.addressof(arr)))

    def run_generations(self, ngen, tune=False, sample_frequency=1, target=0.44):
        self._ck(self.L.zz_run_generations(self.h, int(ngen), int(tune), int(sample_frequency), float(target)))

    def run_sweeps(self, step, nsweeps, tune=False):
        """`nsweeps` fused sweeps with fixed hyper-parameters in one cooperative launch (statistics after every sweep)."""
        self._ck(self.L.zz_run_sweeps(self.h, int(step), int(nsweeps), int(tune)))

    def debug_gen_times(self):
        out = np.zeros(8)
        self._ck(self.L.zz_debug_gen_times(self.h, _ptr(out)))
        return out

    def chain_read(self):
        """(scalars, hyper) of chain 0 for a one-chain handle, else two lists with one entry per chain."""
        sc, hy = (ChainScalars * self.C)(), (Hyper * self.C)()
        self._ck(self.L.zz_chain_read(self.h, C.addressof(sc), C.addressof(hy)))
        if self.C == 1:
            return sc[0], hy[0]
        return list(sc), list(hy)

    def post_predictive(self, step, bins, allreduce=None):
        """Realised discrepancies of one simulated data set: dict(W_L, W_U, B, W_L_r[R], B_r[R]) (PostPred:3-156).  On gene shards pass
        `allreduce(array)`: sums a numpy array in place over the shards (called with float64 / uint32 / uint64 arrays)."""
        bins = np.ascontiguousarray(bins, dtype=np.float64)
        out = np.empty(3 + 2 * self.R)
        if allreduce is None:
            self._ck(self.L.zz_post_predictive(self.h, step, _ptr(bins), len(bins), _ptr(out)))
        else:
            types = {0: (C.c_double, np.float64), 1: (C.c_uint32, np.uint32), 2: (C.c_uint64, np.uint64)}

            @C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int64, C.c_int)
            def cb(_ctx, buf, n, dtype):
                ct, _ = types[dtype]
                allreduce(np.ctypeslib.as_array(C.cast(buf, C.POINTER(ct)), (n,)))
            self._ck(self.L.zz_post_predictive_sharded(self.h, step, _ptr(bins), len(bins), C.cast(cb, C.c_void_p), None, _ptr(out)))
        return dict(W_L=out[0], W_U=out[1], B=out[2], W_L_r=out[3:3 + self.R].copy(), B_r=out[3 + self.R:].copy())

    # ---- sample-time rows written by the device
    def _pinned(self, rows, cols):
        p = self.L.zz_pinned_alloc(max(1, rows * cols) * 8)
        if not p:
            raise ZigzagGpuError("zz_pinned_alloc failed")
        self._sink_ptrs.append(p)
        return p, np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), (rows, cols))

    def set_sample_sink(self, param_capacity, candidates=None, yv_capacity=0, yv_every=1):
        """Allocates the pinned rings and attaches them; returns nothing, read with sample_rows()."""
        self.clear_sample_sink()
        self._sink_ptrs = []
        plen = int(self.L.zz_param_row_len(self.R, self.K))
        sk = SampleSink()
        sk.param_rows, self._sink_param = self._pinned(param_capacity, plen)
        sk.param_capacity = param_capacity
        self._sink_yg = self._sink_vg = None
        if candidates is not None and yv_capacity > 0:
            cand = np.ascontiguousarray(candidates, dtype=np.int32)
            sk.yg_rows, self._sink_yg = self._pinned(yv_capacity, 1 + len(cand))
            sk.vg_rows, self._sink_vg = self._pinned(yv_capacity, 1 + len(cand))
            sk.yv_capacity, sk.candidates, sk.n_candidates, sk.yv_every = yv_capacity, cand.ctypes.data, len(cand), int(yv_every)
        self._ck(self.L.zz_set_sample_sink(self.h, C.addressof(sk)))

    def sample_rows(self):
        """(parameter rows, Yg rows, variance_g rows) written so far (copies); waits for the device."""
        self.sync()
        a, b = C.c_int64(0), C.c_int64(0)
        self._ck(self.L.zz_sample_rows_written(self.h, C.addressof(a), C.addressof(b)))
        yg = None if self._sink_yg is None else self._sink_yg[:b.value].copy()
        vg = None if self._sink_vg is None else self._sink_vg[:b.value].copy()
        return self._sink_param[:a.value].copy(), yg, vg

    def clear_sample_sink(self):
        if getattr(self, "_sink_ptrs", None):
            self._ck(self.L.zz_set_sample_sink(self.h, None))
            for p in self._sink_ptrs:
                self.L.zz_pinned_free(p)
        self._sink_ptrs = []
        self._sink_param = self._sink_yg = self._sink_vg = None

    def sweep_host(self, step, tune, Yg, variance_g, ai, within, spike, XgLikelihood, variance_g_probability, ty, tv):
        """All arguments are host arrays or raw host pointers (ints), updated in place; returns (h2d_bytes, d2h_bytes)."""
        a, b = C.c_int64(0), C.c_int64(0)
        ptrs = [_ptr(x) for x in (Yg, variance_g, ai, within, spike, XgLikelihood, variance_g_probability, ty, tv)]
        self._ck(self.L.zz_sweep_host(self.h, step, int(tune), *ptrs, C.addressof(a), C.addressof(b)))
        return int(a.value), int(b.value)
