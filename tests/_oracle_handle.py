"""Test double of zigzag_b200._capi.Handle for the CPU suite: the same method surface the host mirror (zigzag_b200/zigzag.py)
drives in its host-driven schedules ("reference", "fused"), answered by the CPU oracle instead of the GPU.  It exists so that the
PRODUCT's sharded host logic — zigzag(shard=...), _allreduce, _gather_genes, the sample-time file writing — can run under gloo
with world_size 2 on a box without a GPU.  Test infrastructure only: nothing outside tests/ may import it."""
import ctypes as C

import numpy as np

from oracle import oracle as O


class OracleHandle:
    def __init__(self, num_genes, num_libraries, num_active_components, num_chains=1, device=0, seed=0, gene_offset=0):
        assert num_chains == 1
        self.G, self.R, self.K, self.C = int(num_genes), int(num_libraries), int(num_active_components), 1
        self.seed, self.gene_offset = int(seed), int(gene_offset)
        self.launch_count = 0
        self._nan = 0
        self.s = None
        self.ho = O.Hyper()
        self._trace = np.zeros((self.K + 1, self.G), np.uint32)

    # ---- data / state
    def upload_data(self, Xg, gene_lengths):
        self._X, self._gl = np.asfortranarray(Xg, dtype=np.float64), np.ascontiguousarray(gene_lengths, dtype=np.float64)

    def set_hyper(self, hyper, chain=0):
        C.memmove(C.byref(self.ho), C.byref(hyper), C.sizeof(O.Hyper))      # zz_hyper and zzo_hyper share their layout

    def set_state(self, chain=0, pinned_active=None, **st):
        if self.s is None:
            self.s = O.State(self._X, self._gl, st["Yg"], st["variance_g"], st["alloc_active_inactive"], st["alloc_within_active"],
                             st["inactive_spike_allocation"], tuningParam_yg=st.get("tuningParam_yg"),
                             tuningParam_variance_g=st.get("tuningParam_variance_g"), pinned_active=pinned_active)
        else:
            for k, v in st.items():
                if v is not None:
                    getattr(self.s, k)[:] = v

    def get_state(self, chain=0):
        s = self.s
        return {k: getattr(s, k).copy() for k in O.State.F64 + O.State.I32}

    def refresh_caches(self):
        O.refresh_caches(self.ho, self.s)

    # ---- moves: draws of the fused sweep's Philox map (rows: z_yg, u_yg, u1_vg, u2_vg, u_ai, u_within, z_y, z_v, u_spike)
    def _d(self, step):
        self.launch_count += 1
        return O.sweep_draws(self.seed, 0, int(step), self.gene_offset, self.G)

    def sweep(self, step, tune=False, **kw):
        O.sweep(self.ho, self.s, self._d(step), tune=tune, want_decisions=False)

    def mh_yg(self, step, tune=False, **kw):
        d = self._d(step); O.mh_yg(self.ho, self.s, d[0], d[1], tune)

    def mh_variance_g(self, step, tune=False, **kw):
        d = self._d(step); O.mh_variance_g(self.ho, self.s, d[2], d[3], tune)

    def gibbs_alloc(self, step, **kw):
        d = self._d(step); O.gibbs_alloc_active_inactive(self.ho, self.s, d[4], d[5])

    def gibbs_alloc_within(self, step, **kw):
        d = self._d(step); O.gibbs_alloc_within_active(self.ho, self.s, d[5])

    def mh_spike_alloc(self, step, **kw):
        d = self._d(step); O.mh_spike_alloc(self.ho, self.s, d[6], d[7], d[8])

    # ---- reductions
    @property
    def nstat(self):
        return 3 * (self.K + 1) + 10

    def suffstats(self):
        return O.suffstats(self.ho, self.s)[None, :]

    def varg_hyper_sum(self, s0, s1, tau, chain=0):
        return np.array([O.varg_hyper_sum(self.ho, self.s, s0, s1, tau), O.varg_hyper_sum(self.ho, self.s, self.ho.s0, self.ho.s1, self.ho.tau)])

    def commit_varg_hyper(self, chain=0):
        self.s.variance_g_probability[:] = np.where(self.s.inactive_spike_allocation == 0, O.eval_varg_logprior(self.ho, self.s),
                                                    self.s.variance_g_probability)

    def level2_sums(self, im, iv, am, av, chain=0):
        return O.level2_sums(self.ho, self.s, im, iv, am, av)

    def px_delta(self, alpha_prop, lib_mask=None, chain=0):
        return O.px_delta(self.ho, self.s, alpha_prop, lib_mask)

    def px_commit(self, lib, alpha_new, chain=0):
        O.px_commit(self.ho, self.s, lib, alpha_new)
        self.ho.alpha_r[lib] = alpha_new

    def lnl(self, chain=0):
        return float(O.lnl(self.ho, self.s))

    # ---- sampling / tuning
    def reset_alloc_trace(self):
        self._trace[:] = 0

    def accumulate_alloc_trace(self):
        comp = np.where(self.s.alloc_active_inactive == 1, self.s.alloc_within_active, 0)
        self._trace[comp, np.arange(self.G)] += 1

    def get_alloc_trace(self, chain=0):
        return self._trace.copy()

    def tune(self, target=0.44):
        O.tune_per_gene(self.s, target)

    def nan_flag(self):
        return self._nan

    def sync(self):
        pass

    def clear_sample_sink(self):
        pass

    def close(self):
        pass
