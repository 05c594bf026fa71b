"""Generates tests/golden/chain_g400.npz: the hyper-parameters, stream positions and a state checksum of the oracle's fused-schedule
chain (oracle/zz_oracle.c: zzo_chain_run_fused) after fixed numbers of generations, so that an accidental change of a hyper move, of
the scalar Philox stream or of the schedule shows up in the CPU suite.  Run from the repo root:  python tests/golden/make_golden_chain.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle import oracle as O
from zigzag_b200 import synth


def run():
    G, R, K = 400, 4, 2
    X, gl, _ = synth.simulate(G, R, config_id=4242)
    ch = O.Chain(X, gl, O.Priors.defaults(-1.0, (1.0, 6.0)), K, 31)
    out = {}
    for tag, n, tune in (("burn", 420, True), ("samp", 30, False)):
        ch.run_fused(n, sample_frequency=3, tune=tune)
        hy, sc, st = ch.hyper, ch.scalars(), ch.state_arrays()
        out[tag + "_hyper"] = np.array([hy.s0, hy.s1, hy.tau, hy.spike_probability, hy.inactive_means, hy.inactive_variances, hy.weight_active,
                                        *hy.alpha_r[:R], *hy.active_means[:K], *hy.active_variances[:K], *hy.weight_within_active[:K]])
        out[tag + "_ctr"] = np.array([sc.hyper_ctr, sc.gen, sc.step, sc.n_samples], dtype=np.int64)
        out[tag + "_tuning"] = np.array([sc.t_im, sc.t_av, sc.t_s0, sc.t_s1, sc.t_tau, sc.t_s0tau, *sc.t_am[:K], *sc.t_alpha[:R]])
        out[tag + "_state"] = np.array([st["Yg"].sum(), st["variance_g"].sum(), st["alloc_active_inactive"].sum(), st["inactive_spike_allocation"].sum(),
                                        st["tuningParam_yg"].sum()])
    out["trace"] = ch.alloc_trace(K).sum(axis=1).astype(np.int64)
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "chain_g400.npz")
    np.savez(path, **run())
    print("wrote", path)
