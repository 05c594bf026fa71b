"""Generates tests/golden/moves_g512.npz: inputs and the ORACLE's outputs of every per-gene routine on one seeded 512-gene
case (R = 4, K = 2, with the awkward rows of tests/_util.make_case).  The GPU tests compare against these committed numbers
as well as against the live oracle, so a silent change of the oracle itself is caught too.

  python tests/golden/make_golden_moves.py
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O                       # noqa: E402
from _util import make_case, hyper_pair, oracle_state  # noqa: E402


def main():
    G, R = 512, 4
    X, gl, hd, st = make_case(G, R, 77)
    ho, _ = hyper_pair(hd)
    s = oracle_state(X, gl, st)
    O.refresh_caches(ho, s)
    rng = np.random.Generator(np.random.PCG64(4242))
    d = rng.random((9, G))
    for j in (0, 6, 7):
        d[j] = rng.standard_normal(G)
    out = dict(X=X, gl=gl, draws=d, **{"st_" + k: v for k, v in st.items()}, **{"h_" + k: np.asarray(v, dtype=np.float64) for k, v in hd.items()},
               xglik=s.XgLikelihood.copy(), vgprob=s.variance_g_probability.copy(), level2=O.eval_level2_loglik(ho, s))
    s1 = s.copy(); out["dec_yg"], out["lr_yg"] = O.mh_yg(ho, s1, d[0], d[1]); out["yg_after"] = s1.Yg.copy()
    s2 = s.copy(); out["dec_vg"], out["lr_vg"] = O.mh_variance_g(ho, s2, d[2], d[3]); out["vg_after"] = s2.variance_g.copy()
    s3 = s.copy(); out["dec_alloc"], out["prob_active"] = O.gibbs_alloc_active_inactive(ho, s3, d[4], d[5])
    s4 = s.copy(); out["dec_spike"], out["lr_spike"] = O.mh_spike_alloc(ho, s4, d[6], d[7], d[8]); out["spike_after"] = s4.inactive_spike_allocation.copy()
    s5 = s.copy(); out["dec_sweep"] = O.sweep(ho, s5, d)
    out.update(sweep_yg=s5.Yg.copy(), sweep_vg=s5.variance_g.copy(), sweep_xglik=s5.XgLikelihood.copy(), sweep_vgprob=s5.variance_g_probability.copy(),
               sweep_stats=O.suffstats(ho, s5), sweep_lnl=np.array(O.lnl(ho, s5)))
    np.savez_compressed(os.path.join(HERE, "moves_g512.npz"), **out)
    print("wrote moves_g512.npz; accept rates", out["dec_yg"].mean(), out["dec_vg"].mean(), (out["dec_alloc"] > 0).mean(), out["dec_spike"].sum())


if __name__ == "__main__":
    main()
