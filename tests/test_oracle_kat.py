"""Pins the CPU oracle (the checker of the GPU path) — runs without a GPU.

The reference has no function-level golden vectors (no tests, no seeds, and R is unavailable here), so the oracle is
pinned by: the SURVEY §8c known-answer vector, an mpmath evaluation of the same closed forms, the Random123 Philox4x32-10
known-answer vectors, and agreement with an independent vectorised NumPy restatement (oracle/oracle_np.py)."""
import numpy as np
import mpmath as mp
import pytest
from oracle import oracle as O
from oracle import oracle_np as N
from _util import make_case, hyper_pair, oracle_state, lung_like_hyper

KAT_H = dict(s0=-2.05, s1=-0.355, tau=0.86, alpha_r=[2, 5, 9], spike_probability=0.115, inactive_means=-1.9, inactive_variances=1.78,
             active_means=[2.64, 7.2], active_variances=[1.78, 1.78], weight_active=0.725, weight_within_active=[0.992, 0.008],
             beta=1.0, temperature=1.0, variance_g_upper_bound=float("inf"))


def kat_state(ai=1, k=1):
    X = np.array([[np.log(3.302), -np.inf, np.log(4.96)]])
    st = dict(Yg=np.array([1.1]), variance_g=np.array([0.35]), alloc_active_inactive=np.array([ai], np.int32),
              alloc_within_active=np.array([k], np.int32), inactive_spike_allocation=np.zeros(1, np.int32),
              tuningParam_yg=np.array([0.5]), tuningParam_variance_g=np.array([0.5]))
    return oracle_state(X, np.array([1.2]), st)


def test_survey_known_answer_vector():
    ho, _ = hyper_pair(KAT_H)
    s = kat_state()
    assert O.eval_xg_loglik(ho, s)[0] == -19.185709554063248
    assert O.eval_varg_logprior(ho, s)[0] == 0.04256305502955485
    assert O.eval_level2_loglik(ho, s)[0] == -1.8734249906375684
    assert O.eval_level2_loglik(ho, kat_state(1, 2))[0] == -11.659492406367903
    assert O.eval_level2_loglik(ho, kat_state(0, 1))[0] == -3.857502736971327
    _, pa = O.gibbs_alloc_active_inactive(ho, s, [0.5], [0.5])
    assert abs(pa[0] - 0.95004783959521) < 1e-14


def test_closed_forms_against_mpmath():
    """The fp64-order value of log(1 - p) differs from exact math by 6.4e-11 relative: the oracle must carry the
    fp64-order value; everything else agrees with 40-digit arithmetic to rounding."""
    mp.mp.dps = 40
    h = KAT_H
    gl, y, v = mp.mpf("1.2"), mp.mpf("1.1"), mp.mpf("0.35")
    x = [mp.log(mp.mpf("3.302")), None, mp.log(mp.mpf("4.96"))]
    exact = mp.mpf(0)
    for r, a in enumerate(h["alpha_r"]):
        p = 1 - mp.e ** (-(gl * mp.e ** y) * a)
        if x[r] is None:
            exact += mp.log(1 - p)
        else:
            z = (x[r] - y) / mp.sqrt(v)
            exact += mp.log(p) - (mp.log(mp.sqrt(2 * mp.pi)) + z * z / 2 + mp.log(mp.sqrt(v)))
    ho, _ = hyper_pair(h)
    got = O.eval_xg_loglik(ho, kat_state())[0]
    rel = abs((mp.mpf(got) - exact) / exact)
    assert mp.mpf("1e-11") < rel < mp.mpf("1e-10")          # faithful to fp64 order, not to exact math
    Sg = mp.e ** (mp.mpf(h["s0"]) + mp.mpf(h["s1"]) * y)
    tau = mp.mpf(h["tau"])
    pv = -mp.log(v * mp.sqrt(tau) * mp.sqrt(2 * mp.pi)) - ((mp.log(v) - mp.log(Sg) - tau) / mp.sqrt(tau)) ** 2 / 2
    assert abs(mp.mpf(O.eval_varg_logprior(ho, kat_state())[0]) - pv) < mp.mpf("1e-15")
    act = -mp.log(mp.sqrt(2 * mp.pi) * mp.sqrt(mp.mpf("1.78"))) - (y - mp.mpf("2.64")) ** 2 / (2 * mp.mpf("1.78"))
    assert abs(mp.mpf(O.eval_level2_loglik(ho, kat_state())[0]) - act) < mp.mpf("1e-15")


def test_philox4x32_10_random123_known_answers():
    assert O.philox4x32_10([0] * 4, [0] * 2) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert O.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert O.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_uniform_stream_properties():
    u = np.array([O.uniform4(7, 1, 3, b, g) for b in range(3) for g in range(2000)]).ravel()
    assert u.min() > 0 and u.max() < 1
    assert np.all((u * 2 ** 32 - 0.5) == np.round(u * 2 ** 32 - 0.5))      # (w + 0.5) * 2^-32 exactly
    assert abs(u.mean() - 0.5) < 0.01 and abs(u.var() - 1 / 12) < 0.005
    d = O.sweep_draws(11, 0, 5, 100, 20000)
    for j in (0, 6, 7):
        assert abs(d[j].mean()) < 0.03 and abs(d[j].std() - 1) < 0.03
    assert np.array_equal(O.sweep_draws(11, 0, 5, 100, 300)[:, 50:60], O.sweep_draws(11, 0, 5, 150, 10))   # keyed by global gene index


def _np_state(s):
    return {k: getattr(s, k).copy() for k in ("Yg", "variance_g", "alloc_active_inactive", "alloc_within_active",
                                              "inactive_spike_allocation", "no_detect_spike", "XgLikelihood",
                                              "variance_g_probability", "tuningParam_yg", "tuningParam_variance_g")}


@pytest.mark.parametrize("G,R,K,seed", [(1500, 2, 2, 0), (1200, 5, 3, 1), (900, 16, 1, 2)])
def test_c_oracle_agrees_with_numpy_restatement(G, R, K, seed):
    X, gl, hd, st = make_case(G, R, seed, K)
    ho, _ = hyper_pair(hd)
    s = oracle_state(X, gl, st)
    O.refresh_caches(ho, s)
    px = N.get_px(np.asarray(hd["alpha_r"]), s.Yg, gl)
    xl = N.computeXgLikelihood(X, s.Yg, s.variance_g, px, s.inactive_spike_allocation, s.no_detect_spike)
    assert np.array_equal(np.isneginf(xl), np.isneginf(s.XgLikelihood)) and np.isneginf(xl).sum() >= 1
    fin = np.isfinite(xl)
    assert np.allclose(xl[fin], s.XgLikelihood[fin], rtol=1e-14, atol=0)
    pv = N.computeVarianceGPriorProbability(s.variance_g, N.get_sigmax(hd["s0"], hd["s1"], s.Yg), hd["tau"])
    assert np.allclose(pv, s.variance_g_probability, rtol=1e-14, atol=0)
    l2 = N.level2(hd, s.Yg, s.alloc_active_inactive, s.alloc_within_active, s.inactive_spike_allocation)
    assert np.allclose(l2, O.eval_level2_loglik(ho, s), rtol=1e-15, atol=0)
    rng = np.random.Generator(np.random.PCG64(seed))
    ns = _np_state(s)
    for it in range(3):
        z, u, u1, u2, ua, uw = rng.standard_normal(G), rng.random(G), rng.random(G), rng.random(G), rng.random(G), rng.random(G)
        d1, _ = O.mh_yg(ho, s, z, u); n1 = N.mhYg(hd, X, gl, ns, z, u)
        d2, _ = O.mh_variance_g(ho, s, u1, u2); n2 = N.mhVariance_g(hd, X, gl, ns, u1, u2)
        d3, pa = O.gibbs_alloc_active_inactive(ho, s, ua, uw); n3, npa = N.gibbsAllocationActiveInactive(hd, ns, ua, uw)
        assert np.array_equal(d1, n1) and np.array_equal(d2, n2) and np.array_equal(d3, n3)
        assert np.allclose(pa, npa, rtol=1e-13, atol=1e-300)
        assert np.allclose(s.Yg, ns["Yg"], rtol=1e-15) and np.allclose(s.variance_g, ns["variance_g"], rtol=1e-15)


def test_move_invariants_hold_on_the_oracle():
    """Caches equal a fresh recomputation after any move sequence; in-spike genes are never touched by mhYg / mhVariance_g;
    allocation words stay in range; windows only move under tune=TRUE."""
    G, R = 3000, 4
    X, gl, hd, st = make_case(G, R, 9, edge_cases=False)
    ho, _ = hyper_pair(hd)
    s = oracle_state(X, gl, st)
    O.refresh_caches(ho, s)
    y0, v0, sp0 = s.Yg.copy(), s.variance_g.copy(), s.inactive_spike_allocation == 1
    d = O.sweep_draws(3, 0, 99, 0, G)
    O.mh_yg(ho, s, d[0], d[1]); O.mh_variance_g(ho, s, d[2], d[3])
    assert np.array_equal(s.Yg[sp0], y0[sp0]) and np.array_equal(s.variance_g[sp0], v0[sp0]) and sp0.sum() > 50
    assert (s.Yg[~sp0] != y0[~sp0]).mean() > 0.2
    for step in range(5):
        O.sweep(ho, s, O.sweep_draws(3, 0, step, 0, G), tune=(step % 2 == 0))
    fresh = s.copy(); O.refresh_caches(ho, fresh)
    assert np.allclose(s.XgLikelihood, fresh.XgLikelihood, rtol=1e-13) and np.allclose(s.variance_g_probability, fresh.variance_g_probability, rtol=1e-13)
    assert set(np.unique(s.alloc_within_active)) <= {1, 2} and set(np.unique(s.alloc_active_inactive)) <= {0, 1}
    assert np.all(s.alloc_active_inactive[s.inactive_spike_allocation == 1] == 0)
    ws = s.window_sums("Yg_trace")
    assert ws.min() >= 0 and ws.max() <= 26 and (ws != 23).any()        # 3 tuned sweeps pushed 3 bits
    st_ = O.suffstats(ho, s)
    assert st_[:3].sum() == G and st_[9] == (s.inactive_spike_allocation == 1).sum()


def test_oracle_reproduces_committed_golden_vectors():
    """The committed golden file is what the oracle produces today (guards the checker against silent drift)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "moves_g512.npz"))
    hd = {k[2:]: (g[k] if g[k].ndim else float(g[k])) for k in g.files if k.startswith("h_")}
    st = {k[3:]: g[k] for k in g.files if k.startswith("st_")}
    ho, _ = hyper_pair(hd)
    s = oracle_state(g["X"], g["gl"], st)
    O.refresh_caches(ho, s)
    assert np.array_equal(s.XgLikelihood, g["xglik"]) and np.array_equal(s.variance_g_probability, g["vgprob"])
    dec = O.sweep(ho, s, g["draws"])
    assert np.array_equal(dec, g["dec_sweep"]) and np.array_equal(s.Yg, g["sweep_yg"]) and np.array_equal(s.variance_g, g["sweep_vg"])
    assert float(g["sweep_lnl"]) == O.lnl(ho, s) or (np.isneginf(g["sweep_lnl"]) and np.isneginf(O.lnl(ho, s)))


def test_post_predictive_restatement_sanity():
    """zzo_post_predictive (R/zigzagPostPredictiveSimulation.R:3-156): data simulated FROM the model at the simulation truth gives
    realised-minus-simulated discrepancies that scatter around zero; a wrong detection model (alpha / 20) shows up in the Rumsfeld
    statistic; the draws are counter-based (same step -> same answer)."""
    import numpy as np
    from statistics import NormalDist
    from zigzag_b200 import synth
    from oracle import oracle as O
    G, R = 4000, 4
    X, gl, _ = synth.simulate(G, R, config_id=77)
    hd = synth.truth_hyper_dict(R)
    st = synth.initial_state(X, gl, hd, 5)
    kw = {k: float(hd[k]) for k in ("s0", "s1", "tau", "spike_probability", "inactive_means", "inactive_variances", "weight_active",
                                    "beta", "temperature", "variance_g_upper_bound")}
    ho = O.Hyper.make(hd["alpha_r"], hd["active_means"], hd["active_variances"], hd["weight_within_active"], **kw)
    s = O.State(X, gl, st["Yg"], st["variance_g"], st["alloc_active_inactive"], st["alloc_within_active"], st["inactive_spike_allocation"])
    yo, xf = s.Yg[s.inactive_spike_allocation == 0], X[np.isfinite(X)]
    r = [NormalDist(hd["inactive_means"], hd["inactive_variances"] ** 0.5).inv_cdf(1e-6),
         NormalDist(hd["active_means"][-1], hd["active_variances"][-1] ** 0.5).inv_cdf(1 - 1e-6), yo.min(), yo.max(), xf.min(), xf.max()]
    bins = np.arange(min(r) - 2, max(r) + 2, 0.1)
    a = O.post_predictive(ho, s, 9, 1, bins)
    assert a["W_L"] == O.post_predictive(ho, s, 9, 1, bins)["W_L"] != O.post_predictive(ho, s, 9, 2, bins)["W_L"]
    ws = np.array([[O.post_predictive(ho, s, 9, k, bins)[n] for n in ("W_L", "W_U", "B")] for k in range(8)])
    assert np.abs(ws[:, 0]).max() < 0.5 and np.abs(ws[:, 1]).max() < 0.5 and np.abs(ws[:, 2]).max() < 0.05
    bad = O.Hyper.make(np.asarray(hd["alpha_r"]) / 20, hd["active_means"], hd["active_variances"], hd["weight_within_active"], **kw)
    assert O.post_predictive(bad, s, 9, 1, bins)["B"] > 10 * np.abs(ws[:, 2]).max()


def test_oracle_fused_chain_matches_its_golden_trajectory():
    """tests/golden/chain_g400.npz (make_golden_chain.py): hyper-parameters, scalar tuning, stream positions and state checksums of
    the oracle's fused-schedule chain after 420 burn-in (one tune_all) + 30 sampling generations."""
    import importlib.util, os
    import numpy as np
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_golden_chain", os.path.join(here, "golden", "make_golden_chain.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    got, want = mod.run(), np.load(os.path.join(here, "golden", "chain_g400.npz"))
    for k in want.files:
        if want[k].dtype.kind == "i":
            assert np.array_equal(got[k], want[k]), k
        else:
            assert np.allclose(got[k], want[k], rtol=1e-9, atol=1e-12), k
