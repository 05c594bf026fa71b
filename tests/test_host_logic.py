"""Host-side logic that needs no GPU: tuning rule, accept windows, R-style number formatting, the closed forms the fused
schedule uses for the hyper-parameter moves (checked against the oracle's per-gene sums), synthetic-data determinism."""
import math
import numpy as np
from oracle import oracle as O
from zigzag_b200 import synth
from zigzag_b200 import zigzag as Z
from _util import make_case, hyper_pair, oracle_state


def test_x_tune_matches_oracle_and_clamps():
    for ws, t in ((23, 0.5), (44, 0.5), (80, 1.3), (0, 0.5), (100, 0.2)):
        a = Z.x_tune(ws, t, 0.44, 1e-4, 10)
        b = O.lib().zzo_x_tune(float(ws), t, 0.44, 1e-4, 10.0)
        assert math.isclose(a, b, rel_tol=1e-14) or (a == b)
    assert Z.x_tune(44, 0.5) == 0.5                      # at the target rate the step size is a fixed point
    assert Z.x_tune(0, 0.5, 0.44, 1e-4, 10) == 1e-4 and Z.x_tune(100, 0.5, 0.44, 1e-4, 10) == 10


def test_window_semantics():
    w = Z._Window()
    assert w.sum() == 23                                 # c(rep(0,77), rep(1,23)), I:277
    for _ in range(23):
        w.push0()
    assert w.sum() == 23                                 # the ones are the NEWEST entries: they leave last
    for _ in range(77):
        w.push0()
    assert w.sum() == 0
    w.push0(); w.accept()
    assert w.sum() == 1 and w.w[-1] == 1


def test_r_number_formatting():
    assert Z._rnum(100) == "100" and Z._rnum(-10.0) == "-10" and Z._rnum(0.1235) == "0.1235" and Z._rnum(round(2.64049, 4)) == "2.6405"
    assert Z._rfmt3(0.5) == "0.500" and Z._rfmt3(1) == "1.000"


def _stats_dict(v, K):
    K1 = K + 1
    return dict(n=v[:K1], sy=v[K1:2 * K1], sy2=v[2 * K1:3 * K1], n_inspike=v[3 * K1], n_out=v[3 * K1 + 1], sz=v[3 * K1 + 2],
                sz2=v[3 * K1 + 3], syo=v[3 * K1 + 4], sy2o=v[3 * K1 + 5], syz=v[3 * K1 + 6], lnl=v[3 * K1 + 7], svg=v[3 * K1 + 8], nan=v[3 * K1 + 9])


class _Shim:
    """Just enough of a zigzag object to call the closed-form reductions (no device)."""
    _level2_sums = Z.zigzag._level2_sums
    _varg_sum = Z.zigzag._varg_sum
    variance_g_upper_bound = math.inf


def test_closed_form_hyper_sums_match_per_gene_sums():
    for K, seed in ((2, 0), (3, 1), (1, 2)):
        G, R = 4000, 4
        X, gl, hd, st = make_case(G, R, seed, K, edge_cases=False)
        ho, _ = hyper_pair(hd)
        s = oracle_state(X, gl, st)
        O.refresh_caches(ho, s)
        O.sweep(ho, s, O.sweep_draws(5, 0, 0, 0, G))
        stats = _stats_dict(O.suffstats(ho, s), K)
        z = _Shim()
        z.num_active_components, z.spike_probability = K, ho.spike_probability
        z.s0, z.s1, z.tau = ho.s0, ho.s1, ho.tau
        am = np.array([ho.active_means[k] for k in range(K)]) + 0.07
        av = np.array([ho.active_variances[k] for k in range(K)]) * 0.93
        got = z._level2_sums(ho.inactive_means - 0.05, ho.inactive_variances * 1.2, am, av, stats)
        ref = O.level2_sums(ho, s, ho.inactive_means - 0.05, ho.inactive_variances * 1.2, am, av)
        assert np.allclose(got, ref, rtol=1e-11)
        lp, lc = z._varg_sum(ho.s0 + 0.03, ho.s1 * 0.9, ho.tau * 1.1, stats)
        assert math.isclose(lp, O.varg_hyper_sum(ho, s, ho.s0 + 0.03, ho.s1 * 0.9, ho.tau * 1.1), rel_tol=1e-11)
        assert math.isclose(lc, O.varg_hyper_sum(ho, s, ho.s0, ho.s1, ho.tau), rel_tol=1e-11)
        assert math.isclose(lc, stats["svg"], rel_tol=1e-11)


def test_synthetic_configs_are_deterministic_and_shaped():
    X1, gl1, t1 = synth.simulate(3000, 4, 7)
    X2, gl2, t2 = synth.simulate(3000, 4, 7)
    assert np.array_equal(X1, X2) and np.array_equal(gl1, gl2)
    assert X1.flags["F_CONTIGUOUS"] and abs(gl1.mean() - 1) < 1e-12
    az = np.isneginf(X1).all(1).mean()
    assert 0.02 < az < 0.09                              # (1 - 0.6) * 0.1 in the spike + undetected low genes
    hd = synth.truth_hyper_dict(4)
    st = synth.initial_state(X1, gl1, hd, 3)
    assert not synth._xg_lik_is_neg_inf(X1, st["Yg"], gl1, hd["alpha_r"])[~np.isneginf(X1).all(1)].any()   # I:480-502
    assert np.all(st["alloc_active_inactive"][st["inactive_spike_allocation"] == 1] == 0)                  # I:422
    assert set(synth.CONFIGS) == {"20k_x2", "200k_x8", "1m_x16", "64ch_60k_x4"}
