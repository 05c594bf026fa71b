"""GPU parity tests proper (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs.  North-star checks: (1) per-gene log-likelihoods within 1e-10 relative in fp64; (2) MH accept/reject and
allocation decisions bit-exact given identical state and identical draws."""
import numpy as np
import pytest
from oracle import oracle as O
from zigzag_b200 import _capi
from _util import setup_pair, rel_close, assert_decisions, compare_state, hyper_pair, oracle_state, gpu_handle, make_case

pytestmark = pytest.mark.gpu

CASES = [(4096, 2, 0, 2), (3000, 4, 1, 2), (2500, 8, 2, 3), (2049, 16, 3, 2), (777, 3, 4, 1), (1500, 5, 5, 4)]


def draws_for(G, n, seed, normal=()):
    rng = np.random.Generator(np.random.PCG64(seed))
    d = rng.random((n, G))
    for j in normal:
        d[j] = rng.standard_normal(G)
    return d


@pytest.mark.parametrize("G,R,seed,K", CASES)
def test_evaluators_match_oracle(G, R, seed, K):
    X, gl, hd, ho, hg, s, h = setup_pair(G, R, seed, K)
    e1 = rel_close(h.eval_xg_loglik(), O.eval_xg_loglik(ho, s), 1e-10, "computeXgLikelihood")
    e2 = rel_close(h.eval_varg_logprior(), O.eval_varg_logprior(ho, s), 1e-10, "computeVarianceGPriorProbability")
    e3 = rel_close(h.eval_level2_loglik(), O.eval_level2_loglik(ho, s), 1e-10, "level-2 likelihood")
    assert np.isneginf(s.XgLikelihood).sum() >= 1      # the -Inf rows are in the parity set
    print(f"max rel err xg {e1:.2e} vg {e2:.2e} l2 {e3:.2e}")
    h.refresh_caches()
    compare_state(h, s)


def test_known_answer_vector():
    """SURVEY.md §8c known-answer vector, evaluated on the device."""
    hd = dict(s0=-2.05, s1=-0.355, tau=0.86, alpha_r=[2, 5, 9], spike_probability=0.115, inactive_means=-1.9,
              inactive_variances=1.78, active_means=[2.64, 7.2], active_variances=[1.78, 1.78], weight_active=0.725,
              weight_within_active=[0.992, 0.008], beta=1.0, temperature=1.0, variance_g_upper_bound=float("inf"))
    ho, hg = hyper_pair(hd)
    G = 33
    X = np.tile(np.array([[np.log(3.302), -np.inf, np.log(4.96)]]), (G, 1))
    st = dict(Yg=np.full(G, 1.1), variance_g=np.full(G, 0.35), alloc_active_inactive=np.ones(G, np.int32),
              alloc_within_active=np.ones(G, np.int32), inactive_spike_allocation=np.zeros(G, np.int32),
              tuningParam_yg=np.full(G, 0.5), tuningParam_variance_g=np.full(G, 0.5))
    st["alloc_within_active"][1] = 2
    st["alloc_active_inactive"][2] = 0
    s = oracle_state(X, np.full(G, 1.2), st)
    h = gpu_handle(X, np.full(G, 1.2), hg, s)
    xl, vp, l2 = h.eval_xg_loglik(), h.eval_varg_logprior(), h.eval_level2_loglik()
    assert abs(xl[0] - (-19.185709554063248)) <= 1e-10 * 19.2
    assert abs(vp[0] - 0.04256305502955485) <= 1e-10 * 0.043
    assert abs(l2[0] - (-1.8734249906375684)) <= 2e-10
    assert abs(l2[1] - (-11.659492406367903)) <= 2e-9
    assert abs(l2[2] - (-3.857502736971327)) <= 4e-10


@pytest.mark.parametrize("G,R,seed,K", CASES)
def test_mh_yg_decisions_bit_exact(G, R, seed, K):
    X, gl, hd, ho, hg, s, h = setup_pair(G, R, seed, K)
    for it in range(3):
        d = draws_for(G, 2, 10 * seed + it, normal=(0,))
        dec_o, lr = O.mh_yg(ho, s, d[0], d[1], tune=True)
        dec_g = h.mh_yg(step=it, tune=True, draws=d, decisions=True)
        assert_decisions(dec_g, dec_o, lr, np.log(d[1]), "mhYg")
        compare_state(h, s)
    assert np.array_equal(h.window_sums()[0], s.window_sums("Yg_trace"))
    assert 0.05 < dec_o.mean() < 0.95


@pytest.mark.parametrize("G,R,seed,K", CASES)
def test_mh_variance_g_decisions_bit_exact(G, R, seed, K):
    X, gl, hd, ho, hg, s, h = setup_pair(G, R, seed, K)
    for it in range(3):
        d = draws_for(G, 2, 20 * seed + it)
        dec_o, lr = O.mh_variance_g(ho, s, d[0], d[1], tune=True)
        dec_g = h.mh_variance_g(step=it, tune=True, draws=d, decisions=True)
        assert_decisions(dec_g, dec_o, lr, np.log(d[1]), "mhVariance_g")
        compare_state(h, s)
    assert np.array_equal(h.window_sums()[1], s.window_sums("variance_g_trace"))


@pytest.mark.parametrize("G,R,seed,K", CASES)
def test_gibbs_alloc_decisions_bit_exact(G, R, seed, K):
    X, gl, hd, ho, hg, s, h = setup_pair(G, R, seed, K, pinned_frac=0.02)
    for it in range(2):
        d = draws_for(G, 2, 30 * seed + it)
        dec_o, pa = O.gibbs_alloc_active_inactive(ho, s, d[0], d[1])
        dec_g = h.gibbs_alloc(step=it, draws=d, decisions=True)
        assert np.array_equal(dec_g, dec_o), np.where(dec_g != dec_o)[0][:10]
        compare_state(h, s)
    d = draws_for(G, 1, 31 * seed)
    dec_o = O.gibbs_alloc_within_active(ho, s, d[0])
    dec_g = h.gibbs_alloc_within(step=9, draws=d, decisions=True)
    assert np.array_equal(dec_g, dec_o)
    compare_state(h, s)
    if K > 1:
        assert len(np.unique(dec_o)) == K + 1


@pytest.mark.parametrize("G,R,seed,K", CASES)
def test_mh_spike_alloc_decisions_bit_exact(G, R, seed, K):
    X, gl, hd, ho, hg, s, h = setup_pair(G, R, seed, K)
    total = 0
    for it in range(3):
        d = draws_for(G, 3, 40 * seed + it, normal=(0, 1))
        before = s.inactive_spike_allocation.copy()
        dec_o, lr = O.mh_spike_alloc(ho, s, d[0], d[1], d[2])
        dec_g = h.mh_spike_alloc(step=it, draws=d, decisions=True)
        assert_decisions(dec_g, dec_o, lr, np.log(d[2]), "mhSpikeAllocation")
        compare_state(h, s)
        total += int((before != s.inactive_spike_allocation).sum())
    assert total > 0   # both directions were exercised


MODES = {"production": _capi.MOVE_FUSED, "reference_order": _capi.MOVE_FUSED | _capi.MOVE_REFERENCE_ORDER}


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("G,R,seed,K", CASES)
def test_fused_sweep_equals_sequential_moves(G, R, seed, K, mode):
    """Both fused kernels (production arithmetic of zz_fast.cuh and the reference-operation-order kernel) reproduce the
    oracle's move-by-move decisions on identical draws; likelihood caches stay within 1e-10, Yg / variance_g within 1e-12."""
    X, gl, hd, ho, hg, s, h = setup_pair(G, R, seed, K, pinned_frac=0.01)
    for it in range(4):
        d = draws_for(G, 9, 50 * seed + it, normal=(0, 6, 7))
        dec_o = O.sweep(ho, s, d, tune=True)
        before = s.copy()
        dec_g = h.sweep(step=it, tune=True, moves=MODES[mode], draws=d, decisions=True)
        bad = {j: np.where(dec_g[j, 0] != dec_o[j])[0][:5] for j in range(4) if not np.array_equal(dec_g[j, 0], dec_o[j])}
        assert not bad, (bad, {j: (before.XgLikelihood[b], before.variance_g[b], before.Yg[b], before.inactive_spike_allocation[b]) for j, b in bad.items()})
        compare_state(h, s)
    assert np.array_equal(h.window_sums()[0], s.window_sums("Yg_trace"))
    assert np.array_equal(h.window_sums()[1], s.window_sums("variance_g_trace"))


VARIANTS = {"upper_bound": dict(variance_g_upper_bound=3.0), "tempered": dict(beta=0.5, temperature=0.7), "prior_only": dict(beta=0.0)}


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("variant", list(VARIANTS))
def test_model_switches_that_change_the_arithmetic(variant, mode):
    """variance_g_upper_bound < Inf (plnorm term + hard reject, P:18-27, Pr:26, Pr:603), the likelihood power beta (P:121-125,
    beta < 1e-7 => likelihood 0: the prior-recovery recipe of scripts/test_script.R:42-53) and the temperature."""
    from zigzag_b200 import synth
    G, R = 3000, 4
    hd = dict(synth.truth_hyper_dict(R)); hd.update(VARIANTS[variant])
    X, gl, hd, st = make_case(G, R, 23, hyper=hd)
    if variant == "upper_bound":
        st["variance_g"] = np.minimum(st["variance_g"], 3.0)                     # I:468
    ho, hg = hyper_pair(hd)
    s = oracle_state(X, gl, st); O.refresh_caches(ho, s)
    h = gpu_handle(X, gl, hg, s)
    rel_close(h.eval_xg_loglik(), O.eval_xg_loglik(ho, s), 1e-10, "XgLikelihood")
    rel_close(h.eval_varg_logprior(), O.eval_varg_logprior(ho, s), 1e-10, "variance_g prior")
    rejected_by_bound = 0
    for it in range(4):
        d = draws_for(G, 9, 300 + it, normal=(0, 6, 7))
        v_before = s.variance_g.copy()
        dec_o = O.sweep(ho, s, d, tune=True)
        dec_g = h.sweep(step=it, tune=True, moves=MODES[mode], draws=d, decisions=True)
        assert np.array_equal(dec_g[:, 0, :], dec_o), {j: np.where(dec_g[j, 0] != dec_o[j])[0][:5] for j in range(4)}
        compare_state(h, s)
        if variant == "upper_bound":
            sf = np.exp(st["tuningParam_variance_g"] * (d[2] - 0.5))
            rejected_by_bound += int(((v_before * sf > 3.0) & (s.inactive_spike_allocation == 0)).sum())
            assert s.variance_g[(v_before <= 3.0)].max() <= 3.0
    if variant == "upper_bound":
        assert rejected_by_bound > 0
    if variant == "prior_only":
        assert np.all(s.XgLikelihood == 0)
    st_o, st_g = O.suffstats(ho, s), h.suffstats()[0]
    fin = np.isfinite(st_o)
    assert np.allclose(st_g[fin], st_o[fin], rtol=1e-11, atol=1e-9)


@pytest.mark.parametrize("moves", [_capi.MOVE_YG, _capi.MOVE_VARIANCE_G, _capi.MOVE_ALLOC, _capi.MOVE_SPIKE,
                                   _capi.MOVE_VARIANCE_G | _capi.MOVE_ALLOC])
def test_production_kernel_move_subsets(moves):
    """zz_sweep with a subset mask runs the production arithmetic for just those moves (and keeps caches consistent)."""
    G, R = 3000, 8
    X, gl, hd, ho, hg, s, h = setup_pair(G, R, 31)
    for it in range(2):
        d = draws_for(G, 9, 60 + it, normal=(0, 6, 7))
        if moves & _capi.MOVE_YG: O.mh_yg(ho, s, d[0], d[1])
        if moves & _capi.MOVE_VARIANCE_G: O.mh_variance_g(ho, s, d[2], d[3])
        if moves & _capi.MOVE_ALLOC: O.gibbs_alloc_active_inactive(ho, s, d[4], d[5])
        if moves & _capi.MOVE_SPIKE: O.mh_spike_alloc(ho, s, d[6], d[7], d[8])
        h.sweep(step=it, moves=moves, draws=d)
        compare_state(h, s)


def test_production_and_reference_kernels_interleave():
    """The cached detection term A(y) of the production kernel is rebuilt whenever a reference-order move, a new alpha_r or a
    host state upload changed what it depends on."""
    G, R = 2500, 4
    X, gl, hd, ho, hg, s, h = setup_pair(G, R, 41, edge_cases=False)   # no -Inf rows: px_commit turns them into NaN (as R would)
    step = 0
    for rnd in range(3):
        d = draws_for(G, 9, 70 + rnd, normal=(0, 6, 7))
        O.sweep(ho, s, d); h.sweep(step, draws=d); step += 1
        O.mh_yg(ho, s, d[0], d[1]); h.mh_yg(step, draws=d[0:2]); step += 1
        O.mh_spike_alloc(ho, s, d[6], d[7], d[8]); h.mh_spike_alloc(step, draws=d[6:9]); step += 1
        new_alpha = ho.alpha_r[rnd % R] * 1.03
        O.px_commit(ho, s, rnd % R, new_alpha); h.px_commit(rnd % R, new_alpha)
        compare_state(h, s)
    d = draws_for(G, 9, 99, normal=(0, 6, 7))
    dec_o = O.sweep(ho, s, d); dec_g = h.sweep(step, draws=d, decisions=True)
    assert np.array_equal(dec_g[:, 0], dec_o)
    compare_state(h, s)


def test_philox_stream_matches_oracle_bit_for_bit():
    X, gl, hd, ho, hg, s, h0 = setup_pair(1000, 2, 0)
    h0.close()
    h = gpu_handle(X, gl, hg, s, chains=3, seed=0xDEADBEEFCAFE, gene_offset=123456)
    step = (7 << 33) + 5
    dv, da, dw = (h.philox_draws(step, sl) for sl in (_capi.SLOT_VARIANCE_G, _capi.SLOT_ALLOC, _capi.SLOT_ALLOC_WITHIN))
    for c in range(3):
        for g in (0, 1, 17, 999):
            b0 = O.uniform4(0xDEADBEEFCAFE, c, step, 0, 123456 + g)
            b1 = O.uniform4(0xDEADBEEFCAFE, c, step, 1, 123456 + g)
            assert (dv[0, c, g], dv[1, c, g]) == (b0[3], b1[0])
            assert (da[0, c, g], da[1, c, g]) == (b1[1], b1[2])
            assert dw[0, c, g] == b1[2]
    for d in (dv, da, dw):
        assert d.min() > 0 and d.max() < 1
    # normals: same uniforms, Box-Muller through CUDA's log/sincospi vs libm's log/cos/sin
    d = h.philox_draws(step=3, slot=_capi.SLOT_SPIKE)
    ref = np.array([[O.sweep_draws(0xDEADBEEFCAFE, c, 3, 123456, 1000)[6:9]] for c in range(3)])[:, 0]  # [C][3][G]
    assert np.allclose(d[0], ref[:, 0], rtol=0, atol=1e-13)
    assert np.allclose(d[1], ref[:, 1], rtol=0, atol=1e-13)
    assert np.array_equal(d[2], ref[:, 2])


@pytest.mark.parametrize("G,R,seed,K", CASES[:3])
def test_philox_sweep_matches_oracle_on_exported_draws(G, R, seed, K):
    """draws == NULL path: the device generates its own draws; exported, they drive the oracle to the same decisions."""
    X, gl, hd, ho, hg, s, h = setup_pair(G, R, seed, K)
    for step in range(3):
        d = np.empty((9, G))
        d[0:2] = h.philox_draws(step, _capi.SLOT_YG)[:, 0]
        d[2:4] = h.philox_draws(step, _capi.SLOT_VARIANCE_G)[:, 0]
        d[4:6] = h.philox_draws(step, _capi.SLOT_ALLOC)[:, 0]
        d[6:9] = h.philox_draws(step, _capi.SLOT_SPIKE)[:, 0]
        dec_g = h.sweep(step=step, tune=False, draws=None, decisions=True)
        dec_o = O.sweep(ho, s, d, tune=False)
        assert np.array_equal(dec_g[:, 0, :], dec_o)
        compare_state(h, s)


@pytest.mark.parametrize("G,R,seed,K", CASES)
def test_reductions_match_oracle(G, R, seed, K):
    X, gl, hd, ho, hg, s, h = setup_pair(G, R, seed, K)
    d = draws_for(G, 9, seed, normal=(0, 6, 7))
    O.sweep(ho, s, d); h.sweep(0, draws=d)
    st_o, st_g = O.suffstats(ho, s), h.suffstats()[0]
    K1 = K + 1
    assert np.array_equal(st_g[:K1], st_o[:K1]) and st_g[3 * K1] == st_o[3 * K1] and st_g[3 * K1 + 1] == st_o[3 * K1 + 1]
    assert st_g[:K1].sum() == G
    fin = np.isfinite(st_o)
    assert np.allclose(st_g[fin], st_o[fin], rtol=1e-11, atol=1e-9)
    assert np.array_equal(st_g[~fin], st_o[~fin])
    a = h.varg_hyper_sum(ho.s0 + 0.01, ho.s1 * 1.01, ho.tau * 0.97)
    assert np.isclose(a[0], O.varg_hyper_sum(ho, s, ho.s0 + 0.01, ho.s1 * 1.01, ho.tau * 0.97), rtol=1e-12)
    assert np.isclose(a[1], s.variance_g_probability[s.inactive_spike_allocation == 0].sum(), rtol=1e-12)
    am = np.array([ho.active_means[k] for k in range(K)]) + 0.05
    av = np.array([ho.active_variances[k] for k in range(K)]) * 1.1
    l_g = h.level2_sums(ho.inactive_means - 0.1, ho.inactive_variances * 1.1, am, av)
    l_o = O.level2_sums(ho, s, ho.inactive_means - 0.1, ho.inactive_variances * 1.1, am, av)
    assert np.allclose(l_g, l_o, rtol=1e-12)
    ap = np.array([ho.alpha_r[r] for r in range(R)]) * np.linspace(0.9, 1.1, R)
    mask = np.ones(R, np.uint8); mask[R // 2] = 0
    p_g, p_o = h.px_delta(ap, mask), O.px_delta(ho, s, ap, mask)
    assert p_g[R // 2] == 0
    for r in range(R):   # sums of differences of -Inf rows are NaN on both sides (R != Inf guard, Pr:232)
        assert (np.isnan(p_g[r]) and np.isnan(p_o[r])) or np.isclose(p_g[r], p_o[r], rtol=1e-9, atol=1e-9), (r, p_g[r], p_o[r])
    lg, lo = h.lnl(), O.lnl(ho, s)
    assert (lg == lo) or np.isclose(lg, lo, rtol=1e-12)


def test_px_commit_and_varg_commit():
    for R in (2, 5):
        X, gl, hd, ho, hg, s, h = setup_pair(2000, R, 11, edge_cases=False)
        new_alpha = ho.alpha_r[1] * 1.07
        O.px_commit(ho, s, 1, new_alpha); h.px_commit(1, new_alpha)
        assert h.get_hyper().alpha_r[1] == new_alpha
        compare_state(h, s)
        ho.tau *= 0.9; ho.s0 += 0.02
        hg2 = h.get_hyper(); hg2.tau = ho.tau; hg2.s0 = ho.s0; h.set_hyper(hg2)
        out = s.inactive_spike_allocation == 0
        s.variance_g_probability[out] = O.eval_varg_logprior(ho, s)[out]
        h.commit_varg_hyper()
        compare_state(h, s)


def test_tune_and_windows():
    X, gl, hd, ho, hg, s, h = setup_pair(3000, 4, 21)
    for it in range(12):
        d = draws_for(3000, 9, 700 + it, normal=(0, 6, 7))
        O.sweep(ho, s, d, tune=True); h.sweep(it, tune=True, draws=d)
    wy, wv = h.window_sums()
    assert np.array_equal(wy, s.window_sums("Yg_trace")) and np.array_equal(wv, s.window_sums("variance_g_trace"))
    O.tune_per_gene(s, 0.44); h.tune(0.44)
    g = h.get_state()
    rel_close(g["tuningParam_yg"], s.tuningParam_yg, 1e-12, "tuningParam_yg")
    rel_close(g["tuningParam_variance_g"], s.tuningParam_variance_g, 1e-12, "tuningParam_variance_g")
    assert g["tuningParam_yg"].min() >= 1e-4 and g["tuningParam_yg"].max() <= 10
    h.reset_windows()
    assert np.all(h.window_sums()[0] == 23)


def test_alloc_trace_accumulates_one_hot():
    X, gl, hd, ho, hg, s, h = setup_pair(1500, 3, 5, K=3)
    ref = np.zeros((4, 1500), np.uint32)
    h.reset_alloc_trace()
    for it in range(5):
        d = draws_for(1500, 2, 900 + it)
        O.gibbs_alloc_active_inactive(ho, s, d[0], d[1]); h.gibbs_alloc(it, draws=d)
        comp = np.where(s.alloc_active_inactive == 1, s.alloc_within_active, 0)
        ref[comp, np.arange(1500)] += 1
        h.accumulate_alloc_trace()
    tr = h.get_alloc_trace()
    assert np.array_equal(tr, ref) and np.all(tr.sum(0) == 5)


def test_sharding_does_not_change_decisions():
    """Philox is keyed by the global gene index: a shard [g0, g1) reproduces the full run's decisions."""
    G, R = 5000, 4
    X, gl, hd, ho, hg, s, h = setup_pair(G, R, 8)
    full = [h.sweep(step=t, decisions=True)[:, 0].copy() for t in range(2)]
    full_state = h.get_state()
    X2, gl2, hd2, st2 = make_case(G, R, 8)
    s2 = oracle_state(X2, gl2, st2); O.refresh_caches(ho, s2)
    g0 = 0
    for g1 in (1777, 3500, G):
        sl = slice(g0, g1)
        sub = O.State(X2[sl], gl2[sl], s2.Yg[sl], s2.variance_g[sl], s2.alloc_active_inactive[sl], s2.alloc_within_active[sl],
                      s2.inactive_spike_allocation[sl], s2.XgLikelihood[sl], s2.variance_g_probability[sl],
                      s2.tuningParam_yg[sl], s2.tuningParam_variance_g[sl])
        hs = gpu_handle(np.asfortranarray(X2[sl]), gl2[sl], hg, sub, gene_offset=g0)
        for t in range(2):
            dec = hs.sweep(step=t, decisions=True)[:, 0]
            assert np.array_equal(dec, full[t][:, sl])
        gs = hs.get_state()
        assert np.array_equal(gs["Yg"], full_state["Yg"][sl]) and np.array_equal(gs["variance_g"], full_state["variance_g"][sl])
        hs.close()
        g0 = g1


def test_batched_chains_are_independent_and_match_oracle():
    G, R, C = 1200, 4, 5
    X, gl, hd, ho, hg, s, h0 = setup_pair(G, R, 13)
    h0.close()
    h = gpu_handle(X, gl, hg, s, chains=C, seed=77)
    hy = [ho.copy() for _ in range(C)]
    ss = [s.copy() for _ in range(C)]
    for c in range(C):     # different hyper-parameters per chain
        hy[c].tau = ho.tau * (1 + 0.1 * c); hy[c].inactive_means = ho.inactive_means - 0.2 * c
        g = h.get_hyper(c); g.tau = hy[c].tau; g.inactive_means = hy[c].inactive_means; h.set_hyper(g, c)
        O.refresh_caches(hy[c], ss[c])
    h.refresh_caches()
    for step in range(2):
        d = np.concatenate([h.philox_draws(step, _capi.SLOT_YG), h.philox_draws(step, _capi.SLOT_VARIANCE_G),
                            h.philox_draws(step, _capi.SLOT_ALLOC), h.philox_draws(step, _capi.SLOT_SPIKE)])  # [9][C][G]
        dec = h.sweep(step, decisions=True)
        for c in range(C):
            dec_o = O.sweep(hy[c], ss[c], d[:, c, :])
            assert np.array_equal(dec[:, c, :], dec_o)
            compare_state(h, ss[c], chain=c)
    st = h.suffstats()
    for c in range(C):
        assert np.allclose(st[c], O.suffstats(hy[c], ss[c]), rtol=1e-11, atol=1e-9)
    assert not np.array_equal(dec[0, 0], dec[0, 1])


@pytest.mark.parametrize("mode", ["pageable", "pageable_3chunks", "pinned_zero_copy", "pinned_staged"])
def test_sweep_host_round_trip_equals_resident_sweep(mode, monkeypatch):
    """zz_sweep_host (state in host arrays, up and down every call) leaves exactly what the resident sweep leaves, on each of its
    transfer paths: staged copies in one or several pipelined chunks, and the single kernel that reads / writes pinned host memory."""
    import torch
    G, R = 4000, 4
    X, gl, hd, ho, hg, s, h = setup_pair(G, R, 17)
    h2 = gpu_handle(X, gl, hg, s)
    if mode == "pageable_3chunks":
        monkeypatch.setenv("ZZ_HOST_CHUNKS", "3")
    if mode == "pinned_staged":
        monkeypatch.setenv("ZZ_HOST_NO_ZEROCOPY", "1")
        monkeypatch.setenv("ZZ_HOST_CHUNKS", "2")
    keys = ("Yg", "variance_g", "alloc_active_inactive", "alloc_within_active", "inactive_spike_allocation", "XgLikelihood",
            "variance_g_probability", "tuningParam_yg", "tuningParam_variance_g")
    host = {k: getattr(s, k).copy() for k in keys}
    if mode.startswith("pinned"):
        keep = {k: torch.from_numpy(host[k]).pin_memory() for k in keys}
        host = {k: keep[k].numpy() for k in keys}                    # views of the pinned buffers
    for step in range(2):
        h.sweep(step)
        nb = h2.sweep_host(step, False, *[host[k] for k in host])
        assert nb == (G * 44, G * 44)
    g = h.get_state()
    g2 = h2.get_state()                                              # the handle's resident state follows the host arrays
    for k in host:
        if k == "alloc_within_active":
            act = g["alloc_active_inactive"] == 1
            assert np.array_equal(host[k][act], g[k][act])
        else:
            assert np.array_equal(host[k], g[k]), k
            assert np.array_equal(g2[k], g[k]), k
    h2.sweep(2); h.sweep(2)                                          # ... and a resident sweep can follow directly
    assert np.array_equal(h2.get_state()["Yg"], h.get_state()["Yg"])


def test_nan_ratio_rejects_and_raises_flag():
    X, gl, hd, ho, hg, s, h = setup_pair(512, 2, 3, edge_cases=False)
    assert h.nan_flag() == 0
    y = s.Yg.copy(); y[5] = np.nan
    h.set_state(0, Yg=y)
    dec = h.mh_yg(0, draws=draws_for(512, 2, 1, normal=(0,)), decisions=True)
    assert dec[5] == 0 and h.nan_flag() == 1
    assert h.suffstats()[0][-1] == 1


def test_error_codes_not_exceptions():
    with pytest.raises(_capi.ZigzagGpuError):
        _capi.Handle(0, 2, 2)
    with pytest.raises(_capi.ZigzagGpuError):
        _capi.Handle(10, 1, 2)      # "Data must have > 1 library", I:135
    h = _capi.Handle(10, 2, 2)
    with pytest.raises(_capi.ZigzagGpuError, match="zz_upload_data"):
        h.sweep(0)
    with pytest.raises(_capi.ZigzagGpuError, match="chain"):
        h.eval_xg_loglik(chain=3)


def test_full_size_properties_1m_x16():
    """BASELINE.json config 4 at full size, checked through size-independent properties: after sweeps the two caches equal
    a fresh recomputation (idempotence of refresh), component counts sum to G, windows move only when tuned, and a
    strided sample of genes replays bit-exactly on the oracle."""
    from zigzag_b200 import synth
    X, gl, hd, st, _ = synth.make_config("1m_x16")
    G, R = X.shape
    ho, hg = hyper_pair(hd)
    h = _capi.Handle(G, R, 2, seed=5)
    h.upload_data(X, gl); h.set_hyper(hg)
    h.set_state(0, **st)
    h.refresh_caches()
    before = h.get_state()
    decs = [h.sweep(step=t, decisions=True)[:, 0] for t in range(3)]
    after = h.get_state()
    assert h.nan_flag() == 0
    assert 0.2 < decs[-1][0].mean() < 0.9 and 0.2 < decs[-1][1].mean() < 0.95
    h.refresh_caches()
    fresh = h.get_state()
    rel_close(after["XgLikelihood"], fresh["XgLikelihood"], 1e-10, "XgLikelihood cache vs recompute")
    rel_close(after["variance_g_probability"], fresh["variance_g_probability"], 1e-10, "variance_g_probability cache vs recompute")
    stt = h.suffstats()[0]
    assert stt[:3].sum() == G and stt[-1] == 0
    assert np.all(h.window_sums()[0] == 23)
    # replay a strided sample on the oracle with the exported draws
    idx = np.arange(0, G, 997)
    sub = O.State(X[idx], gl[idx], before["Yg"][idx], before["variance_g"][idx], before["alloc_active_inactive"][idx],
                  before["alloc_within_active"][idx], before["inactive_spike_allocation"][idx], before["XgLikelihood"][idx],
                  before["variance_g_probability"][idx], before["tuningParam_yg"][idx], before["tuningParam_variance_g"][idx])
    for t in range(3):
        d = np.concatenate([h.philox_draws(t, sl)[:, 0][:, idx] for sl in (_capi.SLOT_YG, _capi.SLOT_VARIANCE_G, _capi.SLOT_ALLOC, _capi.SLOT_SPIKE)])
        dec_o = O.sweep(ho, sub, d)
        assert np.array_equal(decs[t][:, idx], dec_o)
    assert np.allclose(after["Yg"][idx], sub.Yg, rtol=1e-12)


def test_committed_golden_vectors():
    """tests/golden/moves_g512.npz (generated by make_golden_moves.py from the oracle, committed): the GPU reproduces the stored
    likelihoods to 1e-10 and the stored decisions exactly, through both the per-move entry points and both sweep kernels."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "moves_g512.npz"))
    hd = {k[2:]: (g[k] if g[k].ndim else float(g[k])) for k in g.files if k.startswith("h_")}
    st = {k[3:]: g[k] for k in g.files if k.startswith("st_")}
    ho, hg = hyper_pair(hd)
    X, gl, d = g["X"], g["gl"], g["draws"]

    def fresh():
        s = oracle_state(X, gl, st)
        s.XgLikelihood[:] = g["xglik"]; s.variance_g_probability[:] = g["vgprob"]
        return gpu_handle(X, gl, hg, s)

    h = fresh()
    rel_close(h.eval_xg_loglik(), g["xglik"], 1e-10, "golden XgLikelihood")
    rel_close(h.eval_varg_logprior(), g["vgprob"], 1e-10, "golden variance_g prior")
    rel_close(h.eval_level2_loglik(), g["level2"], 1e-10, "golden level-2")
    assert np.array_equal(h.mh_yg(0, draws=d[0:2], decisions=True), g["dec_yg"])
    rel_close(h.get_state()["Yg"], g["yg_after"], 1e-12, "golden Yg")
    h = fresh()
    assert np.array_equal(h.mh_variance_g(0, draws=d[2:4], decisions=True), g["dec_vg"])
    rel_close(h.get_state()["variance_g"], g["vg_after"], 1e-12, "golden variance_g")
    h = fresh()
    assert np.array_equal(h.gibbs_alloc(0, draws=d[4:6], decisions=True), g["dec_alloc"])
    h = fresh()
    assert np.array_equal(h.mh_spike_alloc(0, draws=d[6:9], decisions=True), g["dec_spike"])
    assert np.array_equal(h.get_state()["inactive_spike_allocation"], g["spike_after"])
    for mode in MODES.values():
        h = fresh()
        dec = h.sweep(0, moves=mode, draws=d, decisions=True)
        assert np.array_equal(dec[:, 0], g["dec_sweep"])
        stt = h.get_state()
        rel_close(stt["Yg"], g["sweep_yg"], 1e-12, "golden sweep Yg"); rel_close(stt["variance_g"], g["sweep_vg"], 1e-12, "golden sweep variance_g")
        rel_close(stt["XgLikelihood"], g["sweep_xglik"], 1e-10, "golden sweep XgLikelihood")
        fin = np.isfinite(g["sweep_stats"])
        assert np.allclose(h.suffstats()[0][fin], g["sweep_stats"][fin], rtol=1e-11, atol=1e-9)


def test_fused_statistics_epilogue_matches_separate_reduction():
    """zz_set_stats_output_dev: the sweep's last CTA leaves the same statistics (bit for bit) as zz_suffstats computes from the
    per-CTA rows, for several chains, and they match the oracle."""
    import torch
    G, R, C = 6000, 4, 3
    X, gl, hd, ho, hg, s, h0 = setup_pair(G, R, 19)
    h0.close()
    h = gpu_handle(X, gl, hg, s, chains=C, seed=5)
    out = torch.zeros(C * h.nstat, dtype=torch.float64, device="cuda")
    h.set_stream(torch.cuda.current_stream().cuda_stream)
    h.set_stats_output_dev(out.data_ptr())
    for step in range(3):
        h.sweep(step)
        torch.cuda.synchronize()
        fused = out.cpu().numpy().reshape(C, -1).copy()
        sep = h.suffstats()
        assert np.array_equal(fused, sep)
        assert np.all(fused[:, :3].sum(1) == G)
    h.set_stats_output_dev(0)
    out.zero_()
    h.sweep(9); torch.cuda.synchronize()
    assert float(out.abs().sum()) == 0.0          # switched off again


@pytest.mark.parametrize("tune", [False, True])
@pytest.mark.parametrize("G,R,seed,K", CASES + [(5000, 16, 6, 2), (4000, 8, 7, 1)])
def test_straight_line_production_tile_matches_oracle(G, R, seed, K, tune):
    """The kernel bench.py times: zz_sweep with Philox draws and no decision export takes the straight-line tile (K <= 3) or
    the generic production tile (K = 4).  Its draws are exported and replayed on the oracle; equality of every allocation word
    and spike flag plus Yg / variance_g to 1e-12 means every accept/reject decision was identical."""
    X, gl, hd, ho, hg, s, h = setup_pair(G, R, seed, K, pinned_frac=0.01)
    for step in range(4):
        d = np.concatenate([h.philox_draws(step, sl)[:, 0] for sl in
                            (_capi.SLOT_YG, _capi.SLOT_VARIANCE_G, _capi.SLOT_ALLOC, _capi.SLOT_SPIKE)])
        y_before = s.Yg.copy()
        O.sweep(ho, s, d, tune=tune)
        h.sweep(step, tune=tune)                     # draws=None, decisions=False  -> production instantiation
        compare_state(h, s)
        assert (s.Yg != y_before).mean() > 0.2
    if tune:
        assert np.array_equal(h.window_sums()[0], s.window_sums("Yg_trace"))
        assert np.array_equal(h.window_sums()[1], s.window_sums("variance_g_trace"))
    st_o, st_g = O.suffstats(ho, s), h.suffstats()[0]
    fin = np.isfinite(st_o)
    assert np.allclose(st_g[fin], st_o[fin], rtol=1e-11, atol=1e-9) and np.array_equal(st_g[:K + 1], st_o[:K + 1])
    assert h.nan_flag() in (0, 1)


def test_straight_line_tile_handles_beta_and_temperature():
    from zigzag_b200 import synth
    for variant in ("tempered", "prior_only"):
        G, R = 3000, 4
        hd = dict(synth.truth_hyper_dict(R)); hd.update(VARIANTS[variant])
        X, gl, hd, st = make_case(G, R, 29, hyper=hd)
        ho, hg = hyper_pair(hd)
        s = oracle_state(X, gl, st); O.refresh_caches(ho, s)
        h = gpu_handle(X, gl, hg, s)
        for step in range(3):
            d = np.concatenate([h.philox_draws(step, sl)[:, 0] for sl in
                                (_capi.SLOT_YG, _capi.SLOT_VARIANCE_G, _capi.SLOT_ALLOC, _capi.SLOT_SPIKE)])
            O.sweep(ho, s, d); h.sweep(step)
            compare_state(h, s)


@pytest.mark.parametrize("R,alpha_spread", [(4, 40.0), (16, 100.0), (3, 2.0), (6, 30.0)])
def test_production_sweep_outside_the_straight_line_preconditions(R, alpha_spread):
    """The straight-line tile needs max(alpha) <= 16 min(alpha) (one clamp of t serves every library) and R in {2,4,8,16}; anything
    else must fall back to the generic production tile with the per-library clamp — same draws, same decisions, same caches."""
    from zigzag_b200 import synth
    G = 3000
    hd = dict(synth.truth_hyper_dict(R))
    hd["alpha_r"] = np.geomspace(0.5, 0.5 * alpha_spread, R)
    X, gl, hd, st = make_case(G, R, 41, hyper=hd)
    ho, hg = hyper_pair(hd)
    s = oracle_state(X, gl, st); O.refresh_caches(ho, s)
    h = gpu_handle(X, gl, hg, s)
    for step in range(3):
        d = np.concatenate([h.philox_draws(step, sl)[:, 0] for sl in
                            (_capi.SLOT_YG, _capi.SLOT_VARIANCE_G, _capi.SLOT_ALLOC, _capi.SLOT_SPIKE)])
        O.sweep(ho, s, d); h.sweep(step)
        compare_state(h, s)
    st_o, st_g = O.suffstats(ho, s), h.suffstats()[0]
    fin = np.isfinite(st_o)
    assert np.allclose(st_g[fin], st_o[fin], rtol=1e-11, atol=1e-9)


def pp_bins(hd, Yg, spike, X):
    """bins of posteriorPredictiveSimulation (PostPred:9-17)."""
    from statistics import NormalDist
    K = len(hd["active_means"])
    lo_m = NormalDist(hd["inactive_means"], np.sqrt(hd["inactive_variances"])).inv_cdf(0.000001)
    hi_m = NormalDist(hd["active_means"][K - 1], np.sqrt(hd["active_variances"][K - 1])).inv_cdf(0.999999)
    yo = Yg[spike == 0]
    xf = X[np.isfinite(X)]
    r = [lo_m, hi_m, yo.min(), yo.max(), xf.min(), xf.max()]
    return np.arange(min(r) - 2, max(r) + 2, 0.1)


@pytest.mark.parametrize("G,R,seed", [(6000, 4, 51), (3000, 16, 52), (2500, 2, 53)])
def test_post_predictive_discrepancies_match_oracle(G, R, seed):
    """zz_post_predictive: histograms on the device (integer atomics, fixed-point model weights), functionals on the host; the
    oracle simulates the same data set from the same Philox draws and forms the statistics gene by gene."""
    X, gl, hd, ho, hg, s, h = setup_pair(G, R, seed)
    for step in range(2):
        d = np.concatenate([h.philox_draws(step, sl)[:, 0] for sl in (_capi.SLOT_YG, _capi.SLOT_VARIANCE_G, _capi.SLOT_ALLOC, _capi.SLOT_SPIKE)])
        O.sweep(ho, s, d); h.sweep(step)
    bins = pp_bins(hd, s.Yg, s.inactive_spike_allocation, X)
    for step in (100, 101):
        g = h.post_predictive(step, bins)
        o = O.post_predictive(ho, s, 1234, step, bins)      # 1234: the Philox key of gpu_handle()
        for k in ("W_L", "W_U", "B"):
            assert np.isclose(g[k], o[k], rtol=1e-9, atol=1e-11), (k, g[k], o[k])
        assert np.allclose(g["W_L_r"], o["W_L_r"], rtol=1e-9, atol=1e-11) and np.allclose(g["B_r"], o["B_r"], rtol=1e-9, atol=1e-11)
    assert g["W_L"] != 0 and np.all(np.isfinite(g["W_L_r"]))
    g2 = h.post_predictive(100, bins)                       # counter-based draws: the same step reproduces the same data set
    assert g2["W_L"] != g["W_L"] and g2["W_L"] == h.post_predictive(100, bins)["W_L"]
