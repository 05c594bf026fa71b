"""Device-resident generation loop: zz_run_generations against the oracle's fused-schedule chain
(oracle/zz_oracle.c: zzo_chain_run_fused).  Both start from the same per-gene state, hyper-parameters, scalar tuning state and
positions in the two Philox streams; afterwards every hyper-parameter, every scalar tuning parameter and window, the number of
uniforms consumed by the hyper moves, the per-gene state and the allocation trace must agree."""
import ctypes as C
import math
import numpy as np
import pytest
from oracle import oracle as O
from zigzag_b200 import _capi, synth

pytestmark = pytest.mark.gpu


def _start_pair(G, R, K, seed, thr_a):
    X, gl, _ = synth.simulate(G, R, config_id=300 + seed)
    pri = O.Priors.defaults(-1.0, thr_a)
    ch = O.Chain(X, gl, pri, K, seed)
    st, hy = ch.state_arrays(), ch.hyper
    hg = _capi.Hyper()
    C.memmove(C.byref(hg), C.byref(hy), C.sizeof(_capi.Hyper))
    h = _capi.Handle(G, R, K, seed=seed)
    h.upload_data(X, gl)
    h.set_hyper(hg, 0)
    h.set_state(0, Yg=st["Yg"], variance_g=st["variance_g"], alloc_active_inactive=st["alloc_active_inactive"],
                alloc_within_active=st["alloc_within_active"], inactive_spike_allocation=st["inactive_spike_allocation"],
                XgLikelihood=st["XgLikelihood"], variance_g_probability=st["variance_g_probability"],
                tuningParam_yg=st["tuningParam_yg"], tuningParam_variance_g=st["tuningParam_variance_g"])
    h.reset_windows()
    h.reset_alloc_trace()
    h.chain_begin(pri, ch.scalars())
    return ch, h


def _compare(ch, h, K, R, rtol=1e-9):
    sc_g, hy_g = h.chain_read()
    sc_o, hy_o = ch.scalars(), ch.hyper
    for f in ("s0", "s1", "tau", "spike_probability", "inactive_means", "inactive_variances", "weight_active"):
        assert np.isclose(getattr(hy_g, f), getattr(hy_o, f), rtol=rtol, atol=1e-13), f
    for f, n in (("alpha_r", R), ("active_means", K), ("active_variances", K), ("weight_within_active", K)):
        assert np.allclose(list(getattr(hy_g, f))[:n], list(getattr(hy_o, f))[:n], rtol=rtol, atol=1e-13), f
    # identical consumption of the hyper-level stream = every rejection loop and every accept/reject took the same path
    assert sc_g.hyper_ctr == sc_o.hyper_ctr and sc_g.gen == sc_o.gen and sc_g.step == sc_o.step and sc_g.n_samples == sc_o.n_samples
    for f in ("t_im", "t_av", "t_s0", "t_s1", "t_tau", "t_s0tau"):
        assert np.isclose(getattr(sc_g, f), getattr(sc_o, f), rtol=rtol), f
    assert np.allclose(list(sc_g.t_am)[:K], list(sc_o.t_am)[:K], rtol=rtol)
    assert np.allclose(list(sc_g.t_alpha)[:R], list(sc_o.t_alpha)[:R], rtol=rtol)
    assert np.allclose(list(sc_g.active_means_dif)[:K], list(sc_o.active_means_dif)[:K], rtol=rtol)
    for f in ("w_im", "w_av", "w_s0", "w_s1", "w_tau", "w_s0tau"):
        assert bytes(getattr(sc_g, f)) == bytes(getattr(sc_o, f)), f
    assert all(bytes(sc_g.w_am[k]) == bytes(sc_o.w_am[k]) for k in range(K))
    assert all(bytes(sc_g.w_alpha[r]) == bytes(sc_o.w_alpha[r]) for r in range(R))
    g, o = h.get_state(0), ch.state_arrays()
    for k in ("alloc_active_inactive", "inactive_spike_allocation"):
        assert np.array_equal(g[k], o[k]), k
    act = o["alloc_active_inactive"] == 1
    assert np.array_equal(g["alloc_within_active"][act], o["alloc_within_active"][act])
    for k in ("Yg", "variance_g", "tuningParam_yg", "tuningParam_variance_g"):
        assert np.allclose(g[k], o[k], rtol=1e-9, atol=1e-12), k
    # the two caches are rebuilt from the state when somebody asks for them after a device-resident segment; the reference keeps
    # variance_g_probability current for out-of-spike genes only (Pr:113,147,187 overwrite [out_spike_idx])
    out = o["inactive_spike_allocation"] == 0
    for k, m in (("XgLikelihood", np.ones_like(out)), ("variance_g_probability", out)):
        fin = np.isfinite(o[k]) & m
        assert np.array_equal(np.isfinite(g[k])[m], np.isfinite(o[k])[m]) and np.allclose(g[k][fin], o[k][fin], rtol=1e-8, atol=1e-9), k
    assert h.nan_flag() == 0


@pytest.mark.parametrize("G,R,K,thr_a", [(3000, 4, 2, (1.0, 6.0)), (2000, 2, 1, (1.0,)), (1500, 8, 3, (0.5, 3.0, 6.0)), (1111, 16, 2, (1.0, 6.0)),
                                         (900, 5, 2, (1.0, 6.0))])
def test_device_resident_generations_match_the_oracle_chain(G, R, K, thr_a):
    """Whole segments of generations in one cooperative launch (k_generations): per-gene sweep, fixed-point sufficient statistics,
    every hyper move from their closed forms, mhP_x with its gene sum accumulated by the sweep — against the CPU chain, which
    evaluates the reference's per-gene sums.  1111 genes: ragged last tile; R = 5: the runtime-R instantiation."""
    ch, h = _start_pair(G, R, K, 11, thr_a)
    ch.run_fused(60, tune=True); h.run_generations(60, tune=True)
    _compare(ch, h, K, R)
    ch.run_fused(30, sample_frequency=2, tune=False); h.run_generations(30, tune=False, sample_frequency=2)
    _compare(ch, h, K, R)
    assert np.array_equal(h.get_alloc_trace(0), ch.alloc_trace(K))


@pytest.mark.parametrize("G,R,K,thr_a", [(1, 2, 2, (1.0, 6.0)), (7, 3, 1, (1.0,)), (33, 4, 2, (1.0, 6.0)), (64, 16, 2, (1.0, 6.0)), (97, 8, 3, (0.5, 3.0, 6.0))])
def test_device_resident_generations_on_tiny_gene_sets(G, R, K, thr_a):
    """Fewer tiles than warps, down to one gene: the grid shrinks to one worker CTA per tile (+ the steward), most lanes of the only
    tile are idle, every queue is empty from the start.  Same comparison with the CPU chain as above."""
    ch, h = _start_pair(G, R, K, 17, thr_a)
    ch.run_fused(40, tune=True); h.run_generations(40, tune=True)
    _compare(ch, h, K, R)
    ch.run_fused(12, sample_frequency=3, tune=False); h.run_generations(12, tune=False, sample_frequency=3)
    _compare(ch, h, K, R)
    assert np.array_equal(h.get_alloc_trace(0), ch.alloc_trace(K))


def test_device_resident_generations_on_a_small_grid(monkeypatch):
    """The same comparison with 2 CTAs (16 warps) for 94 tiles: several tiles per warp, two static rounds and then the dynamic
    tile queues — the paths a 1 M-gene shard takes on the full grid."""
    monkeypatch.setenv("ZZ_GEN_BLOCKS", "2")
    ch, h = _start_pair(3000, 4, 2, 13, (1.0, 6.0))
    ch.run_fused(50, tune=True); h.run_generations(50, tune=True)
    _compare(ch, h, 2, 4)
    ch.run_fused(20, sample_frequency=5, tune=False); h.run_generations(20, tune=False, sample_frequency=5)
    _compare(ch, h, 2, 4)


def test_run_sweeps_matches_the_oracle_sweeps_and_any_grid_size(monkeypatch):
    """zz_run_sweeps: n fused sweeps in one launch with the hyper-parameters held fixed.  State and decisions as n oracle sweeps;
    the fixed-point statistics are exact sums of the quantised per-gene terms, so they do not depend on the grid size."""
    import torch
    from _util import setup_pair, compare_state
    G, R = 5000, 4
    X, gl, hd, ho, hg, s, h = setup_pair(G, R, seed=5)
    stats = torch.zeros(h.nstat, dtype=torch.float64, device="cuda")
    h.set_stats_output_dev(stats.data_ptr())
    h.run_sweeps(7, 4)
    for i in range(4):
        O.sweep(ho, s, O.sweep_draws(1234, 0, 7 + i, 0, G), want_decisions=False)
    compare_state(h, s)
    st_o, st_g = O.suffstats(ho, s), stats.cpu().numpy()
    K1 = 3
    keep = [j for j in range(h.nstat) if j not in (3 * K1 + 7, 3 * K1 + 8)]        # the two cache sums are not produced here
    assert np.array_equal(st_g[:K1], st_o[:K1])
    assert np.allclose(st_g[keep], st_o[keep], rtol=1e-11, atol=2e-7)              # 32 fraction bits per gene term
    X2, gl2, hd2, ho2, hg2, s2, h2 = setup_pair(G, R, seed=5)
    monkeypatch.setenv("ZZ_GEN_BLOCKS", "5")
    stats2 = torch.zeros(h2.nstat, dtype=torch.float64, device="cuda")
    h2.set_stats_output_dev(stats2.data_ptr())
    h2.run_sweeps(7, 4)
    h2.sync()
    assert torch.equal(stats, stats2)
    a, b = h.get_state(0), h2.get_state(0)
    assert all(np.array_equal(a[k], b[k]) for k in a)


def test_device_resident_burnin_crosses_a_tuning_point():
    """tune_all every 400 generations (B:146): scalar and per-gene step sizes are re-tuned on the device."""
    G, R, K = 1200, 4, 2
    ch, h = _start_pair(G, R, K, 5, (1.0, 6.0))
    ch.run_fused(405, tune=True); h.run_generations(405, tune=True)
    sc, _ = h.chain_read()
    assert sc.t_tau != 0.05 and sc.gen == 405          # re-tuned away from the constructor's value (I:185-199)
    _compare(ch, h, K, R)


def test_run_generations_is_asynchronous_and_needs_chain_begin():
    G, R, K = 600, 4, 2
    X, gl, _ = synth.simulate(G, R, config_id=7)
    h = _capi.Handle(G, R, K, seed=3)
    h.upload_data(X, gl)
    with pytest.raises(_capi.ZigzagGpuError):
        h.run_generations(1)


def test_sample_rows_are_written_by_the_device():
    """zz_set_sample_sink: during a sampling segment the device appends the rows of the three log files to pinned host rings
    (parameters + calculate_lnl at every sample; Yg / variance_g of the candidate genes every yv_every-th sample, capped)."""
    G, R, K = 2500, 4, 2
    ch, h = _start_pair(G, R, K, 21, (1.0, 6.0))
    h.run_generations(30, tune=True)
    cand = np.arange(0, G, 7, dtype=np.int32)
    h.set_sample_sink(param_capacity=16, candidates=cand, yv_capacity=3, yv_every=2)
    h.run_generations(30, tune=False, sample_frequency=2)            # generations 30..59: samples at 30, 32, ..., 58
    prow, yrow, vrow = h.sample_rows()
    assert prow.shape == (15, 9 + R + 3 * K) and list(prow[:, 0]) == list(range(30, 60, 2))
    assert yrow.shape == (3, 1 + len(cand)) and list(yrow[:, 0]) == [32, 36, 40] and list(vrow[:, 0]) == [32, 36, 40]   # (gen/2) % 2 == 0, 3 rows max
    with pytest.raises(_capi.ZigzagGpuError):
        h.run_generations(10, tune=False, sample_frequency=2)          # only one parameter row left: refused before anything runs
    # the last generation (59) was not a sample: run one more sampled generation and compare the row with the state it describes
    h.run_generations(1, tune=False, sample_frequency=2)
    prow, _, _ = h.sample_rows()
    sc, hy = h.chain_read()
    last = prow[-1]
    assert last[0] == 60 and sc.gen == 61
    assert np.isclose(last[1], h.lnl(), rtol=1e-12)
    want = [hy.s0, hy.s1, hy.tau, *hy.alpha_r[:R], hy.spike_probability, hy.inactive_means, hy.inactive_variances, *hy.active_means[:K],
            *hy.active_variances[:K], hy.weight_active, *hy.weight_within_active[:K]]
    assert np.array_equal(last[2:], np.array(want))
    # a fresh sink: per-gene rows of the current state, in the masked form the log files use
    h.set_sample_sink(param_capacity=2, candidates=cand, yv_capacity=2, yv_every=1)
    h.run_generations(1, tune=False, sample_frequency=1)
    _, yrow, vrow = h.sample_rows()
    st = h.get_state(0)
    sp = st["inactive_spike_allocation"][cand] == 1
    assert np.array_equal(yrow[0, 1:], np.where(sp, -10.0, st["Yg"][cand]))
    assert np.array_equal(vrow[0, 1:], np.where(sp, 0.0, st["variance_g"][cand]))
    h.clear_sample_sink()


@pytest.mark.parametrize("G,C_", [(1800, 3), (500, 17)])
def test_batched_chains_run_device_resident_and_match_their_oracle_chains(G, C_):
    """A handle with several chains (shared data, per-chain state and hyper-parameters, chain index in every Philox counter):
    zz_run_generations advances all of them in the same launches — one stage block, one reduction slice and one gate per chain.
    17 chains exceed the fused statistics epilogue (296 doubles): the sweep's per-CTA rows are reduced per chain instead."""
    R, K = 4, 2
    X, gl, _ = synth.simulate(G, R, config_id=333)
    pri = O.Priors.defaults(-1.0, (1.0, 6.0))
    chains = [O.Chain(X, gl, pri, K, 100 + c) for c in range(C_)]          # different seeds: different initial states / hypers
    h = _capi.Handle(G, R, K, num_chains=C_, seed=2024)
    h.upload_data(X, gl)
    scs = []
    for c, ch in enumerate(chains):
        st, hy = ch.state_arrays(), ch.hyper
        hg = _capi.Hyper(); C.memmove(C.byref(hg), C.byref(hy), C.sizeof(_capi.Hyper))
        h.set_hyper(hg, c)
        h.set_state(c, **{k: st[k] for k in ("Yg", "variance_g", "alloc_active_inactive", "alloc_within_active", "inactive_spike_allocation",
                                             "XgLikelihood", "variance_g_probability", "tuningParam_yg", "tuningParam_variance_g")})
        ch.set_stream(2024, c)                                             # from here on: the handle's key, this chain's index
        scs.append(ch.scalars())
    # the constructors consumed different numbers of uniforms; batched chains only need gen / step / n_samples in lock-step
    h.reset_windows(); h.reset_alloc_trace()
    h.chain_begin(pri, scs)
    for ch in chains:
        ch.run_fused(40, tune=True); ch.run_fused(12, sample_frequency=3, tune=False)
    h.run_generations(40, tune=True); h.run_generations(12, tune=False, sample_frequency=3)
    sc_g, hy_g = h.chain_read()
    for c, ch in enumerate(chains):
        sc_o, hy_o = ch.scalars(), ch.hyper
        assert sc_g[c].hyper_ctr == sc_o.hyper_ctr and sc_g[c].gen == sc_o.gen == 52
        for f in ("s0", "s1", "tau", "spike_probability", "inactive_means", "inactive_variances", "weight_active"):
            assert np.isclose(getattr(hy_g[c], f), getattr(hy_o, f), rtol=1e-9, atol=1e-13), (c, f)
        assert np.allclose(list(hy_g[c].alpha_r)[:R], list(hy_o.alpha_r)[:R], rtol=1e-9)
        g, o = h.get_state(c), ch.state_arrays()
        assert np.array_equal(g["alloc_active_inactive"], o["alloc_active_inactive"])
        assert np.allclose(g["Yg"], o["Yg"], rtol=1e-9, atol=1e-12) and np.allclose(g["variance_g"], o["variance_g"], rtol=1e-9)
        assert np.array_equal(h.get_alloc_trace(c), ch.alloc_trace(K))
    assert hy_g[0].tau != hy_g[1].tau


def test_device_loop_reproduces_the_golden_chain_trajectory():
    """The committed trajectory of the fused-schedule chain (tests/golden/chain_g400.npz): started from the same constructor state,
    zz_run_generations lands on the same hyper-parameters, tuning parameters, stream positions and allocation-trace totals."""
    import os
    want = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "chain_g400.npz"))
    G, R, K = 400, 4, 2
    X, gl, _ = synth.simulate(G, R, config_id=4242)
    pri = O.Priors.defaults(-1.0, (1.0, 6.0))
    ch = O.Chain(X, gl, pri, K, 31)                                      # the constructor only: initial state and stream positions
    st, hy = ch.state_arrays(), ch.hyper
    hg = _capi.Hyper(); C.memmove(C.byref(hg), C.byref(hy), C.sizeof(_capi.Hyper))
    h = _capi.Handle(G, R, K, seed=31)
    h.upload_data(X, gl); h.set_hyper(hg, 0)
    h.set_state(0, **{k: st[k] for k in ("Yg", "variance_g", "alloc_active_inactive", "alloc_within_active", "inactive_spike_allocation",
                                         "XgLikelihood", "variance_g_probability", "tuningParam_yg", "tuningParam_variance_g")})
    h.reset_windows(); h.reset_alloc_trace()
    h.chain_begin(pri, ch.scalars())
    for tag, n, tune in (("burn", 420, True), ("samp", 30, False)):
        h.run_generations(n, tune=tune, sample_frequency=3)
        sc, hyg = h.chain_read()
        got = np.array([hyg.s0, hyg.s1, hyg.tau, hyg.spike_probability, hyg.inactive_means, hyg.inactive_variances, hyg.weight_active,
                        *hyg.alpha_r[:R], *hyg.active_means[:K], *hyg.active_variances[:K], *hyg.weight_within_active[:K]])
        assert np.allclose(got, want[tag + "_hyper"], rtol=1e-9, atol=1e-12), tag
        assert [sc.hyper_ctr, sc.gen, sc.step, sc.n_samples] == list(want[tag + "_ctr"]), tag
        tun = np.array([sc.t_im, sc.t_av, sc.t_s0, sc.t_s1, sc.t_tau, sc.t_s0tau, *sc.t_am[:K], *sc.t_alpha[:R]])
        assert np.allclose(tun, want[tag + "_tuning"], rtol=1e-9), tag
        s = h.get_state(0)
        chk = np.array([s["Yg"].sum(), s["variance_g"].sum(), s["alloc_active_inactive"].sum(), s["inactive_spike_allocation"].sum(),
                        s["tuningParam_yg"].sum()])
        assert np.allclose(chk, want[tag + "_state"], rtol=1e-9), tag
    assert np.array_equal(h.get_alloc_trace(0).sum(axis=1), want["trace"])


@pytest.mark.parametrize("name", ["1m_x16", "200k_x8", "20k_x2", "64ch_60k_x4"])
def test_device_loop_full_size_invariants(name):
    """Every BASELINE.json workload at its full size through the device-resident loop (1 M x 16; 200 k x 8 with gene-length-dependent
    detection; 20 k x 2; 64 batched chains x 60 k x 4): counts add up, nothing non-finite, every chain's hyper-parameters stay in the
    neighbourhood of the simulation truth they started from, the allocation trace counts every gene once per sample."""
    X, gl, hd, st, C_ = synth.make_config(name)
    G, R = X.shape
    hg = _capi.Hyper.make(hd["alpha_r"], hd["active_means"], hd["active_variances"], hd["weight_within_active"],
                          **{k: float(hd[k]) for k in ("s0", "s1", "tau", "spike_probability", "inactive_means", "inactive_variances",
                                                       "weight_active", "beta", "temperature", "variance_g_upper_bound")})
    h = _capi.Handle(G, R, 2, num_chains=C_, seed=8)
    h.upload_data(X, gl)
    for c in range(C_):
        h.set_hyper(hg, c); h.set_state(c, **st)
    h.refresh_caches(); h.reset_alloc_trace()
    thr_a = (1.0, 6.0)
    h.chain_begin(_capi.Priors.defaults(-1.0, thr_a), [_capi.ChainScalars.initial(hd["active_means"], thr_a, R) for _ in range(C_)])
    h.run_sweeps(1 << 40, 400)                      # the per-gene state equilibrates under the true hyper-parameters first
    h.run_generations(30, tune=True)
    h.run_generations(10, tune=False, sample_frequency=2)
    scs, hys = h.chain_read()
    if C_ == 1:
        scs, hys = [scs], [hys]
    assert h.nan_flag() == 0
    # the genes pin every hyper-parameter (posterior sd ~ 1 / sqrt(G)): started at the simulation truth with an equilibrated per-gene
    # state, a chain stays there — a wrong gene sum in any hyper move would let it walk away at once
    f = min(10.0, math.sqrt(1e6 / G)) * (2.0 if R == 2 else 1.0)
    for sc, hy in zip(scs, hys):
        assert sc.gen == 40 and sc.n_samples == 1 + 5
        assert abs(hy.tau - hd["tau"]) < 0.03 * f and abs(hy.s0 - hd["s0"]) < 0.03 * f and abs(hy.s1 - hd["s1"]) < 0.01 * f, (hy.tau, hy.s0, hy.s1)
        assert abs(hy.weight_active - hd["weight_active"]) < 0.005 * f and abs(hy.spike_probability - hd["spike_probability"]) < 0.005 * f
        assert np.allclose(hy.alpha_r[:R], hd["alpha_r"], rtol=0.01 * f) and abs(hy.inactive_means - hd["inactive_means"]) < 0.02 * f
        assert np.allclose(hy.active_means[:2], hd["active_means"], atol=0.05 * f) and abs(hy.inactive_variances - 2.0) < 0.05 * f
    if C_ > 1:                                       # independent chains: distinct Philox streams, so distinct trajectories
        assert len({hy.tau for hy in hys}) == C_
    for c in (0, C_ - 1):
        s = h.get_state(c)
        assert np.isfinite(s["Yg"]).all() and (s["variance_g"] > 0).all()
        tr = h.get_alloc_trace(c)
        assert tr.sum() == 6 * G and np.array_equal(tr.sum(axis=0), np.full(G, 6))
    st_ = h.suffstats()
    assert all(st_[c][:3].sum() == G for c in range(C_))
