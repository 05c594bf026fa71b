"""SURVEY §8 (f4) and the reference's own validation recipes, on the GPU:
  * the constructor-side initial state on the device (zz_init_state, R/zigzagInitialize.R:416-504) against the oracle's
    restatement on the same Philox draws, and its independence of the sharding;
  * Metropolis-coupled chains and the stepping-stone ladder on the chain axis (R/zigzagDevelopment.R:9-208);
  * the beta = 0 prior-recovery chain of scripts/test_script.R:42-53: with the data switched off every hyper-parameter must
    reproduce its prior, whose moments are known in closed form — a check of all nine scalar moves that needs no R;
  * parameter recovery on simulator output with the prior settings of scripts/cmd_zigzag_sim_batch.r:16-26."""
import math
import numpy as np
import pytest

from oracle import oracle as O
from zigzag_b200 import _capi, synth
from zigzag_b200.zigzag import zigzag
from _util import make_case, hyper_pair, oracle_state

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("G,R,K", [(3000, 4, 2), (1500, 16, 3), (700, 2, 1)])
def test_device_initial_state_matches_the_oracle(G, R, K):
    X, gl, hd, st = make_case(G, R, 11, K, edge_cases=False)
    hd = dict(hd); hd["alpha_r"] = np.linspace(0.5, 3.0, R)          # small alpha_r: -Inf likelihoods are common, the re-draw loop runs
    ho, hg = hyper_pair(hd)
    m, sd = O.rowvar_moments(X)
    pinned = (np.random.Generator(np.random.PCG64(4)).random(G) < 0.02).astype(np.int32)
    pinned[np.isneginf(X).all(1)] = 0
    s = oracle_state(X, gl, st, pinned)
    O.init_state(ho, s, 99, 0, 5, 0, m, sd)
    h = _capi.Handle(G, R, K, seed=99)
    h.upload_data(X, gl); h.set_hyper(hg, 0); h.set_state(0, pinned_active=pinned)
    h.init_state(5, m, sd)
    g = h.get_state(0)
    for k in ("alloc_active_inactive", "inactive_spike_allocation"):
        assert np.array_equal(g[k], getattr(s, k)), k
    act = s.alloc_active_inactive == 1
    assert np.array_equal(g["alloc_within_active"][act], s.alloc_within_active[act])
    assert np.all(g["alloc_active_inactive"][pinned == 1] == 1)
    assert np.allclose(g["Yg"], s.Yg, rtol=1e-12, atol=1e-13) and np.allclose(g["variance_g"], s.variance_g, rtol=1e-11)
    fin = np.isfinite(s.XgLikelihood)
    assert fin.all() and np.isfinite(g["XgLikelihood"]).all()          # I:480-502: no gene is left at zero likelihood
    assert np.allclose(g["XgLikelihood"], s.XgLikelihood, rtol=1e-10, atol=1e-9)
    assert np.allclose(g["variance_g_probability"], s.variance_g_probability, rtol=1e-10, atol=1e-9)
    assert np.all(g["tuningParam_yg"] == 0.5) and np.array_equal(h.window_sums(0)[0], np.full(G, 23))
    # a gene shard initialises to the same state as the same genes inside the whole set (Philox is keyed by the global gene index)
    lo, hi = G // 3, min(G, G // 3 + 501)
    hs = _capi.Handle(hi - lo, R, K, seed=99, gene_offset=lo)
    hs.upload_data(np.asfortranarray(X[lo:hi]), gl[lo:hi]); hs.set_hyper(hg, 0); hs.set_state(0, pinned_active=pinned[lo:hi])
    hs.init_state(5, m, sd)
    gs = hs.get_state(0)
    for k in ("Yg", "variance_g", "alloc_active_inactive", "inactive_spike_allocation"):
        assert np.array_equal(gs[k], g[k][lo:hi]), k


def _sim(G, R, seed):
    X, gl, truth = synth.simulate(G, R, seed)
    return np.where(np.isneginf(X), 0.0, np.exp(X)), gl, truth


def test_mc3_swaps_temperatures_on_the_chain_axis(tmp_path):
    data, gl, _ = _sim(2000, 4, 31)
    mm = zigzag(data, gl, threshold_i=-1.0, threshold_a=(1.0, 6.0), seed=7, output_directory=str(tmp_path), init="device")
    mm.burnin(sample_frequency=50, ngen=400, write_to_files=False, schedule="device")
    acc, prop = mm.mc3(nchains=4, burnin_ngen=200, mcmc_ngen=300, sample_frequency=10, swap_every=1, mcmcprefix="mc3")
    assert prop == 300 * 3 and 0 < acc <= prop
    assert sorted(mm.mc3_slots) == [0, 1, 2, 3] and mm.mc3_temperatures[0] == 1.0 and mm.temperature == 1.0
    pa = mm.mc3_prob_active
    assert pa.shape == (2000,) and pa.min() >= 0 and pa.max() <= 1 and 0.4 < pa.mean() < 0.8
    rows = open(tmp_path / "mc3_mcmc_output" / "mc3_model_parameters.log").read().splitlines()
    assert len(rows) == 1 + 30 and all(len(r.split("\t")) == 19 for r in rows)
    assert mm.h.nan_flag() == 0 and np.isfinite(mm.calculate_lnl())
    # equal temperatures: the swap ratio is exactly 0, every proposal is accepted (log u < 0)
    acc0, prop0 = mm.mc3(nchains=3, burnin_ngen=50, mcmc_ngen=40, dT=0.0, write_to_files=False)
    assert acc0 == prop0 == 40 * 2


def test_stepping_stone_ladder_on_the_chain_axis(tmp_path):
    data, gl, _ = _sim(1500, 4, 32)
    mm = zigzag(data, gl, threshold_i=-1.0, threshold_a=(1.0, 6.0), seed=8, output_directory=str(tmp_path), init="device", write_settings=False)
    mm.burnin(sample_frequency=50, ngen=400, write_to_files=False, schedule="device")
    total, terms, betas = mm.estimateMarginalLikelihood(burnin_gen=300, mcmc_gen=600, ss_sample_frequency=10, stones=6)
    assert len(terms) == 5 and np.all(np.diff(betas) > 0) and betas[-1] == 1.0 and abs(betas[0] - (1 / 6) ** 5) < 1e-15
    assert np.isfinite(total) and total < 0 and total == pytest.approx(sum(terms))
    # deterministic given the seeds; a second object with the same seeds repeats the estimate
    m2 = zigzag(data, gl, threshold_i=-1.0, threshold_a=(1.0, 6.0), seed=8, output_directory=str(tmp_path), init="device", write_settings=False)
    m2.burnin(sample_frequency=50, ngen=400, write_to_files=False, schedule="device")
    assert m2.estimateMarginalLikelihood(burnin_gen=300, mcmc_gen=600, ss_sample_frequency=10, stones=6)[0] == total


def _ess(x):
    """Effective sample size from the initial positive sequence of autocorrelations."""
    x = np.asarray(x, dtype=np.float64) - np.mean(x)
    n = len(x)
    if x.var() == 0:
        return 1.0
    f = np.fft.rfft(x, 2 * n)
    ac = np.fft.irfft(f * np.conj(f))[:n] / (x.var() * n)
    s = 0.0
    for k in range(1, n // 2):
        if ac[k] <= 0.02:
            break
        s += ac[k]
    return n / (1 + 2 * s)


def test_beta_zero_chain_recovers_the_priors(tmp_path):
    """scripts/test_script.R:42-53: one gene without any detection, beta = 0."""
    mm = zigzag(np.zeros((1, 2)), None, num_active_components=2, threshold_i=0.0, threshold_a=(0.0, 5.0), beta=0, seed=12,
                output_directory=str(tmp_path), write_settings=False)
    mm.burnin(sample_frequency=20, ngen=4000, write_to_files=False, schedule="device")
    rows = []
    for _ in range(3000):                                        # 3000 samples, 40 generations apart
        mm._device_segment(40, False, sample_frequency=1 << 30)
        rows.append([mm.tau, mm.s0, abs(mm.s1), mm.alpha_r[0], mm.alpha_r[1], mm.threshold_i - mm.inactive_means, mm.active_means_dif[0],
                     mm.active_means_dif[1], mm.weight_active, mm.weight_within_active[0], math.log10(mm.inactive_variances)])
    x = np.array(rows)
    lo, hi = math.log10(0.01), math.log10(5)
    # (name, prior mean, prior sd) — R/zigzagInitialize.R:7-41 defaults
    pri = [("tau ~ Gamma(1, 1)", 1.0, 1.0), ("s0 ~ N(-1, 2)", -1.0, 2.0), ("|s1| ~ Gamma(1, 2)", 0.5, 0.5),
           ("alpha_1 ~ Gamma(1, 1/10)", 10.0, 10.0), ("alpha_2", 10.0, 10.0), ("threshold_i - inactive_mean ~ Gamma(1, 1/3)", 3.0, 3.0),
           ("active_means_dif_1 ~ Gamma(1, 1/3)", 3.0, 3.0), ("active_means_dif_2", 3.0, 3.0), ("weight_active ~ Beta(2, 2)", 0.5, math.sqrt(0.05)),
           ("weight_within_active_1 ~ Beta(1, 1)", 0.5, math.sqrt(1 / 12)), ("log10 variance ~ U(log10 .01, log10 5)", (lo + hi) / 2, (hi - lo) / math.sqrt(12))]
    for j, (name, mu, sd) in enumerate(pri):
        ess = max(_ess(x[:, j]), 20.0)
        z = (x[:, j].mean() - mu) / (sd / math.sqrt(ess))
        assert abs(z) < 4.5, f"{name}: sample mean {x[:, j].mean():.4f} vs prior mean {mu} (z = {z:.2f}, ESS {ess:.0f})"
        assert 0.7 < x[:, j].std() / sd < 1.35, f"{name}: sample sd {x[:, j].std():.4f} vs prior sd {sd:.4f}"
    assert mm.h.nan_flag() == 0


def test_parameters_are_recovered_from_simulated_data(tmp_path):
    """scripts/simulate_data.cf.r + scripts/cmd_zigzag_sim_batch.r:16-26 (their prior settings): the posterior concentrates on the
    simulation truth and the P(active) classification recovers the simulated component memberships."""
    G, R = 6000, 4
    data, gl, truth = _sim(G, R, 33)
    mm = zigzag(data, gl * 2000.0, shared_variance=True, threshold_a=(1, 6), threshold_i=-1, num_active_components=2,
                weight_active_shape_1=10, weight_active_shape_2=10, spike_prior_shape_1=2, spike_prior_shape_2=8, tau_shape=5, tau_rate=20,
                s0_mu=-1, s0_sigma=1 / 5, s1_shape=1, s1_rate=10, alpha_r_shape=15, alpha_r_rate=1, inactive_means_prior_shape=1,
                inactive_means_prior_rate=1, active_means_dif_prior_shape=1, active_means_dif_prior_rate=1, shared_variance_prior_min=1,
                shared_variance_prior_max=3, seed=21, output_directory=str(tmp_path), write_settings=False, init="device")
    mm.burnin(sample_frequency=10, ngen=5000, write_to_files=False, schedule="device")
    keep = []
    for _ in range(200):
        mm._device_segment(100, False, sample_frequency=10)
        keep.append([mm.weight_active, mm.spike_probability, mm.inactive_means, mm.active_means[0], mm.active_means[1], mm.inactive_variances,
                     mm.tau, mm.s0, mm.s1, mm.weight_within_active[0], *mm.alpha_r])
    post = np.array(keep).mean(0)
    t = synth.TRUTH
    want = [t["weight_active"], t["spike_probability"], t["inactive_means"], t["active_means"][0], t["active_means"][1], t["shared_variance"],
            t["tau"], t["s0"], t["s1"], t["weight_within_active"][0], *truth["alpha_r"]]
    tol = [0.03, 0.03, 0.15, 0.12, 0.6, 0.2, 0.12, 0.15, 0.03, 0.02] + [2.5] * R   # component 2 holds 3 % of the genes and overlaps the tail of component 1
    names = ["weight_active", "spike_probability", "inactive_mean", "active_mean_1", "active_mean_2", "variance", "tau", "s0", "s1",
             "weight_within_1"] + [f"alpha_{r}" for r in range(R)]
    report = "; ".join(f"{n} {p:.3f} (truth {w:.3f})" for n, p, w in zip(names, post, want))
    print(report)
    bad = [n for n, p, w, e in zip(names, post, want, tol) if not abs(p - w) < e]
    assert not bad, f"not recovered: {bad} — {report}"
    pa = mm.prob_active()
    called = pa > 0.5
    is_active = truth["comp"] > 0
    # the level-2 components overlap (means -2 and 2, sd 1.41): even the Bayes rule that SEES the true Yg and parameters misclassifies ~7.5 %
    t_sd = math.sqrt(t["shared_variance"])
    dens = lambda y, m: np.exp(-0.5 * ((y - m) / t_sd) ** 2)
    w1, w2 = t["weight_within_active"]
    bayes = t["weight_active"] * (w1 * dens(truth["Y"], t["active_means"][0]) + w2 * dens(truth["Y"], t["active_means"][1])) > \
        (1 - t["weight_active"]) * (1 - t["spike_probability"]) * dens(truth["Y"], t["inactive_means"])
    bayes_acc = (bayes == is_active)[~truth["in_spike"]].mean()
    acc = (called == is_active).mean()
    print(f"classification accuracy {acc:.4f}, Bayes rule on the true Yg {bayes_acc:.4f}")
    assert acc > bayes_acc - 0.03 and acc > 0.9
    assert mm.h.nan_flag() == 0


def test_non_shared_variance_model_recovers_its_priors(tmp_path):
    """shared_variance = FALSE, shared_active_variance = FALSE (R/zigzagProposals.R:473-509 mhInactiveVariances, 707-766 mhActiveVariances;
    slot weights I:228-237), host-driven fused schedule.  Same recipe as above: beta = 0, one gene without detections — the three
    variances must reproduce their three independent log10-uniform priors."""
    mm = zigzag(np.zeros((1, 2)), None, num_active_components=2, threshold_i=0.0, threshold_a=(0.0, 5.0), beta=0, seed=14, shared_variance=False,
                shared_active_variance=False, inactive_variances_prior_min=0.1, inactive_variances_prior_max=10.0, active_variances_prior_min=0.01,
                active_variances_prior_max=2.0, output_directory=str(tmp_path), write_settings=True)
    assert mm.proposal_list[4] == mm.mhInactiveVariances and mm.proposal_list[6] == mm.mhActiveVariances
    assert mm.proposal_probs[4] == 7 and mm.proposal_probs[6] == 7 and len(set(mm.active_variances)) == 2
    rows = dict(l.rstrip("\n").split("\t") for l in open(tmp_path / "hyperparameter_settings.txt"))
    assert rows["inactive_variances_prior_max"] == "10" and rows["active_variances_prior_min"] == "0.01" and "shared_variance_prior_min" not in rows
    with pytest.raises(NotImplementedError):
        mm.burnin(ngen=10, write_to_files=False, schedule="device")
    mm.burnin(sample_frequency=100, ngen=1200, write_to_files=False, schedule="fused")
    assert mm.inactive_variance_tuningParam != 0.5                       # tune_all reached the new parameters
    x = []
    for i in range(12000):
        mm._generation(False, "fused")
        if i % 4 == 0:
            x.append([math.log10(mm.inactive_variances), math.log10(mm.active_variances[0]), math.log10(mm.active_variances[1])])
    x = np.array(x)
    for j, (lo, hi) in enumerate([(-1.0, 1.0), (-2.0, math.log10(2.0)), (-2.0, math.log10(2.0))]):
        mu, sd = (lo + hi) / 2, (hi - lo) / math.sqrt(12)
        ess = max(_ess(x[:, j]), 20.0)
        assert abs(x[:, j].mean() - mu) < 4.5 * sd / math.sqrt(ess), (j, x[:, j].mean(), mu, ess)
        assert 0.8 < x[:, j].std() / sd < 1.2 and x[:, j].min() >= lo and x[:, j].max() <= hi
    assert abs(np.corrcoef(x[:, 1], x[:, 2])[0, 1]) < 0.15              # the two active variances move independently
    # the reference schedule draws the two moves with their weights
    mm.burnin(sample_frequency=100, ngen=200, write_to_files=False, schedule="reference")
    assert mm.h.nan_flag() == 0
