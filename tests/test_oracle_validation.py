"""The reference's own validation recipe that needs no R output: the beta = 0 chain of scripts/test_script.R:42-53.  With the data
likelihood switched off (P:121-125) and one never-detected gene, the marginal posterior of every hyper-parameter is its PRIOR,
whose moments are known in closed form (constructor defaults, R/zigzagInitialize.R:7-41).  The oracle's schedule-faithful chain
(15 weighted moves per generation, oracle/zz_oracle.c:zzo_chain_burnin / zzo_chain_mcmc) must reproduce them: a pin of all ten
scalar moves — proposals, Hastings ratios, prior ratios, accept rules — against the model itself rather than against a run of R.
tests/test_gpu_f4.py holds the device-resident loop to the same bar."""
import ctypes as C
import math

import numpy as np

from oracle import oracle as O


def _ess(x):
    x = np.asarray(x, dtype=np.float64) - np.mean(x)
    n = len(x)
    f = np.fft.rfft(x, 2 * n)
    ac = np.fft.irfft(f * np.conj(f))[:n] / (x.var() * n)
    s = 0.0
    for k in range(1, n // 2):
        if ac[k] <= 0.02:
            break
        s += ac[k]
    return n / (1 + 2 * s)


def test_oracle_chain_with_beta_zero_recovers_the_priors():
    X, gl = np.full((1, 2), -np.inf), np.ones(1)
    ch = O.Chain(X, gl, O.Priors.defaults(0.0, (0.0, 5.0)), 2, seed=3)
    h = ch.hyper
    h.beta = 0.0
    L = O.lib()
    L.zzo_chain_state.restype, L.zzo_chain_state.argtypes = C.c_void_p, [C.c_void_p]
    L.zzo_refresh_caches(C.byref(h), C.c_void_p(L.zzo_chain_state(ch.ptr)))
    ch.burnin(4000)
    rows = []
    for _ in range(8000):
        ch.mcmc(40, 40)
        h = ch.hyper
        rows.append([h.tau, h.s0, abs(h.s1), h.alpha_r[0], h.alpha_r[1], 0.0 - h.inactive_means, h.active_means[0] - 0.0, h.active_means[1] - 5.0,
                     h.weight_active, h.weight_within_active[0], math.log10(h.inactive_variances)])
    x = np.array(rows)
    lo, hi = math.log10(0.01), math.log10(5)
    pri = [("tau ~ Gamma(1, 1)", 1.0, 1.0), ("s0 ~ N(-1, 2)", -1.0, 2.0), ("|s1| ~ Gamma(1, 2)", 0.5, 0.5), ("alpha_1 ~ Gamma(1, 1/10)", 10.0, 10.0),
           ("alpha_2", 10.0, 10.0), ("threshold_i - inactive_mean ~ Gamma(1, 1/3)", 3.0, 3.0), ("active_means_dif_1 ~ Gamma(1, 1/3)", 3.0, 3.0),
           ("active_means_dif_2", 3.0, 3.0), ("weight_active ~ Beta(2, 2)", 0.5, math.sqrt(0.05)), ("weight_within_active_1 ~ Beta(1, 1)", 0.5, math.sqrt(1 / 12)),
           ("log10 variance ~ U(log10 .01, log10 5)", (lo + hi) / 2, (hi - lo) / math.sqrt(12))]
    for j, (name, mu, sd) in enumerate(pri):
        ess = max(_ess(x[:, j]), 20.0)
        z = (x[:, j].mean() - mu) / (sd / math.sqrt(ess))
        assert abs(z) < 4.5, f"{name}: sample mean {x[:, j].mean():.4f} vs prior mean {mu} (z = {z:.2f}, ESS {ess:.0f})"
        assert 0.8 < x[:, j].std() / sd < 1.25, f"{name}: sample sd {x[:, j].std():.4f} vs prior sd {sd:.4f}"
    # quantiles of two of them against the prior's CDF (a stronger statement than two moments)
    q = np.quantile(x[:, 0], [0.1, 0.5, 0.9])                                # Exp(1): -log(1 - p)
    assert np.allclose(q, -np.log(1 - np.array([0.1, 0.5, 0.9])), rtol=0.12)
    q = np.quantile(x[:, 9], [0.1, 0.5, 0.9])                                # U(0, 1)
    assert np.allclose(q, [0.1, 0.5, 0.9], atol=0.03)
