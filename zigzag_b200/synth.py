"""Deterministic synthetic inputs for the benchmark / parity configurations.

Follows the reference's generative simulator scripts/simulate_data.cf.r:46-126 and its initial-state recipe
R/zigzagInitialize.R:416-504, with the hyper-parameters set to the simulation truth so throughput is measured in
the stationary regime.  Pure NumPy; nothing here touches the GPU or the oracle.
"""
import numpy as np

SEED0 = 20261019

TRUTH = dict(weight_active=0.6, weight_within_active=(0.95, 0.05), spike_probability=0.1, inactive_means=-2.0,
             active_means=(2.0, 7.0), shared_variance=2.0, tau=0.8, s0=-1.5, s1=-0.1)
THRESHOLD_I, THRESHOLD_A = -1.0, (1.0, 6.0)   # scripts/cmd_zigzag_sim_batch.r:17

CONFIGS = {            # name -> (config_id, G, R, chains)
    "20k_x2": (1, 20_000, 2, 1),
    "200k_x8": (2, 200_000, 8, 1),
    "1m_x16": (3, 1_000_000, 16, 1),
    "64ch_60k_x4": (4, 60_000, 4, 64),
}


def truth_hyper_dict(R):
    t = TRUTH
    return dict(s0=t["s0"], s1=t["s1"], tau=t["tau"], alpha_r=np.linspace(10.0, 20.0, R),
                spike_probability=t["spike_probability"], inactive_means=t["inactive_means"],
                inactive_variances=t["shared_variance"], active_means=np.array(t["active_means"]),
                active_variances=np.full(2, t["shared_variance"]), weight_active=t["weight_active"],
                weight_within_active=np.array(t["weight_within_active"]), beta=1.0, temperature=1.0,
                variance_g_upper_bound=float("inf"))


def simulate(G, R, config_id=0, unit_gene_lengths=False):
    """Returns (Xg [G,R] log-TPM with -Inf = not detected, gene_lengths [G] unit mean, truth dict)."""
    rng = np.random.Generator(np.random.PCG64(SEED0 + config_id))
    t = TRUTH
    gl = rng.lognormal(np.log(2000.0), 0.5, G)
    gl = gl / gl.mean()                                   # I:145
    if unit_gene_lengths:
        gl = np.ones(G)
    alpha = np.linspace(10.0, 20.0, R)
    active = rng.random(G) < t["weight_active"]
    in_spike = (~active) & (rng.random(G) < t["spike_probability"])
    comp = np.where(active, 1 + (rng.random(G) >= t["weight_within_active"][0]).astype(np.int64), 0)
    mu = np.where(comp == 0, t["inactive_means"], np.where(comp == 1, t["active_means"][0], t["active_means"][1]))
    Y = mu + np.sqrt(t["shared_variance"]) * rng.standard_normal(G)
    sig2 = np.exp(t["s0"] + t["s1"] * Y + t["tau"] + np.sqrt(t["tau"]) * rng.standard_normal(G))  # simulate_data.cf.r:15-21
    X = np.empty((G, R), order="F")
    for r in range(R):
        x = Y + np.sqrt(sig2) * rng.standard_normal(G)
        p_nd = np.exp(-alpha[r] * gl * np.exp(Y))          # simulate_data.cf.r:24-27, 108-114
        x[rng.random(G) < p_nd] = -np.inf
        x[in_spike] = -np.inf
        X[:, r] = x
    return X, gl, dict(Y=Y, sigma2=sig2, comp=comp, in_spike=in_spike, alpha_r=alpha)


def _xg_lik_is_neg_inf(X, y, gl, alpha):
    """True where computeXgLikelihood would be -Inf: a non-detected library whose 1 - p_x rounds to 0 (P:82)."""
    bad = np.zeros(len(y), dtype=bool)
    t = gl * np.exp(y)
    for r in range(X.shape[1]):
        p = 1 - np.exp(-(t * alpha[r]))
        bad |= np.isneginf(X[:, r]) & ((1 - p) == 0)
    return bad


def initial_state(X, gl, hyper, seed):
    """Initial per-gene state per R/zigzagInitialize.R:416-504 given hyper-parameters `hyper` (dict)."""
    G, R = X.shape
    rng = np.random.Generator(np.random.PCG64(seed))
    det = ~np.isneginf(X)
    ndet = det.sum(1)
    nd_spike = (ndet == 0)
    spike = nd_spike.astype(np.int32)                                       # I:418
    ai = (rng.random(G) < hyper["weight_active"]).astype(np.int32)         # I:367
    ai[nd_spike] = 0                                                        # I:422
    ww = np.asarray(hyper["weight_within_active"], dtype=np.float64)
    within = (1 + np.searchsorted(np.cumsum(ww)[:-1], rng.random(G), side="right")).astype(np.int32)  # I:372
    Xz = np.where(det, X, 0.0)
    im, iv = hyper["inactive_means"], hyper["inactive_variances"]
    with np.errstate(invalid="ignore", divide="ignore"):
        rwm = np.where(ndet > 0, Xz.sum(1) / np.maximum(ndet, 1), im + np.sqrt(iv) * rng.standard_normal(G))   # I:431-438
        d2 = np.where(det, (X - rwm[:, None]) ** 2, 0.0).sum(1)
        full = ndet == R
        rv_full = d2[full] / (R - 1)
        if len(rv_full) > 1:
            m, s = np.log(rv_full.mean()), np.sqrt(rv_full.var(ddof=1))     # I:446
        else:                                                               # data without detections: Yg = rwm (I:452-455)
            m, s = -np.inf, 0.0
        rwv = np.where(ndet >= 2, d2 / np.maximum(ndet - 1, 1), np.exp(m + s * rng.standard_normal(G)))        # I:441-448
    Yg = rwm + np.sqrt(rwv) * rng.standard_normal(G)                        # I:449
    vg = np.exp(0.5 + 0.2 * rng.standard_normal(G))                         # I:465
    alpha = np.asarray(hyper["alpha_r"], dtype=np.float64)
    for _ in range(1000):                                                   # I:480-502
        bad = _xg_lik_is_neg_inf(X, Yg, gl, alpha) & ~nd_spike
        if not bad.any():
            break
        n = int(bad.sum())
        Yg[bad] = im + np.sqrt(iv) * rng.standard_normal(n)
        Sg = np.exp(hyper["s0"] + hyper["s1"] * Yg[bad])
        vg[bad] = np.exp(np.log(Sg) + hyper["tau"] + np.sqrt(hyper["tau"]) * rng.standard_normal(n))
    return dict(Yg=Yg, variance_g=vg, alloc_active_inactive=ai, alloc_within_active=within,
                inactive_spike_allocation=spike, tuningParam_yg=np.full(G, 0.5), tuningParam_variance_g=np.full(G, 0.5))


def make_config(name, G=None):
    """(X, gl, hyper dict, state dict, chains) for one of the BASELINE.json configurations."""
    cid, G0, R, chains = CONFIGS[name]
    G = G or G0
    X, gl, _ = simulate(G, R, cid)
    hyper = truth_hyper_dict(R)
    state = initial_state(X, gl, hyper, SEED0 + 1000 + cid)
    return X, gl, hyper, state, chains
