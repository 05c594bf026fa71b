"""Shared helpers for the parity tests: builds the same seeded case on both sides (oracle = checker, GPU = product)."""
import numpy as np
from oracle import oracle as O
from zigzag_b200 import _capi, synth

SCALARS = ("s0", "s1", "tau", "spike_probability", "inactive_means", "inactive_variances", "weight_active", "beta",
           "temperature", "variance_g_upper_bound")


def hyper_pair(hd):
    kw = {k: float(hd[k]) for k in SCALARS}
    args = (hd["alpha_r"], hd["active_means"], hd["active_variances"], hd["weight_within_active"])
    return O.Hyper.make(*args, **kw), _capi.Hyper.make(*args, **kw)


def lung_like_hyper(R, K=2):
    """Posterior centres read off the tutorial figures."""
    return dict(s0=-2.05, s1=-0.355, tau=0.86, alpha_r=np.linspace(12.0, 19.0, R), spike_probability=0.115,
                inactive_means=-1.9, inactive_variances=1.78, active_means=np.array([2.64, 7.2, 9.0, 11.0][:K]),
                active_variances=np.full(K, 1.78), weight_active=0.725,
                weight_within_active=np.array([[1.0], [0.992, 0.008], [0.9, 0.08, 0.02], [0.7, 0.2, 0.07, 0.03]][K - 1]),
                beta=1.0, temperature=1.0, variance_g_upper_bound=float("inf"))


def make_case(G, R, seed=0, K=2, edge_cases=True, hyper=None):
    """Synthetic case with the awkward rows the reference's data contains: all-zero genes (in and out of the spike),
    partially detected genes whose 1 - p_x underflows (-Inf likelihood), tiny / huge variances, pinned-active genes."""
    X, gl, _ = synth.simulate(G, R, config_id=100 + seed)
    hd = dict(synth.truth_hyper_dict(R)) if hyper is None else dict(hyper)
    if K != 2 and hyper is None:
        hd = lung_like_hyper(R, K)
    st = synth.initial_state(X, gl, hd, seed + 7)
    rng = np.random.Generator(np.random.PCG64(seed + 99))
    if edge_cases and G >= 64:
        az = np.where(np.isneginf(X).all(1))[0]
        # some all-zero genes sit outside the spike (inactive slab) so both directions of mhSpikeAllocation are exercised
        out = az[::2]
        st["inactive_spike_allocation"][out] = 0
        st["Yg"][out] = hd["inactive_means"] + np.sqrt(hd["inactive_variances"]) * rng.standard_normal(len(out))
        # a few highly expressed genes with one dropout: 1 - p_x rounds to 0 => -Inf likelihood (lung genes 2774, 3081, ...)
        hi = np.where(~np.isneginf(X).any(1) & (st["Yg"] > 3.0))[0][:5]
        X[hi, 0] = -np.inf
        st["variance_g"][rng.integers(0, G, 8)] = 1e-3
        st["variance_g"][rng.integers(0, G, 8)] = 50.0
        st["tuningParam_yg"] = rng.uniform(0.05, 2.0, G)
        st["tuningParam_variance_g"] = rng.uniform(0.05, 2.0, G)
    return X, gl, hd, st


def oracle_state(X, gl, st, pinned=None):
    return O.State(X, gl, st["Yg"], st["variance_g"], st["alloc_active_inactive"], st["alloc_within_active"],
                   st["inactive_spike_allocation"], tuningParam_yg=st["tuningParam_yg"],
                   tuningParam_variance_g=st["tuningParam_variance_g"], pinned_active=pinned)


def gpu_handle(X, gl, hg, s, chains=1, seed=1234, gene_offset=0, pinned=None):
    """Handle whose state (including the two caches) is copied from oracle state `s`."""
    G, R = X.shape
    h = _capi.Handle(G, R, hg.num_active_components, num_chains=chains, seed=seed, gene_offset=gene_offset)
    h.upload_data(X, gl)
    for c in range(chains):
        h.set_hyper(hg, c)
        h.set_state(c, Yg=s.Yg, variance_g=s.variance_g, alloc_active_inactive=s.alloc_active_inactive,
                    alloc_within_active=s.alloc_within_active, inactive_spike_allocation=s.inactive_spike_allocation,
                    XgLikelihood=s.XgLikelihood, variance_g_probability=s.variance_g_probability,
                    tuningParam_yg=s.tuningParam_yg, tuningParam_variance_g=s.tuningParam_variance_g, pinned_active=pinned)
    return h


def setup_pair(G, R, seed=0, K=2, chains=1, pinned_frac=0.0, **kw):
    X, gl, hd, st = make_case(G, R, seed, K, **kw)
    ho, hg = hyper_pair(hd)
    pinned = None
    if pinned_frac > 0:
        pinned = (np.random.Generator(np.random.PCG64(seed + 5)).random(G) < pinned_frac).astype(np.int32)
        pinned[np.isneginf(X).all(1)] = 0
    s = oracle_state(X, gl, st, pinned)
    O.refresh_caches(ho, s)
    h = gpu_handle(X, gl, hg, s, chains=chains, pinned=pinned)
    return X, gl, hd, ho, hg, s, h


def rel_close(a, b, rtol=1e-10, name="", atol=0.0):
    """north-star check 1: |a-b| <= rtol*|b| (+ atol) elementwise; infinities and NaNs must sit at identical positions."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (name, a.shape, b.shape)
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin), f"{name}: finite masks differ at {np.where(np.isfinite(a) != fin)[0][:10]}"
    assert np.array_equal(a[~fin], b[~fin], equal_nan=True), f"{name}: non-finite values differ"
    err = np.abs(a[fin] - b[fin])
    tol = rtol * np.abs(b[fin]) + atol
    bad = err > tol
    assert not bad.any(), f"{name}: {bad.sum()} values beyond rtol={rtol}; worst rel err {np.max(err / np.maximum(np.abs(b[fin]), 1e-300)):.3e}"
    return float(np.max(err / np.maximum(np.abs(b[fin]), 1e-300))) if fin.any() else 0.0


def assert_decisions(dec_gpu, dec_ref, log_ratio=None, log_u=None, name=""):
    """north-star check 2: bit-exact decisions.  A mismatch is only tolerated (and reported) when it is a numerical
    tie, |log(u) - R| < 1e-9, which neither libm nor R's own long-double sums resolve identically."""
    dec_gpu, dec_ref = np.asarray(dec_gpu).ravel(), np.asarray(dec_ref).ravel()
    diff = np.where(dec_gpu != dec_ref)[0]
    if len(diff) == 0:
        return 0
    assert log_ratio is not None, f"{name}: {len(diff)} decisions differ, first at gene {diff[0]}"
    gap = np.abs(np.asarray(log_u).ravel()[diff] - np.asarray(log_ratio).ravel()[diff])
    assert np.all(gap < 1e-9), f"{name}: {len(diff)} decisions differ and are not ties (min gap {gap.min():.3e} max {gap.max():.3e})"
    print(f"[near-tie report] {name}: {len(diff)} tied decisions, gaps {gap}")
    return len(diff)


def compare_state(h, s, chain=0, rtol=1e-12, caches=True):
    g = h.get_state(chain)
    for k in ("alloc_active_inactive", "inactive_spike_allocation"):
        assert np.array_equal(g[k], getattr(s, k)), k
    act = s.alloc_active_inactive == 1
    assert np.array_equal(g["alloc_within_active"][act], s.alloc_within_active[act]), "alloc_within_active"
    # Yg' = Yg + tuning * z can land arbitrarily close to 0 (cancellation): a few ulps of the O(1..10) operands are allowed
    rel_close(g["Yg"], s.Yg, rtol, "Yg", atol=2e-14)
    rel_close(g["variance_g"], s.variance_g, rtol, "variance_g")
    if caches:
        rel_close(g["XgLikelihood"], s.XgLikelihood, 1e-10, "XgLikelihood")
        rel_close(g["variance_g_probability"], s.variance_g_probability, 1e-10, "variance_g_probability")
    return g
