"""Second, independent restatement of the reference's per-gene arithmetic in vectorised NumPy — TEST INFRASTRUCTURE ONLY.

Written to mirror the *vectorised* structure of the R code (whole-vector temporaries, a loop over libraries, index subsets),
whereas oracle/zz_oracle.c is its per-gene scalar reading; tests/test_oracle_kat.py requires the two to agree to 1e-13, which
guards each against transcription slips in the other.  Citations as in zz_oracle.c (P = R/zigzagProbability.R, ...).
"""
import numpy as np

SQRT2PI = np.sqrt(2 * np.pi)          # I:126
M_LN_SQRT_2PI = 0.918938533204672741780329736406


def dnorm_log(x, mu, sd):             # nmath dnorm, log = TRUE
    z = (x - mu) / sd
    return -(M_LN_SQRT_2PI + 0.5 * z * z + np.log(sd))


def get_sigmax(s0, s1, yy):           # U:144-148
    return np.exp(s0 + s1 * yy)


def get_px(alpha_r, yy, gl):          # U:150-159
    return 1 - np.exp(-np.outer(gl * np.exp(yy), alpha_r))


def computeXgLikelihood(xg, yg, sx, px, spike_allocation, nd_spike, beta=1.0):       # P:57-127
    n, R = xg.shape
    lnl = np.full(n, np.nan)
    out = np.where(spike_allocation == 0)[0]
    ins = np.where(spike_allocation == 1)[0]
    if len(out) > 0:
        cols = []
        for lib in range(R):
            x = xg[out, lib]
            ll = np.empty(len(out))
            nd = np.where(x == -np.inf)[0]
            det = np.where(x != -np.inf)[0]
            with np.errstate(divide="ignore"):
                ll[nd] = np.log(1 - px[out, lib][nd])                                    # P:82
                ll[det] = np.log(px[out, lib][det]) + dnorm_log(x[det], yg[out][det], np.sqrt(sx[out][det]))   # P:84-85
            cols.append(ll)
        acc = np.zeros(len(out), dtype=np.longdouble)                                    # rowSums accumulates in long double
        for c in cols:
            acc += c
        lnl[out] = acc.astype(np.float64)
    with np.errstate(divide="ignore"):
        lnl[ins] = np.log((spike_allocation[ins] * nd_spike[ins]).astype(np.float64))   # P:117
    if beta < 0.0000001:
        lnl[:] = 0
    return beta * lnl


def computeVarianceGPriorProbability(x, sx, gtau):                                       # P:11-31 (variance_g_upper_bound = Inf)
    rt = np.sqrt(gtau)
    q = (np.log(x) - np.log(sx) - gtau) / rt
    return -np.log(x * rt * SQRT2PI) - 0.5 * (q * q)


def computeInactiveLikelihood(y, spike_p, im, iv, spike_allocation):                     # P:210-238
    is_ = np.sqrt(iv)
    lnl = np.empty(len(y))
    ins, out = spike_allocation == 1, spike_allocation == 0
    lnl[ins] = np.log(spike_p)
    lnl[out] = np.log(1 - spike_p) - np.log(SQRT2PI * is_) - (y[out] - im) ** 2 / (2 * iv)
    return lnl


def computeActiveLikelihood(y, comp, am, av):                                            # P:240-273
    lnl = np.full(len(y), np.nan)
    for k in range(1, len(am) + 1):
        idx = comp == k
        lnl[idx] = -np.log(SQRT2PI * np.sqrt(av[k - 1])) - (y[idx] - am[k - 1]) ** 2 / (2 * av[k - 1])
    return lnl


def level2(h, y, ai, within, spike):
    out = np.empty(len(y))
    ina, act = ai == 0, ai == 1
    out[ina] = computeInactiveLikelihood(y[ina], h["spike_probability"], h["inactive_means"], h["inactive_variances"], spike[ina])
    out[act] = computeActiveLikelihood(y[act], within[act], h["active_means"], h["active_variances"])
    return out


def mhYg(h, X, gl, st, z, u):                                                           # Pr:250-333; st is a dict of arrays, updated in place
    out = np.where(st["inactive_spike_allocation"] == 0)[0]
    Yg, vg = st["Yg"], st["variance_g"]
    yp = Yg[out] + st["tuningParam_yg"][out] * z[out]
    Sgp = get_sigmax(h["s0"], h["s1"], yp)
    pxp = get_px(h["alpha_r"], yp, gl[out])
    zeros = np.zeros(len(out), dtype=np.int64)
    LXp = computeXgLikelihood(X[out], yp, vg[out], pxp, zeros, st["no_detect_spike"][out], h["beta"])
    PVp = computeVarianceGPriorProbability(vg[out], Sgp, h["tau"])
    ai, wi = st["alloc_active_inactive"][out], st["alloc_within_active"][out]
    Lp = LXp + level2(h, yp, ai, wi, zeros)
    Lp = Lp + PVp
    Lc = st["XgLikelihood"][out] + level2(h, Yg[out], ai, wi, zeros)
    Lc = Lc + st["variance_g_probability"][out]
    with np.errstate(invalid="ignore"):
        R = h["temperature"] * (Lp - Lc)
        acc = np.log(u[out]) < R
    Yg[out] = np.where(acc, yp, Yg[out])
    st["XgLikelihood"][out] = np.where(acc, LXp, st["XgLikelihood"][out])
    st["variance_g_probability"][out] = np.where(acc, PVp, st["variance_g_probability"][out])
    dec = np.zeros(len(Yg), np.uint8); dec[out] = acc
    return dec


def mhVariance_g(h, X, gl, st, u1, u2):                                                  # Pr:7-56
    out = np.where(st["inactive_spike_allocation"] == 0)[0]
    Yg, vg = st["Yg"], st["variance_g"]
    sf = np.exp(st["tuningParam_variance_g"][out] * (u1[out] - 0.5))
    vp = vg[out] * sf
    HR = np.log(sf)
    px = get_px(h["alpha_r"], Yg[out], gl[out])
    zeros = np.zeros(len(out), dtype=np.int64)
    LXp = computeXgLikelihood(X[out], Yg[out], vp, px, zeros, st["no_detect_spike"][out], h["beta"])
    PVp = computeVarianceGPriorProbability(vp, get_sigmax(h["s0"], h["s1"], Yg[out]), h["tau"])
    Lp, Lc = LXp + PVp, st["XgLikelihood"][out] + st["variance_g_probability"][out]
    with np.errstate(invalid="ignore"):
        R = h["temperature"] * (Lp - Lc) + HR
        R[vp > h["variance_g_upper_bound"]] = -np.inf
        acc = np.log(u2[out]) < R
    vg[out] = np.where(acc, vp, vg[out])
    st["XgLikelihood"][out] = np.where(acc, LXp, st["XgLikelihood"][out])
    st["variance_g_probability"][out] = np.where(acc, PVp, st["variance_g_probability"][out])
    dec = np.zeros(len(Yg), np.uint8); dec[out] = acc
    return dec


def gibbsAllocationActiveInactive(h, st, u_ai, u_w, pinned=None):                        # Pr:335-410
    Yg, sa = st["Yg"], st["inactive_spike_allocation"]
    K = len(h["active_means"])
    lik0 = np.where(sa == 1, np.log(h["spike_probability"]),
                    np.log(1 - h["spike_probability"]) - np.log(SQRT2PI * np.sqrt(h["inactive_variances"])) - (Yg - h["inactive_means"]) ** 2 / (2 * h["inactive_variances"]))
    L0 = np.exp(lik0)
    L1 = np.zeros(len(Yg))
    for k in range(1, K + 1):
        L1 = L1 + h["weight_within_active"][k - 1] * np.exp(computeActiveLikelihood(Yg, np.full(len(Yg), k), h["active_means"], h["active_variances"]))
    wa = h["weight_active"]
    with np.errstate(invalid="ignore", divide="ignore"):
        pa = wa * L1 * (1 - sa) / ((1 - wa) * L0 + wa * L1 * (1 - sa))
    pa[np.isnan(pa)] = wa
    a = (u_ai < pa).astype(np.int32)
    if pinned is not None:
        a[pinned == 1] = 1
    st["alloc_active_inactive"][:] = a
    act = np.where(a == 1)[0]
    if len(act):
        logL = np.stack([np.log(h["weight_within_active"][k - 1]) + computeActiveLikelihood(Yg[act], np.full(len(act), k), h["active_means"], h["active_variances"]) - 0
                         for k in range(1, K + 1)], axis=1)
        Lk = np.exp(logL)
        with np.errstate(invalid="ignore", divide="ignore"):
            probs = Lk / Lk.sum(1, dtype=np.longdouble).astype(np.float64)[:, None]
            cum = np.cumsum(probs.astype(np.longdouble), axis=1).astype(np.float64)
            rs = np.floor(cum / u_w[act][:, None])
        new = (rs == 0).sum(1) + 1
        bad = np.isnan(rs).any(1)
        new = np.minimum(new, K)
        st["alloc_within_active"][act[~bad]] = new[~bad]
    return np.where(a == 1, st["alloc_within_active"], 0).astype(np.uint8), pa
