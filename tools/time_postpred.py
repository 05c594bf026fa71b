"""Times zz_post_predictive (one simulated data set + all discrepancy statistics) for a config.  Usage: python tools/time_postpred.py [config]"""
import sys, time, numpy as np
sys.path.insert(0, ".")
from statistics import NormalDist
from zigzag_b200 import _capi, synth
name = sys.argv[1] if len(sys.argv) > 1 else "1m_x16"
X, gl, hd, st, chains = synth.make_config(name)
G, R = X.shape
hg = _capi.Hyper.make(hd["alpha_r"], hd["active_means"], hd["active_variances"], hd["weight_within_active"],
                      **{k: float(hd[k]) for k in ("s0","s1","tau","spike_probability","inactive_means","inactive_variances","weight_active","beta","temperature","variance_g_upper_bound")})
h = _capi.Handle(G, R, 2, seed=1)
h.upload_data(X, gl); h.set_hyper(hg, 0); h.set_state(0, **st); h.refresh_caches()
yo, xf = st["Yg"][st["inactive_spike_allocation"] == 0], X[np.isfinite(X)]
r = [NormalDist(hd["inactive_means"], hd["inactive_variances"] ** 0.5).inv_cdf(1e-6),
     NormalDist(hd["active_means"][-1], hd["active_variances"][-1] ** 0.5).inv_cdf(1 - 1e-6), yo.min(), yo.max(), xf.min(), xf.max()]
bins = np.arange(min(r) - 2, max(r) + 2, 0.1)
for k in range(3): out = h.post_predictive(k, bins)
t0 = time.perf_counter()
n = 20
for k in range(n): out = h.post_predictive(10 + k, bins)
dt = (time.perf_counter() - t0) / n
print(f"{name}: zz_post_predictive {dt*1e3:.3f} ms per simulated data set ({len(bins)} bins; {G*R/dt/1e9:.2f} G simulated values/s); "
      f"W_L {out['W_L']:.4f} W_U {out['W_U']:.4f} B {out['B']:.5f}")
