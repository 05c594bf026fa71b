"""One short device-resident launch for ncu: `ncu --set full -k regex:k_generations -c 2 python tools/prof_generations.py [sweeps|generations] [n]`."""
import sys
sys.path.insert(0, ".")
from zigzag_b200 import _capi, synth
mode = sys.argv[1] if len(sys.argv) > 1 else "sweeps"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 5
name = sys.argv[3] if len(sys.argv) > 3 else "1m_x16"
X, gl, hd, st, chains = synth.make_config(name)
G, R = X.shape
K = len(hd["active_means"])
hg = _capi.Hyper.make(hd["alpha_r"], hd["active_means"], hd["active_variances"], hd["weight_within_active"],
                      **{k: float(hd[k]) for k in ("s0", "s1", "tau", "spike_probability", "inactive_means", "inactive_variances", "weight_active", "beta", "temperature", "variance_g_upper_bound")})
h = _capi.Handle(G, R, K, num_chains=chains, seed=1)
h.upload_data(X, gl)
for c in range(chains):
    h.set_hyper(hg, c); h.set_state(c, **st)
h.refresh_caches()
thr_a = (1.0, 6.0)[:K]
h.chain_begin(_capi.Priors.defaults(-1.0, thr_a), [_capi.ChainScalars.initial(hd["active_means"], thr_a, R) for _ in range(chains)])
for rep in range(2):
    if mode == "sweeps":
        h.run_sweeps(1000 + 100 * rep, n)
    else:
        h.run_generations(n, tune=False, sample_frequency=10 ** 6)
    h.sync()
print("done", mode, n)
