"""Generations per second of a whole MCMC generation (fused per-gene sweep + one pass of every hyper-parameter move):
device-resident loop (zz_run_generations, the host only enqueues) against the host-driven loop, which runs the same hyper
moves in Python with one device reduction + host read-back each (zigzag_b200/zigzag.py, schedule="fused").
Usage: python tools/time_generations.py [config] [generations]"""
import sys, time, numpy as np
sys.path.insert(0, ".")
import torch
from zigzag_b200 import _capi, synth
name = sys.argv[1] if len(sys.argv) > 1 else "1m_x16"
ngen = int(sys.argv[2]) if len(sys.argv) > 2 else 200
X, gl, hd, st, chains = synth.make_config(name)
G, R = X.shape
K = len(hd["active_means"])
hg = _capi.Hyper.make(hd["alpha_r"], hd["active_means"], hd["active_variances"], hd["weight_within_active"],
                      **{k: float(hd[k]) for k in ("s0","s1","tau","spike_probability","inactive_means","inactive_variances","weight_active","beta","temperature","variance_g_upper_bound")})
thr_a = (1.0, 6.0)[:K]
h = _capi.Handle(G, R, K, num_chains=chains, seed=1)
h.upload_data(X, gl)
for c in range(chains):
    h.set_hyper(hg, c); h.set_state(c, **st)
h.refresh_caches()
h.chain_begin(_capi.Priors.defaults(-1.0, thr_a), [_capi.ChainScalars.initial(hd["active_means"], thr_a, R) for _ in range(chains)])
for tune in (True, False):
    h.run_generations(20, tune=tune); h.sync()
    l0 = h.launch_count
    t0 = time.perf_counter()
    h.run_generations(ngen, tune=tune, sample_frequency=10); h.sync()
    dt = (time.perf_counter() - t0) / ngen
    print(f"{name}: device-resident {'burn-in' if tune else 'sampling'} generation {dt*1e6:.1f} us "
          f"({(h.launch_count-l0)/ngen:.1f} launches, {chains} chain(s)) -> {1/dt:.0f} generations/s, {G*chains/dt/1e9:.2f} G gene-updates/s inside full chains")
    tl = h.debug_gen_times() / (ngen if tune else 10) / 1e3
    print(f"   last segment: {tl[0]:.1f} us per generation between the last ticket and the release, {tl[1]:.1f} us of it in the scalar stage"
          + (f" (Gibbs draws + move chains {tl[4]:.1f}, tune_all {tl[5]:.1f}, derive + publish {tl[6]:.1f}; -DZZ_GEN_PROFILE build)" if tl[4] > 0 else ""))
h.run_sweeps(10**6, 20); h.sync()
t0 = time.perf_counter(); h.run_sweeps(10**6 + 20, ngen); h.sync(); dt = (time.perf_counter() - t0) / ngen
print(f"{name}: zz_run_sweeps (fixed hyper-parameters, statistics every sweep) {dt*1e6:.1f} us per sweep -> {G*chains/dt/1e9:.2f} G gene-updates/s; releasing CTA {h.debug_gen_times()[0]/ngen/1e3:.1f} us")
sc, hy = h.chain_read()
if chains > 1: sc, hy = sc[0], hy[0]
print("   after", sc.gen, "generations: tau %.4f s0 %.4f s1 %.4f w_active %.4f spike %.4f alpha[0] %.3f" % (hy.tau, hy.s0, hy.s1, hy.weight_active, hy.spike_probability, hy.alpha_r[0]))
if "--host" in sys.argv:
    # the same generation driven from the host: zigzag_b200/zigzag.py (schedule="fused"): every hyper move is Python + a device
    # reduction whose result is read back before the move can be decided
    from zigzag_b200.zigzag import zigzag
    z = zigzag(X, gl, threshold_i=-1.0, threshold_a=thr_a, seed=2, write_settings=False, output_directory="/tmp/zz_time_out")
    for _ in range(5): z._generation(True, "fused")
    z.h.sync(); t0 = time.perf_counter()
    n = max(10, ngen // 10)
    for _ in range(n): z._generation(True, "fused")
    z.h.sync(); dt = (time.perf_counter() - t0) / n
    print(f"{name}: host-driven burn-in generation {dt*1e6:.1f} us -> {1/dt:.0f} generations/s")
