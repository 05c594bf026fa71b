"""Fixed and per-gene cost of the persistent sweep: zz_run_sweeps at several shard sizes of the 1M x 16 workload (time per sweep =
a + b * genes).  Usage: python tools/scan_sweeps.py [nsweeps]"""
import sys, time, numpy as np
sys.path.insert(0, ".")
import torch
from zigzag_b200 import _capi, synth
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 200
X, gl, hd, st, _ = synth.make_config("1m_x16")
R = X.shape[1]
hg = _capi.Hyper.make(hd["alpha_r"], hd["active_means"], hd["active_variances"], hd["weight_within_active"],
                      **{k: float(hd[k]) for k in ("s0","s1","tau","spike_probability","inactive_means","inactive_variances","weight_active","beta","temperature","variance_g_upper_bound")})
rows = []
for G in (15625, 31250, 62500, 125000, 250000, 500000, 1000000):
    h = _capi.Handle(G, R, 2, seed=1)
    h.set_stream(torch.cuda.current_stream().cuda_stream)
    h.upload_data(np.asfortranarray(X[:G]), gl[:G]); h.set_hyper(hg, 0); h.set_state(0, **{k: v[:G] for k, v in st.items()}); h.refresh_caches()
    h.run_sweeps(0, 20); h.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h.run_sweeps(100, 2); e0.record(); h.run_sweeps(200, ns); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / ns
    lead = h.debug_gen_times()[0] / ns / 1e3
    rows.append((G, us))
    print(f"G = {G:8d}: {us:7.2f} us per sweep, {G/us/1e3:6.2f} G gene-updates/s, tiles per warp {G/32/3552:5.2f}, releasing CTA {lead:.2f} us")
    h.close()
(g1, t1), (g2, t2) = rows[-2], rows[-1]
b = (t2 - t1) / (g2 - g1)
print(f"slope between the two largest sizes: {b*1e6:.1f} us per 1M genes ({1/b/1e3:.2f} G gene-updates/s asymptotically), intercept {t2 - b*g2:.1f} us")
