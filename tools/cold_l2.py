"""Sweep time with a cold L2: a 512 MB buffer is rewritten before every timed launch (the state the production sweep touches at
1 M x 16 is 78 MB — Xg enters only through per-gene sufficient statistics — so consecutive sweeps of a chain find it in the
126 MB L2; this measures what a sweep costs when it does not)."""
import sys, numpy as np
sys.path.insert(0, ".")
import torch
from zigzag_b200 import _capi, synth
name = sys.argv[1] if len(sys.argv) > 1 else "1m_x16"
X, gl, hd, st, chains = synth.make_config(name)
G, R = X.shape
hg = _capi.Hyper.make(hd["alpha_r"], hd["active_means"], hd["active_variances"], hd["weight_within_active"],
                      **{k: float(hd[k]) for k in ("s0","s1","tau","spike_probability","inactive_means","inactive_variances","weight_active","beta","temperature","variance_g_upper_bound")})
h = _capi.Handle(G, R, 2, num_chains=chains, seed=1)
h.set_stream(torch.cuda.current_stream().cuda_stream)
h.upload_data(X, gl)
for c in range(chains):
    h.set_hyper(hg, c); h.set_state(c, **st)
h.refresh_caches()
stats = torch.zeros(chains * h.nstat, dtype=torch.float64, device="cuda")
h.set_stats_output_dev(stats.data_ptr(), allreduce=False)
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for mode in ("launch (k_sweep_fast)", "persistent (k_generations, 1 sweep per launch)"):
    fn = (lambda s: h.sweep(s)) if mode.startswith("launch") else (lambda s: h.run_sweeps(s, 1))
    for cold in (False, True):
        for i in range(10): fn(i)
        torch.cuda.synchronize()
        ts = []
        for i in range(40):
            if cold: flush.fill_(i & 255)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(100 + i); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        print(f"{name} {mode}, {'L2 flushed before every launch' if cold else 'warm'}: median {np.median(ts):.1f} us, min {np.min(ts):.1f} us")
