import sys, numpy as np
sys.path.insert(0, ".")
import torch
from zigzag_b200 import _capi, synth
X, gl, hd, st, _ = synth.make_config("1m_x16")
R = X.shape[1]; G = X.shape[0]
hg = _capi.Hyper.make(hd["alpha_r"], hd["active_means"], hd["active_variances"], hd["weight_within_active"],
                      **{k: float(hd[k]) for k in ("s0","s1","tau","spike_probability","inactive_means","inactive_variances","weight_active","beta","temperature","variance_g_upper_bound")})
h = _capi.Handle(G, R, 2, seed=1)
h.set_stream(torch.cuda.current_stream().cuda_stream)
h.upload_data(X, gl); h.set_hyper(hg, 0); h.set_state(0, **st); h.refresh_caches()
h.run_sweeps(0, 50); h.sync()
for rep in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h.run_sweeps(100, 2); e0.record(); h.run_sweeps(200 + 1000 * rep, 200); e1.record(); torch.cuda.synchronize()
    print(f"{e0.elapsed_time(e1) * 1e3 / 200:.2f}", end=" ")
print()
