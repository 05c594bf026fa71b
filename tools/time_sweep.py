import sys, os, time, numpy as np
sys.path.insert(0, ".")
import torch
from zigzag_b200 import _capi, synth
name = sys.argv[1] if len(sys.argv) > 1 else "1m_x16"
X, gl, hd, st, chains = synth.make_config(name, int(sys.argv[2]) if len(sys.argv) > 2 else None)
G, R = X.shape
hg = _capi.Hyper.make(hd["alpha_r"], hd["active_means"], hd["active_variances"], hd["weight_within_active"],
                      **{k: float(hd[k]) for k in ("s0","s1","tau","spike_probability","inactive_means","inactive_variances","weight_active","beta","temperature","variance_g_upper_bound")})
h = _capi.Handle(G, R, 2, num_chains=chains, seed=1)
h.set_stream(torch.cuda.current_stream().cuda_stream)
h.upload_data(X, gl)
for c in range(chains):
    h.set_hyper(hg, c); h.set_state(c, **st)
h.refresh_caches()
if os.environ.get("ZZ_TS_STATS"):      # also time the fused statistics epilogue (what bench.py's step runs)
    stats = torch.zeros(chains * h.nstat, dtype=torch.float64, device="cuda")
    h.set_stats_output_dev(stats.data_ptr(), allreduce=False)
for i in range(20): h.sweep(i)
torch.cuda.synchronize()
n = 200
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(n): h.sweep(100 + i)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
e0.record()
for i in range(50): h.debug_refresh_lps()
e1.record(); torch.cuda.synchronize()
print(f"   detection-loop-only kernel (k_refresh_lps): {e0.elapsed_time(e1)/50*1e3:.1f} us")
print(f"{os.environ.get('ZIGZAG_GPU_LIB','default')}: {name} sweep+spike {ms*1e3:.1f} us  -> {G*chains/ms/1e6:.2f} G gene-updates/s, {G*chains*(8*R+60)/ms/1e6/6547.2*100:.1f}% of HBM roofline")
