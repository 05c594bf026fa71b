"""Times zz_sweep_host (the host-buffer round trip bench.py reports as `e2e`) for one config; ZZ_HOST_CHUNKS overrides the
number of pipeline chunks.  Usage: python tools/time_host.py [config] [steps]"""
import sys, os, time, numpy as np
sys.path.insert(0, ".")
import torch
from zigzag_b200 import _capi, synth
name = sys.argv[1] if len(sys.argv) > 1 else "1m_x16"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
X, gl, hd, st, chains = synth.make_config(name)
G, R = X.shape
hg = _capi.Hyper.make(hd["alpha_r"], hd["active_means"], hd["active_variances"], hd["weight_within_active"],
                      **{k: float(hd[k]) for k in ("s0","s1","tau","spike_probability","inactive_means","inactive_variances","weight_active","beta","temperature","variance_g_upper_bound")})
h = _capi.Handle(G, R, 2, num_chains=1, seed=1)
h.upload_data(X, gl)
h.set_hyper(hg, 0); h.set_state(0, **st)
h.refresh_caches()
f64 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).pin_memory()
i32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).pin_memory()
cur = h.get_state()
host = [f64(cur["Yg"]), f64(cur["variance_g"]), i32(cur["alloc_active_inactive"]), i32(cur["alloc_within_active"]),
        i32(cur["inactive_spike_allocation"]), f64(cur["XgLikelihood"]), f64(cur["variance_g_probability"]),
        f64(cur["tuningParam_yg"]), f64(cur["tuningParam_variance_g"])]
ptrs = [x.data_ptr() for x in host]
for i in range(3): nb = h.sweep_host(10_000 + i, False, *ptrs)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(steps): nb = h.sweep_host(20_000 + i, False, *ptrs)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / steps
print(f"chunks={os.environ.get('ZZ_HOST_CHUNKS','auto')}: {name} sweep_host {dt*1e3:.3f} ms/step -> {G/dt/1e9:.3f} G gene-updates/s; "
      f"{(nb[0]+nb[1])/dt/1e9:.1f} GB/s PCIe (both directions summed)")
