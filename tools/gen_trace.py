"""Per-CTA timeline of the persistent kernel: where the time of a sweep / generation goes.  Needs a build with the stamps compiled in
(they cost registers in the tile loop, so the product build has none):
    ZZ_BUILD_OUT=$PWD/scratch/lib_trace.so ZZ_NVCC_EXTRA=-DZZ_GEN_TRACE_BUILD python zigzag_b200/build.py --force
    ZIGZAG_GPU_LIB=$PWD/scratch/lib_trace.so python tools/gen_trace.py [workload] [genes]   (prints a summary per launch kind)"""
import os, sys, numpy as np
sys.path.insert(0, ".")
path = os.environ.setdefault("ZZ_GEN_TRACE", "/tmp/zz_gen_trace.bin")
if os.path.exists(path): os.remove(path)
import torch
from zigzag_b200 import _capi, synth
name = sys.argv[1] if len(sys.argv) > 1 else "1m_x16"
X, gl, hd, st, chains = synth.make_config(name)
G = int(sys.argv[2]) if len(sys.argv) > 2 else X.shape[0]
X, gl, st = np.asfortranarray(X[:G]), gl[:G], {k: v[:G] for k, v in st.items()}
R = X.shape[1]
hg = _capi.Hyper.make(hd["alpha_r"], hd["active_means"], hd["active_variances"], hd["weight_within_active"],
                      **{k: float(hd[k]) for k in ("s0","s1","tau","spike_probability","inactive_means","inactive_variances","weight_active","beta","temperature","variance_g_upper_bound")})
h = _capi.Handle(G, R, 2, seed=1)
h.upload_data(X, gl); h.set_hyper(hg, 0); h.set_state(0, **st); h.refresh_caches()
thr_a = (1.0, 6.0)
h.chain_begin(_capi.Priors.defaults(-1.0, thr_a), _capi.ChainScalars.initial(hd["active_means"], thr_a, R))
h.run_sweeps(0, 16); h.run_sweeps(100, 16); h.run_generations(16, tune=True); h.run_generations(16, tune=True); h.sync()
raw = np.fromfile(path, dtype=np.uint64)
pos, k = 0, 0
kinds = ["sweeps (warm-up)", "sweeps", "generations (warm-up)", "generations"]
while pos < len(raw):
    assert raw[pos] == 0x7a7a747261636531
    ng, nc, stw = int(raw[pos + 1]), int(raw[pos + 2]), int(raw[pos + 3])
    t = raw[pos + 4: pos + 4 + ng * nc * 4].astype(np.int64).reshape(ng, nc, 4); pos += 4 + ng * nc * 4
    w = t[:, :nc - stw]                                  # workers
    s = t[:, nc - 1] if stw else None
    print(f"== {kinds[k] if k < len(kinds) else k}: {ng} generations, {nc - stw} worker CTAs" + (" + steward" if stw else ""))
    for gi in range(2, ng):
        rel = s[gi - 1][2] if stw else w[gi][:, 0].min()   # release of the previous generation = start of this one
        f = lambda a: f"min {a.min()/1e3:6.2f} p50 {np.median(a)/1e3:6.2f} p95 {np.percentile(a,95)/1e3:6.2f} max {a.max()/1e3:6.2f}"
        line = (f"gen {gi:2d}: release seen +[{f(w[gi][:,0]-rel)}]  tiles start +[{f(w[gi][:,1]-rel)}]  tiles done +[{f(w[gi][:,2]-rel)}]  ticket +[{f(w[gi][:,3]-rel)}]")
        if stw: line += f"  steward: last ticket seen +{(s[gi][0]-rel)/1e3:6.2f}, totals +{(s[gi][1]-rel)/1e3:6.2f}, release +{(s[gi][2]-rel)/1e3:6.2f}"
        if gi in (2, 3, ng - 1): print(line)
    if stw:
        per = np.diff(s[2:, 2]) / 1e3
        busy = (w[3:, :, 2] - w[3:, :, 1]).mean() / 1e3
        print(f"   period {per.mean():.2f} us; mean time a CTA spends in tiles {busy:.2f} us; mean wait for the slowest CTA {((s[3:,0][:,None] - w[3:,:,3]).mean())/1e3:.2f} us; "
              f"last ticket -> release {((s[3:,2]-s[3:,0]).mean())/1e3:.2f} us; release -> tiles start {((w[3:,:,1] - s[2:-1,2][:,None]).mean())/1e3:.2f} us")
    k += 1
