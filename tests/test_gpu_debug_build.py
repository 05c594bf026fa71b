"""Memory safety by construction, checked on the device: the GPU pool has compute-sanitizer closed, so the suite runs the
bounds-checked build (zigzag_b200/libzigzag_gpu_dbg.so: the same sources with -DZZ_BOUNDS_CHECK) once over every kernel family that
addresses memory through an index it READ from memory — the gene permutation and its inverse (upload / get / set state, host-buffer
sweep), slot lists, candidate slots of the sample sink, tile tickets of the persistent kernel, component numbers of the statistics and
the allocation trace.  A violated assert traps the kernel and fails the run; and because the two builds share every arithmetic
instruction, the results must be bit-identical to the release build's."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import json, sys, numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
from zigzag_b200 import _capi, synth
from _util import make_case, hyper_pair
X, gl, hd, st = make_case(5000, 4, 21)
_, hg = hyper_pair(hd)
h = _capi.Handle(5000, 4, 2, seed=5)
h.upload_data(X, gl); h.set_hyper(hg, 0); h.set_state(0, **st); h.refresh_caches()
out = {}
h.mh_yg(1, tune=True); h.mh_variance_g(2, tune=True); h.gibbs_alloc(3); h.gibbs_alloc_within(4); h.mh_spike_alloc(5)   # reference-order kernels
h.sweep(6, tune=True); h.sweep(7)                                                                                     # production sweep
out["stats"] = h.suffstats()[0].tolist()
thr_a = (1.0, 6.0)
h.chain_begin(_capi.Priors.defaults(-1.0, thr_a), _capi.ChainScalars.initial(hd["active_means"], thr_a, 4))
cand = np.array([0, 17, 4999, 2500], np.int32)
h.set_sample_sink(64, cand, 64, 1)
h.run_generations(30, tune=True); h.run_generations(20, tune=False, sample_frequency=4)                               # persistent kernel, dynamic tiles
prow, yrow, vrow = h.sample_rows(); h.clear_sample_sink()
out["rows"] = [float(prow.sum()), float(yrow.sum()), float(vrow.sum())]
h.run_sweeps(1000, 5)
s = h.get_state(0)
host = [np.ascontiguousarray(s[k]) for k in ("Yg", "variance_g", "alloc_active_inactive", "alloc_within_active", "inactive_spike_allocation",
                                             "XgLikelihood", "variance_g_probability", "tuningParam_yg", "tuningParam_variance_g")]
h.sweep_host(2000, False, *host)                                                                                      # pageable host buffers: chunked list mode
out["host"] = [float(host[0].sum()), float(host[5].sum()), int(host[2].sum())]
m = np.log(np.var(X[~np.isneginf(X).any(1)], axis=1, ddof=1).mean())
h.init_state(9, float(m), 0.7)
h.accumulate_alloc_trace()
out["trace"] = int(h.get_alloc_trace(0).sum())
out["state"] = [float(h.get_state(0)["Yg"].sum()), float(h.get_state(0)["variance_g"].sum())]
hb = _capi.Handle(3000, 4, 2, num_chains=3, seed=6)                                                                   # batched chains: last-arriver stage
hb.upload_data(X[:3000], gl[:3000])
for c in range(3):
    hb.set_hyper(hg, c); hb.set_state(c, **{k: v[:3000] for k, v in st.items()})
hb.refresh_caches()
hb.chain_begin(_capi.Priors.defaults(-1.0, thr_a), [_capi.ChainScalars.initial(hd["active_means"], thr_a, 4) for _ in range(3)])
hb.run_generations(25, tune=True)
hb.accumulate_alloc_trace_from(2, 0); hb.copy_chain_tuning(1, 2)
out["batched"] = [float(hb.get_state(c)["Yg"].sum()) for c in range(3)]
out["nan"] = [h.nan_flag(), hb.nan_flag()]
print("RESULT " + json.dumps(out))
'''


def _run(lib):
    env = dict(os.environ)
    if lib:
        env["ZIGZAG_GPU_LIB"] = lib
    else:
        env.pop("ZIGZAG_GPU_LIB", None)
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "device assert" not in r.stdout + r.stderr
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[7:])


@pytest.mark.timeout(900)
def test_bounds_checked_build_runs_clean_and_bit_identical():
    dbg = os.path.join(ROOT, "zigzag_b200", "libzigzag_gpu_dbg.so")
    if not os.path.exists(dbg):
        from zigzag_b200 import build
        build.build_debug()
    a, b = _run(dbg), _run(None)
    # (the case carries the awkward rows of make_case — genes at -Inf likelihood raise the NaN flag in both builds alike)
    assert a == b, (a, b)
    assert all(np.isfinite(v) for v in a["batched"] + a["state"])
