import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from oracle import oracle as O
from zigzag_b200 import _capi
from _util import *
G, R, C = 1200, 4, 5
X, gl, hd, ho, hg, s, h0 = setup_pair(G, R, 13)
h0.close()
h = gpu_handle(X, gl, hg, s, chains=C, seed=77)
hy = [ho.copy() for _ in range(C)]
ss = [s.copy() for _ in range(C)]
for c in range(C):
    hy[c].tau = ho.tau * (1 + 0.1 * c); hy[c].inactive_means = ho.inactive_means - 0.2 * c
    g = h.get_hyper(c); g.tau = hy[c].tau; g.inactive_means = hy[c].inactive_means; h.set_hyper(g, c)
    O.refresh_caches(hy[c], ss[c])
h.refresh_caches()
for c in range(C):
    try:
        compare_state(h, ss[c], chain=c); print("init state ok", c)
    except AssertionError as e: print("init state BAD", c, str(e)[:200])
for step in range(2):
    d = np.concatenate([h.philox_draws(step, _capi.SLOT_YG), h.philox_draws(step, _capi.SLOT_VARIANCE_G),
                        h.philox_draws(step, _capi.SLOT_ALLOC), h.philox_draws(step, _capi.SLOT_SPIKE)])
    dec = h.sweep(step, decisions=True)
    for c in range(C):
        dec_o = O.sweep(hy[c], ss[c], d[:, c, :])
        for j in range(4):
            bad = np.where(dec[j, c] != dec_o[j])[0]
            print("step", step, "chain", c, "move", j, "nbad", len(bad), bad[:8])
