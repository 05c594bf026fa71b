"""zz_get_state, both transfer paths: a pageable destination is filled through the handle's pinned bounce region by host threads, a
pinned destination receives the device-to-host copy directly.  Same bytes either way, for any thread count, for NULL fields, and
for sizes around the threading threshold (2 MiB per array) and page boundaries."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import sys, numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
import torch
from zigzag_b200 import _capi
from _util import make_case, hyper_pair
for G in (777, 262144 + 3, 300001):                       # one thread; exactly over the 2 MiB threshold for doubles; ragged
    X, gl, hd, st = make_case(G, 2, 5)
    _, hg = hyper_pair(hd)
    h = _capi.Handle(G, 2, 2, seed=3)
    h.upload_data(X, gl); h.set_hyper(hg, 0); h.set_state(0, **st); h.refresh_caches()
    h.sweep(1, tune=True); h.sweep(2)
    a = h.get_state(0)                                    # pageable, fresh arrays
    b = h.get_state(0)                                    # second call reuses the bounce region
    names = list(a)
    assert all(np.array_equal(a[k], b[k]) for k in names)
    assert np.array_equal(a["Yg"], st["Yg"]) is False     # the sweeps moved the state
    pin = {k: torch.from_numpy(np.zeros_like(a[k])).pin_memory() for k in names}     # direct path
    h._ck(h.L.zz_get_state(h.h, 0, *[pin[k].data_ptr() for k in names]))
    assert all(np.array_equal(a[k], pin[k].numpy()) for k in names), "pinned and pageable destinations differ"
    # a subset of the fields, mixed pinned / pageable
    y = np.empty(G); w = np.empty(G, np.int32)
    args = [_capi._ptr(y), None, None, _capi._ptr(w), None, pin["XgLikelihood"].data_ptr(), None, None, None]
    pin["XgLikelihood"].zero_()
    h._ck(h.L.zz_get_state(h.h, 0, *args))
    assert np.array_equal(y, a["Yg"]) and np.array_equal(w, a["alloc_within_active"]) and np.array_equal(pin["XgLikelihood"].numpy(), a["XgLikelihood"])
    # a device pointer where a host pointer belongs is handed to the copy as it is (unified addressing): no host-side memcpy into it
    dev = torch.zeros(G, dtype=torch.float64, device="cuda")
    rc = h.L.zz_get_state(h.h, 0, *([dev.data_ptr()] + [None] * 8))
    torch.cuda.synchronize()
    assert rc != 0 or np.array_equal(dev.cpu().numpy(), a["Yg"]), "device destination: neither refused nor filled"
    h.close()
print("state pull OK")
'''


@pytest.mark.timeout(600)
@pytest.mark.parametrize("threads", ["1", "4", "7"])
def test_get_state_pageable_and_pinned_destinations_agree(threads):
    env = dict(os.environ, ZZ_HOST_COPY_THREADS=threads)
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}], capture_output=True, text=True, timeout=500, env=env)
    assert r.returncode == 0 and "state pull OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
