"""The multi-rank path on a box with ONE GPU: two processes share the device, their handles exchange the statistics through CUDA-IPC
peer memory inside the persistent kernel (tests/p2p_same_device_check.py).  The exchange payload is the integer fixed-point totals, so
two shards must reproduce the single-handle chain EXACTLY (not to a tolerance), and a missing peer must be fail-stop."""
import os
import numpy as np
import pytest
import torch.multiprocessing as mp

from p2p_same_device_check import worker

pytestmark = pytest.mark.gpu


def _spawn(mode, port_base):
    out = mp.Manager().dict()
    mp.spawn(worker, args=(2, port_base + os.getpid() % 300, mode, out), nprocs=2, join=True)
    return out


@pytest.mark.timeout(600)
def test_two_shards_on_one_device_reproduce_the_single_handle_chain_exactly():
    out = _spawn("chain", 31000)
    a, b = out[0], out[1]
    assert a["flag"] == 0 and b["flag"] == 0
    assert a["vec"] == b["vec"], "the ranks hold different hyper-parameters"
    # integer all-reduce: the sharded totals are the single-handle totals bit for bit, so is everything derived from them
    assert a["vec"] == a["full_vec"], (a["vec"], a["full_vec"])
    for r in (a, b):
        assert np.array_equal(r["yg"], a["full_yg"][r["lo"]:r["hi"]]) and np.array_equal(r["ai"], a["full_ai"][r["lo"]:r["hi"]])
    assert a["vec"][-1] == 18.0


@pytest.mark.timeout(600)
def test_a_missing_peer_is_fail_stop():
    out = _spawn("timeout", 31400)
    r = out[0]
    assert r["flag"] & 4, r                       # bit 2: the exchange timed out
    assert r["all_nan"], "the statistics of a timed-out exchange must be poisoned, not stale"
    assert 1.5 < r["seconds"] < 30, r


@pytest.mark.timeout(600)
def test_posterior_predictive_statistics_on_gene_shards():
    """zz_post_predictive_sharded: the histograms of two shards, summed inside the call, give the statistics of the whole gene set.
    (init="device" + Philox keyed by the global gene index: the two-shard chain IS the one-handle chain, so the states agree too.)"""
    from p2p_same_device_check import postpred_worker
    out = mp.Manager().dict()
    mp.spawn(postpred_worker, args=(2, 31800 + os.getpid() % 300, out), nprocs=2, join=True)
    a, b = out[0], out[1]
    assert a["pp"] == b["pp"], "the ranks hold different statistics"
    assert abs(a["tau"] - a["one_tau"]) < 1e-9 * abs(a["one_tau"])
    for k, v in a["one"].items():
        assert np.allclose(a["pp"][k], v, rtol=1e-7, atol=1e-9), (k, a["pp"][k], v)
    assert np.isfinite(a["pp"]["W_L"]) and np.isfinite(a["pp"]["B"])
