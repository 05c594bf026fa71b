"""N > 1 host logic on CPU (gloo, world_size 2): genes shard by contiguous blocks, every reduction that feeds a hyper-parameter
move is all-reduced, and the replicated hyper-parameter RNG keeps all ranks on identical decisions.  The
per-shard numbers come from the oracle here (the checker standing in for the device), the plumbing is the product's."""
import os
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as O
from _util import make_case, hyper_pair, oracle_state


def _worker(rank, world, port, G, R, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    X, gl, hd, st = make_case(G, R, 3, edge_cases=False)
    ho, _ = hyper_pair(hd)
    g0, g1 = rank * G // world, (rank + 1) * G // world
    sub = {k: v[g0:g1] for k, v in st.items()}
    s = oracle_state(np.asfortranarray(X[g0:g1]), gl[g0:g1], sub)
    O.refresh_caches(ho, s)
    draws = O.sweep_draws(11, 0, 0, g0, g1 - g0)            # Philox keyed by the GLOBAL gene index
    dec = O.sweep(ho, s, draws)
    t = torch.from_numpy(O.suffstats(ho, s))
    dist.all_reduce(t)
    # replicated hyper move: same seed on every rank + all-reduced sums => identical accept decisions
    rng = np.random.Generator(np.random.PCG64(123))
    taup = ho.tau * math_exp(0.05 * (rng.random() - 0.5))
    lp = torch.tensor([O.varg_hyper_sum(ho, s, ho.s0, ho.s1, taup), O.varg_hyper_sum(ho, s, ho.s0, ho.s1, ho.tau)])
    dist.all_reduce(lp)
    accept = bool(np.log(rng.random()) < float(lp[0] - lp[1]))
    out[rank] = (t.numpy().copy(), dec, float(lp[0]), accept, s.Yg.copy())
    dist.destroy_process_group()


def math_exp(x):
    import math
    return math.exp(x)


@pytest.mark.timeout(180)
def test_two_rank_shards_reduce_to_the_single_rank_result():
    G, R, world = 3001, 4, 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 500)
    mp.spawn(_worker, args=(world, port, G, R, out), nprocs=world, join=True)
    X, gl, hd, st = make_case(G, R, 3, edge_cases=False)
    ho, _ = hyper_pair(hd)
    s = oracle_state(X, gl, st)
    O.refresh_caches(ho, s)
    dec = O.sweep(ho, s, O.sweep_draws(11, 0, 0, 0, G))
    full = O.suffstats(ho, s)
    a, b = out[0], out[1]
    assert np.array_equal(a[0], b[0])                                   # all ranks hold the identical reduced vector
    K1 = 3
    assert np.array_equal(a[0][:K1], full[:K1]) and a[0][:K1].sum() == G  # counts are exact
    assert np.allclose(a[0], full, rtol=1e-12, atol=1e-9)
    assert np.array_equal(np.concatenate([a[1], b[1]], axis=1), dec)    # sharding does not change any decision
    assert np.array_equal(np.concatenate([a[4], b[4]]), s.Yg)
    assert a[2] == b[2] and a[3] == b[3]                                # replicated hyper decision is identical


def test_shard_ranges_partition_the_genes():
    for G in (10, 3001, 1_000_000):
        for world in (1, 2, 4, 8):
            edges = [(r * G // world, (r + 1) * G // world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == G
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            assert max(e[1] - e[0] for e in edges) - min(e[1] - e[0] for e in edges) <= 1
