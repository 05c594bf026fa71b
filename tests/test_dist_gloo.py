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


def _product_worker(rank, world, port, G, R, outdir, out):
    """The PRODUCT's host mirror on gene shards: zigzag(shard=...), its _allreduce / _gather_genes and the sample-time file writing
    under gloo.  The device is replaced by the oracle-backed test double (no GPU in the CPU suite)."""
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from zigzag_b200 import _capi, synth, zigzag as Z
    from _oracle_handle import OracleHandle
    _capi.Handle = OracleHandle
    X, gl, _ = synth.simulate(G, R, 77)
    data = np.where(np.isneginf(X), 0.0, np.exp(X))
    z = Z.zigzag(data, gl, threshold_i=-1.0, threshold_a=(1.0, 6.0), seed=5, output_directory=outdir,
                 candidate_gene_list=[f"gene_{i}" for i in (1, G // 2, G // 2 + 1, G)], shard=(rank, world),
                 process_group=dist.group.WORLD, write_settings=(rank == 0))
    assert isinstance(z.h, OracleHandle) and z.h.G == z.g1 - z.g0
    # sampling points at i = 300, 400 (B:116: i > 2 * sample_frequency), candidate-gene rows at i = 400 ((i / sf) % 4 == 0): the
    # gather over the shards is a collective and must be entered by every rank
    z.burnin(sample_frequency=100, ngen=420, write_to_files=True, burninprefix="t", schedule="fused")
    z.mcmc(sample_frequency=10, ngen=40, write_to_files=True, mcmcprefix="t", schedule="reference")
    st = z.state()
    out[rank] = dict(hyper=z._hyper_dict(), lnl=list(z.lnl_trace), gen=z.gen, n_samples=z.n_samples,
                     yg=st["Yg"].copy(), spike=st["inactive_spike_allocation"].copy(), g=(z.g0, z.g1),
                     tune=(z.tuningParam_tau, z.inactive_mean_tuningParam), launches=z.h.launch_count)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_product_host_path_on_two_gene_shards_with_file_writing(tmp_path):
    G, R, world = 600, 4, 2
    out = mp.Manager().dict()
    port = 30100 + (os.getpid() % 500)
    mp.spawn(_product_worker, args=(world, port, G, R, str(tmp_path), out), nprocs=world, join=True)
    a, b = out[0], out[1]
    assert a["g"] == (0, G // 2) and b["g"] == (G // 2, G) and a["launches"] > 0
    for k, v in a["hyper"].items():                                       # replicated hyper-parameters stay bit-identical
        assert np.array_equal(np.asarray(v), np.asarray(b["hyper"][k])), k
    assert a["lnl"] == b["lnl"] and a["gen"] == b["gen"] and a["n_samples"] == b["n_samples"] and a["tune"] == b["tune"]
    assert a["tune"][0] != 0.05                                            # tune_all ran (generation 400)
    rows = lambda f: [l.rstrip("\n").split("\t") for l in open(os.path.join(str(tmp_path), f))]
    par = rows("t_burnin/t_model_parameters.log")
    assert [r[0] for r in par[1:]] == ["300", "400"] and len(par[0]) == 5 + R + 3 + 2 * 2 + 1 + 2
    yg = rows("t_burnin/t_yg_candidate_genes.log")
    assert yg[0] == ["gen", "gene_1", f"gene_{G // 2}", f"gene_{G // 2 + 1}", f"gene_{G}"] and [r[0] for r in yg[1:]] == ["400"]
    par = rows("t_mcmc_output/t_model_parameters.log")
    assert len(par) - 1 == a["n_samples"] - 1 and all(len(r) == len(par[0]) for r in par)
    yg = rows("t_mcmc_output/t_yg_candidate_genes.log")
    last = [float(x) for x in yg[-1][1:]]                                  # the last row holds genes of BOTH shards, from the final state
    full_y = np.concatenate([a["yg"], b["yg"]]); full_s = np.concatenate([a["spike"], b["spike"]])
    want = [(-10.0 if full_s[i] else round(float(full_y[i]), 3)) for i in (0, G // 2 - 1, G // 2, G - 1)]
    assert last == want
    pa = rows("t_mcmc_output/t_probability_active.tab")
    assert len(pa) == G + 1 and all(0.0 <= float(r[1]) <= 1.0 for r in pa[1:])


def test_shard_ranges_partition_the_genes():
    for G in (10, 3001, 1_000_000):
        for world in (1, 2, 4, 8):
            edges = [(r * G // world, (r + 1) * G // world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == G
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            assert max(e[1] - e[0] for e in edges) - min(e[1] - e[0] for e in edges) <= 1
