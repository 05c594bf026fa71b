"""Two processes, ONE GPU: the peer-memory all-reduce of the persistent kernel between two gene shards whose handles live on the same
device (CUDA IPC works between processes on one GPU; the two cooperative kernels time-slice).  Lets a single-GPU box exercise the
multi-rank path: run by tests/test_gpu_p2p_one_device.py under torch.multiprocessing with a gloo group for the handshake.

  mode "chain":   both ranks run device-resident generations on their shard; rank 0 also runs the whole gene set on one handle —
                  hyper-parameters and per-gene state must agree, the ranks' hyper-parameters must be bit-identical.
  mode "timeout": rank 1 connects but never launches; rank 0's exchange must give up after ZZ_COMM_TIMEOUT_S, poison its statistics
                  with NaN and raise flag bit 2 (fail-stop) instead of hanging or consuming stale words."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def worker(rank, world, port, mode, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    if mode == "timeout":
        os.environ["ZZ_COMM_TIMEOUT_S"] = "2"
    import torch
    import torch.distributed as dist
    from zigzag_b200 import _capi, synth
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    G, R, K = 8000, 4, 2
    X, gl, _ = synth.simulate(G, R, config_id=777)
    hd = synth.truth_hyper_dict(R)
    st = synth.initial_state(X, gl, hd, 99)
    hg = _capi.Hyper.make(hd["alpha_r"], hd["active_means"], hd["active_variances"], hd["weight_within_active"],
                          **{k: float(hd[k]) for k in ("s0", "s1", "tau", "spike_probability", "inactive_means", "inactive_variances",
                                                       "weight_active", "beta", "temperature", "variance_g_upper_bound")})
    thr_a = (1.0, 6.0)
    pri, sc0 = _capi.Priors.defaults(-1.0, thr_a), _capi.ChainScalars.initial(hd["active_means"], thr_a, R)

    def handle(lo, hi):
        h = _capi.Handle(hi - lo, R, K, device=0, seed=31, gene_offset=lo)
        h.upload_data(np.asfortranarray(X[lo:hi]), gl[lo:hi]); h.set_hyper(hg, 0)
        h.set_state(0, **{k: v[lo:hi] for k, v in st.items()}); h.refresh_caches()
        return h

    lo, hi = rank * G // world, (rank + 1) * G // world
    h = handle(lo, hi)
    blobs = [None] * world
    dist.all_gather_object(blobs, h.comm_init(rank, world))
    h.comm_connect(blobs)
    dist.barrier()
    if mode == "timeout":
        if rank == 0:
            stats = torch.zeros(h.nstat, dtype=torch.float64, device="cuda")
            h.set_stats_output_dev(stats.data_ptr(), allreduce=True)
            t0 = time.time()
            h.run_sweeps(1, 3); h.sync()
            out[0] = dict(seconds=time.time() - t0, flag=h.nan_flag(), all_nan=bool(torch.isnan(stats).all().item()))
        dist.barrier()
        dist.destroy_process_group()
        return
    h.chain_begin(pri, sc0)
    h.run_generations(12, tune=True)
    h.run_generations(6, tune=False, sample_frequency=2)
    sc, hy = h.chain_read()
    mine = h.get_state(0)
    vec = [hy.s0, hy.s1, hy.tau, hy.spike_probability, hy.inactive_means, hy.inactive_variances, hy.weight_active, *hy.alpha_r[:R],
           *hy.active_means[:K], *hy.weight_within_active[:K], float(sc.hyper_ctr), float(sc.gen)]
    res = dict(vec=vec, flag=h.nan_flag(), yg=mine["Yg"], ai=mine["alloc_active_inactive"], lo=lo, hi=hi)
    if rank == 0:
        f = handle(0, G)
        f.chain_begin(pri, sc0)
        f.run_generations(12, tune=True)
        f.run_generations(6, tune=False, sample_frequency=2)
        scf, hyf = f.chain_read()
        full = f.get_state(0)
        res.update(full_vec=[hyf.s0, hyf.s1, hyf.tau, hyf.spike_probability, hyf.inactive_means, hyf.inactive_variances, hyf.weight_active,
                             *hyf.alpha_r[:R], *hyf.active_means[:K], *hyf.weight_within_active[:K], float(scf.hyper_ctr), float(scf.gen)],
                   full_yg=full["Yg"], full_ai=full["alloc_active_inactive"])
    out[rank] = res
    dist.barrier()
    dist.destroy_process_group()


def postpred_worker(rank, world, port, out):
    """Posterior-predictive statistics on two gene shards (zz_post_predictive_sharded through the host mirror's all-reduce, gloo) against
    the same call on one handle holding every gene."""
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    import torch
    import torch.distributed as dist
    from zigzag_b200 import synth
    from zigzag_b200.zigzag import zigzag
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    G, R = 6000, 4
    X, gl, _ = synth.simulate(G, R, config_id=778)
    data = np.where(np.isneginf(X), 0.0, np.exp(X))
    kw = dict(threshold_i=-1.0, threshold_a=(1.0, 6.0), seed=17, write_settings=False, output_directory="/tmp/zz_pp_check", init="device")
    mm = zigzag(data, gl, shard=(rank, world), process_group=dist.group.WORLD, **kw)
    mm.burnin(sample_frequency=50, ngen=60, write_to_files=False, schedule="fused")
    pp = mm.posteriorPredictiveSimulation()
    res = dict(pp={k: np.asarray(v).tolist() for k, v in pp.items()}, tau=mm.tau)
    if rank == 0:
        one = zigzag(data, gl, **kw)
        one.burnin(sample_frequency=50, ngen=60, write_to_files=False, schedule="fused")
        res["one"] = {k: np.asarray(v).tolist() for k, v in one.posteriorPredictiveSimulation().items()}
        res["one_tau"] = one.tau
    out[rank] = res
    dist.barrier()
    dist.destroy_process_group()
