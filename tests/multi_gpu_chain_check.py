"""Multi-GPU check of the device-resident generation loop (run by tests/test_multi_gpu.py under torchrun when the box has >= 2 GPUs).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/multi_gpu_chain_check.py

Every rank owns a contiguous gene shard and runs zz_run_generations; the sweep statistics and the gene sums of the hyper moves
are all-reduced over peer memory inside the kernels, the scalar moves are replicated.  Rank 0 also runs the whole gene set on one
GPU: hyper-parameters, stream positions and the per-gene state of every shard must agree with it, and all ranks must hold
bit-identical hyper-parameters."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from zigzag_b200 import _capi, synth


def make_handle(X, gl, hg, st, lo, hi, dev, offset):
    sl = slice(lo, hi)
    h = _capi.Handle(hi - lo, X.shape[1], 2, device=dev, seed=77, gene_offset=offset)
    h.upload_data(np.asfortranarray(X[sl]), gl[sl])
    h.set_hyper(hg, 0)
    h.set_state(0, **{k: v[sl] for k, v in st.items()})
    h.refresh_caches()
    return h


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    G, R, K = 16000, 4, 2
    X, gl, _ = synth.simulate(G, R, config_id=901)
    hd = synth.truth_hyper_dict(R)
    st = synth.initial_state(X, gl, hd, 4242)
    hg = _capi.Hyper.make(hd["alpha_r"], hd["active_means"], hd["active_variances"], hd["weight_within_active"],
                          **{k: float(hd[k]) for k in ("s0", "s1", "tau", "spike_probability", "inactive_means", "inactive_variances",
                                                       "weight_active", "beta", "temperature", "variance_g_upper_bound")})
    thr_a = (1.0, 6.0)
    pri, sc0 = _capi.Priors.defaults(-1.0, thr_a), _capi.ChainScalars.initial(hd["active_means"], thr_a, R)
    lo, hi = rank * G // world, (rank + 1) * G // world
    h = make_handle(X, gl, hg, st, lo, hi, local, lo)
    blobs = [None] * world
    dist.all_gather_object(blobs, h.comm_init(rank, world))
    h.comm_connect(blobs)
    dist.barrier()
    for mode in ("closed", "gene_sums"):
        if mode == "gene_sums":
            os.environ["ZZ_CHAIN_GENE_SUMS"] = "1"
        h.chain_begin(pri, sc0) if mode == "closed" else None
        h.run_generations(40, tune=True)
        h.run_generations(20, tune=False, sample_frequency=4)
    sc, hy = h.chain_read()
    mine = h.get_state(0)
    vec = torch.tensor([hy.s0, hy.s1, hy.tau, hy.spike_probability, hy.inactive_means, hy.inactive_variances, hy.weight_active,
                        *hy.alpha_r[:R], *hy.active_means[:K], *hy.weight_within_active[:K], float(sc.hyper_ctr)], dtype=torch.float64, device="cuda")
    allv = [torch.zeros_like(vec) for _ in range(world)]
    dist.all_gather(allv, vec)
    assert all(torch.equal(allv[0], v) for v in allv), "ranks hold different hyper-parameters"
    assert h.nan_flag() == 0
    ok = torch.ones(1, device="cuda")
    if rank == 0:
        os.environ.pop("ZZ_CHAIN_GENE_SUMS", None)
        f = make_handle(X, gl, hg, st, 0, G, local, 0)
        f.chain_begin(pri, sc0)
        for mode in ("closed", "gene_sums"):
            if mode == "gene_sums":
                os.environ["ZZ_CHAIN_GENE_SUMS"] = "1"
            f.run_generations(40, tune=True)
            f.run_generations(20, tune=False, sample_frequency=4)
        scf, hyf = f.chain_read()
        full = f.get_state(0)
        assert scf.hyper_ctr == sc.hyper_ctr and scf.gen == sc.gen == 120, (scf.hyper_ctr, sc.hyper_ctr)
        for name in ("s0", "s1", "tau", "spike_probability", "inactive_means", "inactive_variances", "weight_active"):
            assert np.isclose(getattr(hy, name), getattr(hyf, name), rtol=1e-9), name
        assert np.allclose(hy.alpha_r[:R], hyf.alpha_r[:R], rtol=1e-9)
        for k in ("alloc_active_inactive", "inactive_spike_allocation"):
            assert np.array_equal(mine[k], full[k][lo:hi]), k
        assert np.allclose(mine["Yg"], full["Yg"][lo:hi], rtol=1e-9, atol=1e-12)
        assert np.allclose(mine["variance_g"], full["variance_g"][lo:hi], rtol=1e-9)
        print(f"multi-GPU device-resident chain OK: world={world}, 120 generations, hyper_ctr={sc.hyper_ctr}, tau={hy.tau:.6f} "
              f"(single GPU {hyf.tau:.6f}), alpha[0]={hy.alpha_r[0]:.5f}")
    dist.all_reduce(ok)
    # ---- the host mirror on gene shards: burnin()/mcmc() with schedule="device"
    from zigzag_b200.zigzag import zigzag
    mm = zigzag(X, gl, threshold_i=-1.0, threshold_a=thr_a, seed=9, device=local, shard=(rank, world), process_group=dist.group.WORLD,
                write_settings=False, output_directory="/tmp/zz_multi_gpu_check")
    mm.burnin(sample_frequency=20, ngen=80, write_to_files=False, schedule="device")
    mm.mcmc(sample_frequency=5, ngen=40, write_to_files=False, compute_probs=False, schedule="device")
    v = torch.tensor([mm.tau, mm.s0, mm.weight_active, float(mm.gen), float(mm.n_samples)], dtype=torch.float64, device="cuda")
    allv = [torch.zeros_like(v) for _ in range(world)]
    dist.all_gather(allv, v)
    assert all(torch.equal(allv[0], x) for x in allv), "host mirror: ranks diverged"
    assert mm.gen == 81 + 41 and len(mm.prob_active()) == hi - lo
    if rank == 0:
        print(f"host mirror, schedule=device on {world} shards OK: gen={mm.gen}, tau={mm.tau:.5f}, samples={mm.n_samples}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
