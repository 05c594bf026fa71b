#!/usr/bin/env python
"""bench.py — gene-updates/s of zigzag's per-gene MCMC sweep on B200 (BASELINE.json metric).

One "step" = one fused sweep (mhYg -> mhVariance_g -> allocation Gibbs (+within) -> mhSpikeAllocation) over every gene
of the workload plus the sufficient-statistic reduction that feeds the hyper-parameter moves; with N > 1 GPUs each rank
owns a gene shard and the step ends with ONE NCCL all-reduce of the ~19-double statistics vector.
Workload at N=1: BASELINE.json configs[3], synthetic 1M genes x 16 replicates (the configuration the metric is quoted on).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload 1m_x16|200k_x8|20k_x2|64ch_60k_x4]

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU oracle port of the reference's R code (R itself is
not installable here: no R in the image, no network) with all host threads on the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "gene_updates_per_sec"
UNIT = "gene-updates/s"


def b_alg(R):
    """Algorithmic bytes per gene-update, SURVEY.md §8d: read Xg row 8R + gene_length 8 + Yg 8 + variance_g 8 + two tuning
    parameters 16 + flags 2; write Yg 8 + variance_g 8 + flags 2."""
    return 8 * R + 60


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the benchmark runs (B200_PROFILING.md recipe)."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        win, allr = [], []
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                rec = (float(f[2]), float(f[3]), f[5:9])
            except ValueError:
                continue
            allr.append(rec)
            if t0 - 0.05 <= ts <= t1 + 0.15:
                win.append(rec)
        use = win if win else allr
        if not use:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in use for i in range(4) if r[2][i].lower().startswith("active")})
        return {"sm_mhz": float(np.median([r[0] for r in use])), "sm_max_mhz": float(max(r[1] for r in use)),
                "reasons": reasons, "samples": len(use), "window": "timed region" if win else "whole run"}


def make_oracle_case(X, gl, hd, st):
    from oracle import oracle as O
    kw = {k: float(hd[k]) for k in ("s0", "s1", "tau", "spike_probability", "inactive_means", "inactive_variances",
                                    "weight_active", "beta", "temperature", "variance_g_upper_bound")}
    ho = O.Hyper.make(hd["alpha_r"], hd["active_means"], hd["active_variances"], hd["weight_within_active"], **kw)
    s = O.State(X, gl, st["Yg"], st["variance_g"], st["alloc_active_inactive"], st["alloc_within_active"],
                st["inactive_spike_allocation"])
    O.refresh_caches(ho, s, omp=True)
    return O, ho, s


def cpu_sweep_rate(X, gl, hd, st, sample_genes, steps, warmup, seed=1):
    """Times the oracle's fused sweep (C port of the reference's R arithmetic, OpenMP over genes) on a bounded sample."""
    n = min(sample_genes, X.shape[0])
    sub = {k: v[:n] for k, v in st.items()}
    from oracle import oracle as _O
    cores = _O.set_threads(os.cpu_count() or 1)      # torchrun exports OMP_NUM_THREADS=1: ask for every host thread explicitly
    O, ho, s = make_oracle_case(np.asfortranarray(X[:n]), gl[:n], hd, sub)
    times = []
    for it in range(warmup + steps):
        d = O.sweep_draws(seed, 0, it, 0, n)
        t = time.perf_counter()
        O.sweep(ho, s, d, omp=True, want_decisions=False)
        dt = time.perf_counter() - t
        if it >= warmup:
            times.append(dt)
    tot = float(np.sum(times))
    return n * len(times) / tot, cores, n, tot / len(times)


def run_reference(args, rank, world):
    """Reference arm: the CPU implementation of the path (oracle port; R is unavailable) on the box's host cores."""
    if rank != 0:
        return
    from zigzag_b200 import synth
    cid, G, R, chains = synth.CONFIGS[args.workload]
    X, gl, hd, st, _ = synth.make_config(args.workload)
    sample = min(G, args.cpu_sample)
    rate, cores, n, per = cpu_sweep_rate(X, gl, hd, st, sample, args.steps, max(1, min(args.warmup, 2)))
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"synthetic {G} genes x {R} replicates, K=2 (BASELINE.json configs[3])" if args.workload == "1m_x16" else args.workload,
                       "genes_per_step": n, "libraries": R},
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"fused sweep over the first {n} genes of the workload, oracle/zz_oracle.c with OpenMP ({cores} threads); R is not installed so the R package itself cannot be timed"},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="1m_x16")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--cpu-sample", type=int, default=1_000_000)
    ap.add_argument("--cpu-steps", type=int, default=4)
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--collective", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: statistics all-reduce fused into the reduction kernel over NVLink peer memory (default) or NCCL")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from zigzag_b200 import _capi, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: zigzag_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # one process per GPU: run (and first-touch the pinned host arrays of the e2e leg) on the CPUs of the GPU's own NUMA node
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
        except Exception:
            pass
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    cid, G_cfg, R, chains = synth.CONFIGS[args.workload]
    if args.scaling == "strong" and world > 1:
        G_tot = G_cfg
        g0, g1 = rank * G_tot // world, (rank + 1) * G_tot // world
        X, gl, hd, st, _ = synth.make_config(args.workload)
        X, gl, st = np.asfortranarray(X[g0:g1]), gl[g0:g1], {k: v[g0:g1] for k, v in st.items()}
        G, offset = g1 - g0, g0
    else:  # weak: every rank owns a full-size shard of G_cfg genes (its own synthetic draw), total = world * G_cfg
        G, offset, G_tot = G_cfg, rank * G_cfg, world * G_cfg
        Xr, glr, _ = synth.simulate(G, R, cid + 1000 * rank)
        hd = synth.truth_hyper_dict(R)
        st = synth.initial_state(Xr, glr, hd, synth.SEED0 + 1000 + cid + 7919 * rank)
        X, gl = Xr, glr

    hg = _capi.Hyper.make(hd["alpha_r"], hd["active_means"], hd["active_variances"], hd["weight_within_active"],
                          **{k: float(hd[k]) for k in ("s0", "s1", "tau", "spike_probability", "inactive_means",
                                                       "inactive_variances", "weight_active", "beta", "temperature",
                                                       "variance_g_upper_bound")})
    h = _capi.Handle(G, R, 2, num_chains=chains, device=local_rank, seed=20261019, gene_offset=offset)
    stream = torch.cuda.current_stream()
    h.set_stream(stream.cuda_stream)
    h.upload_data(X, gl)
    for c in range(chains):
        h.set_hyper(hg, c)
        h.set_state(c, **st)
    h.refresh_caches()
    h.sync()

    nstat = h.nstat
    stats = torch.zeros(chains * nstat, dtype=torch.float64, device=dev)

    use_p2p = world > 1 and args.collective == "p2p"
    if use_p2p:     # CUDA IPC handles of every rank's inbox, exchanged once over the process group
        blobs = [None] * world
        dist.all_gather_object(blobs, h.comm_init(rank, world))
        h.comm_connect(blobs)
        dist.barrier()

    # the sweep kernel's last CTA reduces the per-CTA statistics (and all-reduces them over NVLink peer memory when N > 1)
    # straight into `stats`: ONE launch per step
    h.set_stats_output_dev(stats.data_ptr(), allreduce=use_p2p)

    def reduce_stats():
        if world > 1 and not use_p2p:
            dist.all_reduce(stats)                        # the one collective of the path, NCCL variant

    def step(i):
        h.sweep(i)                              # 1 launch: k_sweep_fast (four moves + statistics + final reduction [+ all-reduce])
        reduce_stats()

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = h.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t_wall0 = time.time()
    ev0.record()
    for i in range(args.steps):                 # nothing but the K steps between the two events: consecutive sweeps are launched
        step(args.warmup + i)                   # programmatically dependent, an event record in between would serialise them
    ev1.record()
    torch.cuda.synchronize()
    t_wall1 = time.time()
    if world > 1:
        dist.barrier()
    launches = h.launch_count - launches0
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    value = G_tot * chains * args.steps / (ms_total * 1e-3)
    st_host = stats.cpu().numpy().reshape(chains, nstat)
    assert abs(st_host[0][:3].sum() - G_tot) < 0.5, "component counts do not sum to the number of genes"
    if use_p2p:     # the fused all-reduce must agree with NCCL on the same local vectors and be bit-identical on every rank
        loc = torch.zeros_like(stats)
        h.suffstats_dev(loc.data_ptr())
        dist.all_reduce(loc)
        assert torch.allclose(loc, stats, rtol=1e-12, atol=1e-9), "peer-memory all-reduce disagrees with NCCL"
        both = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(both, stats)
        assert all(torch.equal(both[0], b) for b in both), "ranks hold different reduced vectors"
        assert h.nan_flag() == 0

    if world == 1 or use_p2p:
        # one launch per step and nothing else on the stream: the sweep kernel's average duration IS the step time
        sweep_ms = ms_total / args.steps
    else:
        # NCCL variant: the step also holds the all-reduce kernel; time the sweep launches alone over the same number of steps
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        for i in range(args.steps):
            h.sweep(args.warmup + args.steps + i)
        k1.record()
        torch.cuda.synchronize()
        sweep_ms = k0.elapsed_time(k1) / args.steps

    # ---- e2e: the same sweep through the host-buffer entry point (mutable state in pinned host memory, H2D + D2H every step)
    e2e = None
    if chains == 1:
        f64 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).pin_memory()
        i32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).pin_memory()
        cur = h.get_state()
        host = [f64(cur["Yg"]), f64(cur["variance_g"]), i32(cur["alloc_active_inactive"]), i32(cur["alloc_within_active"]),
                i32(cur["inactive_spike_allocation"]), f64(cur["XgLikelihood"]), f64(cur["variance_g_probability"]),
                f64(cur["tuningParam_yg"]), f64(cur["tuningParam_variance_g"])]
        ptrs = [x.data_ptr() for x in host]
        nb = (0, 0)
        for i in range(2):
            nb = h.sweep_host(10_000 + i, False, *ptrs)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for i in range(args.e2e_steps):
            nb = h.sweep_host(20_000 + i, False, *ptrs)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": G_tot * args.e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": nb[0] * world, "d2h_bytes_per_step": nb[1] * world,
               "steps": args.e2e_steps, "what": "zz_sweep_host: mutable per-gene state (Yg, variance_g, tuning parameters, allocations) up from pinned host arrays, fused sweep, updated state + both likelihood caches back; Xg resident since zz_upload_data"}

    # ---- informational: a WHOLE generation (sweep + the nine hyper-parameter moves + commits) device-resident (zz_run_generations)
    chain = None
    if chains == 1:
        try:
            thr_a = (1.0, 6.0)[:len(hd["active_means"])]
            h.chain_begin(_capi.Priors.defaults(-1.0, thr_a), _capi.ChainScalars.initial(hd["active_means"], thr_a, R))
            h.run_generations(10, tune=False, sample_frequency=10)
            h.sync()
            if world > 1:
                dist.barrier()
            l0, t0 = h.launch_count, time.perf_counter()
            h.run_generations(100, tune=False, sample_frequency=10)
            h.sync()
            dtg = (time.perf_counter() - t0) / 100
            chain = {"generation_us": dtg * 1e6, "generations_per_sec": 1.0 / dtg, "launches_per_generation": (h.launch_count - l0) / 100,
                     "what": "zz_run_generations: fused sweep + gibbsMixtureWeights, mhInactiveMeans, mhActiveMeansDif, mhSharedVariance, "
                             "gibbsSpikeProb, mhTau, mhSg, mhS0Tau, mhP_x on the device, host only enqueues (this rank's wall clock)"}
        except Exception as e:      # never let the informational leg break the contract line
            chain = {"error": str(e)[:200]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    balg = b_alg(R)
    achieved = G * chains * balg / (sweep_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            with open(tp) as f:
                traffic = json.load(f).get(f"k_sweep:{args.workload}")
        except Exception:
            traffic = None
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, cores, n, per = cpu_sweep_rate(X, gl, hd, st, args.cpu_sample, args.cpu_steps, 1)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{args.cpu_steps} fused sweeps over the first {n} genes of the workload, oracle/zz_oracle.c + OpenMP ({cores} threads, {per:.3f} s/sweep)"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"synthetic {G_cfg} genes x {R} replicates, K=2 active components, {chains} chain(s) (BASELINE.json configs: {args.workload})",
                       "genes_per_gpu": G, "genes_total": G_tot, "libraries": R, "chains": chains,
                       "step": "fused sweep (mhYg, mhVariance_g, allocation, spike allocation) + suff-stat reduction" + ((" fused with a peer-memory all-reduce over NVLink" if use_p2p else " + NCCL all-reduce") if world > 1 else ""),
                       "l2": f"working set {(G * (8 * R + 8) + G * chains * 50) / 1e6:.0f} MB per GPU vs 126 MB L2: inputs larger than L2, no flush" if G * (8 * R + 58) > 126e6 else "working set fits in L2 (reported as-is; see DESIGN.md)",
                       "rng": "Philox4x32-10 on device"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": "k_sweep_fast", "kernel_ms": sweep_ms, "bytes_per_gene_update": balg, "peak_source": peak_src},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "chain": chain}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
