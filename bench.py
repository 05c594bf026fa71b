#!/usr/bin/env python
"""bench.py — gene-updates/s of zigzag's per-gene MCMC sweep on B200 (BASELINE.json metric).

One "step" = one fused sweep (mhYg -> mhVariance_g -> allocation Gibbs (+within) -> mhSpikeAllocation) over every gene of the
workload plus the sufficient-statistic reduction that feeds the hyper-parameter moves; with N > 1 GPUs each rank owns a gene
shard and every step ends with the all-reduce of the ~19-entry statistics vector, done inside the kernel over
NVLink peer memory.  Workload: BASELINE.json configs[3], synthetic 1M genes x 16 replicates, gene-sharded over the N GPUs
(strong scaling, as the configuration is named); `--scaling weak` gives every GPU a full 1M-gene shard instead.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload 1m_x16|200k_x8|20k_x2|64ch_60k_x4]
                  [--scaling strong|weak] [--mode persistent|launch] [--blocks B]

Timing protocol: B blocks (default 5) of exactly K steps each, every block bracketed by CUDA events on the
launching stream; per block the MAX over ranks is taken, the line reports the MEDIAN block (all block times are in the line).
In the default mode a block is ONE cooperative launch of the persistent kernel (zz_run_sweeps: K sweeps, a grid barrier and the
statistics [all-reduce] after each); `--mode launch` issues K programmatically dependent zz_sweep launches instead.  Right
before each timed block a short untimed launch of the same kernel is enqueued on the same stream: its in-kernel exchange aligns
the ranks on the device, so no host-side start skew falls into the timed region.

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU oracle port of the reference's R code (R itself is not
installable here: no R in the image, no network) with all host threads on the same workload.
"""
import argparse
import json
import os
import shutil
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
CONFIG_NAMES = {"20k_x2": "configs[1]", "200k_x8": "configs[2]", "1m_x16": "configs[3]", "64ch_60k_x4": "configs[4]"}


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


def config_dict(workload, scaling, world):
    """The `config` object: built identically by both arms (ours and --impl reference) from the command line alone."""
    from zigzag_b200 import synth
    cid, G, R, chains = synth.CONFIGS[workload]
    tot = G * (world if scaling == "weak" else 1)
    return {"workload": f"synthetic {G} genes x {R} replicates, K=2 active components, {chains} chain(s) (BASELINE.json {CONFIG_NAMES[workload]}: {workload})",
            "genes_total": tot, "libraries": R, "chains": chains, "scaling": scaling,
            "step": "fused sweep (mhYg, mhVariance_g, allocation Gibbs, spike allocation) over every gene + suff-stat reduction"
                    + (" + all-reduce of the statistics over the gene shards" if world > 1 and chains == 1 else ""),
            "rng": "Philox4x32-10, counter-based, keyed by (seed, chain, step, global gene index)"}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the benchmark runs (B200_PROFILING.md recipe)."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
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


def numpy_sweep_rate(X, gl, hd, st, sample_genes, steps, seed=1):
    """The stand-in for single-threaded, vectorised R: the NumPy restatement of the R vector code (oracle/oracle_np.py:
    mhYg, mhVariance_g, gibbsAllocationActiveInactive + WithinActive — three of the four per-gene moves, the spike move touches
    ~10 % of the genes) on ONE thread, bounded sample."""
    from oracle import oracle_np as N
    n = min(sample_genes, X.shape[0])
    O, ho, so = make_oracle_case(np.asfortranarray(X[:n]), gl[:n], hd, {k: v[:n] for k, v in st.items()})
    Xs, gls = np.ascontiguousarray(X[:n]), np.ascontiguousarray(gl[:n])
    s = {k: np.array(v[:n]) for k, v in st.items()}
    s.update(no_detect_spike=so.no_detect_spike.copy(), XgLikelihood=so.XgLikelihood.copy(), variance_g_probability=so.variance_g_probability.copy())
    h = dict(hd)
    t_all = []
    try:
        import threadpoolctl
        lim = threadpoolctl.threadpool_limits(1)
    except Exception:
        lim = None
    try:
        for it in range(steps + 1):
            d = O.sweep_draws(seed, 0, it, 0, n)
            t = time.perf_counter()
            N.mhYg(h, Xs, gls, s, d[0], d[1])
            N.mhVariance_g(h, Xs, gls, s, d[2], d[3])
            N.gibbsAllocationActiveInactive(h, s, d[4], d[5])
            if it > 0:
                t_all.append(time.perf_counter() - t)
    finally:
        if lim is not None:
            lim.restore_original_limits()
    per = float(np.mean(t_all))
    return n / per, n, per


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
            "warmup": args.warmup, "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config_dict(args.workload, args.scaling, world),
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"fused sweep over the first {n} genes of the workload, oracle/zz_oracle.c with OpenMP ({cores} threads); "
                                       f"R is not installed (Rscript: {shutil.which('Rscript')}) so the R package itself cannot be timed"},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--blocks", type=int, default=5, help="timed blocks of --steps steps each; the median block is reported")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="1m_x16", choices=sorted(CONFIG_NAMES))
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="strong (default): the named workload is sharded over the GPUs (genes; chains for the batched-chains "
                         "workload); weak: every GPU owns a full copy of the workload's size")
    ap.add_argument("--mode", default="persistent", choices=["persistent", "launch"],
                    help="persistent: a block is one cooperative launch of k_generations (zz_run_sweeps); launch: one zz_sweep launch per step")
    ap.add_argument("--cpu-sample", type=int, default=1_000_000)
    ap.add_argument("--cpu-steps", type=int, default=4)
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the informational legs (generation rate, unsaturated data, weak-scaling leg)")
    ap.add_argument("--collective", default="p2p", choices=["p2p", "nccl"],
                    help="--mode launch, N > 1: statistics all-reduce fused into the sweep kernel over NVLink peer memory (default) or NCCL")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    args.blocks = max(args.blocks, 1)

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
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def maxr(x):
        """MAX over ranks of a list of floats (device-timed values)."""
        t = torch.tensor(list(x), dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    cid, G_cfg, R, chains_cfg = synth.CONFIGS[args.workload]
    X, gl, hd, st, _ = synth.make_config(args.workload)
    X_full, gl_full, st_full = X, gl, st
    shard_chains = chains_cfg > 1                       # the batched-chains workload shards CHAINS (no data-path collective)
    chains, seed = chains_cfg, 20261019
    if args.scaling == "strong" and world > 1:
        G_tot = G_cfg
        if shard_chains:
            if chains_cfg % world:
                raise SystemExit("--gpus must divide the number of chains")
            chains, G, offset, seed = chains_cfg // world, G_cfg, 0, seed + 7919 * rank
        else:
            g0, g1 = rank * G_tot // world, (rank + 1) * G_tot // world
            X, gl, st = np.asfortranarray(X[g0:g1]), gl[g0:g1], {k: v[g0:g1] for k, v in st.items()}
            G, offset = g1 - g0, g0
    else:   # weak (or N = 1): every rank owns the full workload; ranks differ by their Philox key / global gene offset
        G, G_tot = G_cfg, world * G_cfg
        offset, seed = (0, seed + 7919 * rank) if shard_chains else (rank * G_cfg, seed)
    updates_per_step = G_tot * chains_cfg               # gene-updates of one step over the whole job
    use_p2p = world > 1 and not shard_chains and (args.mode == "persistent" or args.collective == "p2p")

    def make_handle(Xh, glh, sth, nchains, off, sd):
        hg = _capi.Hyper.make(hd["alpha_r"], hd["active_means"], hd["active_variances"], hd["weight_within_active"],
                              **{k: float(hd[k]) for k in ("s0", "s1", "tau", "spike_probability", "inactive_means",
                                                           "inactive_variances", "weight_active", "beta", "temperature",
                                                           "variance_g_upper_bound")})
        hh = _capi.Handle(Xh.shape[0], R, 2, num_chains=nchains, device=local_rank, seed=sd, gene_offset=off)
        hh.set_stream(torch.cuda.current_stream().cuda_stream)
        hh.upload_data(Xh, glh)
        for c in range(nchains):
            hh.set_hyper(hg, c)
            hh.set_state(c, **sth)
        hh.refresh_caches()
        hh.sync()
        return hh, hg

    def connect(hh):
        blobs = [None] * world
        dist.all_gather_object(blobs, hh.comm_init(rank, world))
        hh.comm_connect(blobs)
        dist.barrier()

    h, hg = make_handle(X, gl, st, chains, offset, seed)
    nstat = h.nstat
    stats = torch.zeros(chains * nstat, dtype=torch.float64, device=dev)
    if use_p2p:
        connect(h)      # CUDA IPC handles of every rank's inbox, exchanged once over the process group
    # the statistics of every step land in `stats` (all-reduced over NVLink peer memory when N > 1): written by the kernel itself
    h.set_stats_output_dev(stats.data_ptr(), allreduce=use_p2p)

    def timed_blocks(hh, nblocks, K, step0, persistent):
        """Enqueues nblocks x ([untimed aligning launch] [event] K steps [event]); returns per-block ms (this rank) and the launches timed."""
        evs, step, l_timed = [], step0, 0
        for b in range(nblocks):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if persistent:
                hh.run_sweeps(step, 2); step += 2             # untimed: aligns the ranks on the device, keeps the clocks up
                e0.record()
                l0 = hh.launch_count
                hh.run_sweeps(step, K); step += K
            else:
                hh.sweep(step); step += 1
                if world > 1 and not use_p2p and not shard_chains:
                    dist.all_reduce(stats)
                e0.record()
                l0 = hh.launch_count
                for i in range(K):          # nothing but the K steps between the two events: consecutive sweeps are launched
                    hh.sweep(step + i)      # programmatically dependent, an event record in between would serialise them
                    if world > 1 and not use_p2p and not shard_chains:
                        dist.all_reduce(stats)                # the one collective of the path, NCCL variant
                step += K
            e1.record()
            l_timed += hh.launch_count - l0
            evs.append((e0, e1))
        torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in evs], l_timed, step

    persistent = args.mode == "persistent"
    # ---- warm-up: W untimed steps
    _, _, step = timed_blocks(h, 1, args.warmup, 0, persistent)
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    torch.cuda.synchronize()
    t_wall0 = time.time()
    # clocks are sampled every 20 ms: repeat the timed protocol until the region is ~0.3 s long (every pass is timed the same way,
    # the reported blocks are those of the LAST pass; the number of passes is agreed between the ranks)
    ms_blocks, launches, step = timed_blocks(h, args.blocks, args.steps, step, persistent)
    passes = 1 + min(199, int(0.3 / max(maxr([time.time() - t_wall0])[0], 1e-4)))
    for _ in range(passes - 1):
        ms_blocks, launches, step = timed_blocks(h, args.blocks, args.steps, step, persistent)
    t_wall1 = time.time()
    if world > 1:
        dist.barrier()
    ms_blocks = maxr(ms_blocks)
    ms_total = float(np.median(ms_blocks))
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    value = updates_per_step * args.steps / (ms_total * 1e-3)
    st_host = stats.cpu().numpy().reshape(chains, nstat)
    n_comp = st_host[0][:3].sum()
    assert abs(n_comp - (G if shard_chains else G_tot)) < 0.5, f"component counts {n_comp} do not sum to the number of genes"
    if use_p2p:     # the in-kernel all-reduce must agree with NCCL on the same local vectors and be bit-identical on every rank
        loc = torch.zeros_like(stats)
        h.set_stats_output_dev(0, allreduce=False)
        h.suffstats_dev(loc.data_ptr())
        h.set_stats_output_dev(stats.data_ptr(), allreduce=True)
        dist.all_reduce(loc)
        keep = [j for j in range(nstat) if j not in (3 * 3 + 7, 3 * 3 + 8)]     # the two likelihood-cache sums are not carried by the persistent kernel
        assert torch.allclose(loc[keep], stats[keep], rtol=1e-9, atol=1e-4), "peer-memory all-reduce disagrees with NCCL"
        both = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(both, stats)
        assert all(torch.equal(both[0], b) for b in both), "ranks hold different reduced vectors"
        assert h.nan_flag() == 0, f"device flag {h.nan_flag()} (4 = all-reduce time-out)"

    sweep_ms = ms_total / args.steps
    kernel = "k_generations (zz_run_sweeps: one cooperative launch per block)" if persistent else "k_sweep_fast (zz_sweep: one launch per step)"
    if not persistent and world > 1 and not use_p2p and not shard_chains:
        # NCCL variant: the step also holds the all-reduce kernel; time the sweep launches alone over the same number of steps
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        for i in range(args.steps):
            h.sweep(step + i)
        k1.record()
        torch.cuda.synchronize()
        step += args.steps
        sweep_ms = k0.elapsed_time(k1) / args.steps

    extras = {}
    # ---- informational: the same protocol on data where NO tile takes the saturated-detection shortcut (alpha_r / 1000: every
    # gene evaluates the full per-library detection loop), and the share of genes that take it in the headline workload
    if not args.no_extras and persistent:
        try:
            saved = [h.get_state(c) for c in range(chains)]   # the chain wanders off under alpha_r / 1000: put the state back afterwards
            cur = saved[0]
            t_min = gl * np.exp(cur["Yg"]) * float(np.min(hd["alpha_r"]))
            sat = float(np.mean((t_min >= 37.5) | (cur["inactive_spike_allocation"] == 1)))
            hd2 = dict(hd); hd2["alpha_r"] = np.asarray(hd["alpha_r"]) / 1000.0
            hg2 = _capi.Hyper.make(hd2["alpha_r"], hd["active_means"], hd["active_variances"], hd["weight_within_active"],
                                   **{k: float(hd[k]) for k in ("s0", "s1", "tau", "spike_probability", "inactive_means",
                                                                "inactive_variances", "weight_active", "beta", "temperature",
                                                                "variance_g_upper_bound")})
            for c in range(chains):
                h.set_hyper(hg2, c)
            h.refresh_caches()
            _, _, step = timed_blocks(h, 1, args.warmup, step, True)
            msb, _, step = timed_blocks(h, args.blocks, args.steps, step, True)
            msb = float(np.median(maxr(msb)))
            for c in range(chains):
                h.set_hyper(hg, c)
                h.set_state(c, **saved[c])
            h.refresh_caches()
            del saved
            extras["data_dependence"] = {
                "genes_saturated_frac": sat,
                "unsaturated_value": updates_per_step * args.steps / (msb * 1e-3), "unsaturated_ms_per_step": msb / args.steps,
                "what": "genes_saturated_frac: share of this rank's genes whose detection term is exactly 0 / -Inf at the current Yg "
                        "(gene_length * exp(Yg) * min(alpha_r) >= 37.5, or in the spike) — tiles made of such genes skip the per-library loop; "
                        "unsaturated_*: the same timed protocol with alpha_r / 1000, where every gene runs the full loop (lower bound over data)"}
        except Exception as e:
            extras["data_dependence"] = {"error": str(e)[:200]}

    # ---- L2 check: the per-gene state the production sweep touches (Xg enters through per-gene sufficient statistics) can be smaller
    # than the 126 MB L2 and then stays there between the steps of a block, as it does between the generations of a real chain.
    # Timed here: isolated single-sweep launches, L2-warm and with a 512 MB buffer rewritten before every launch.
    l2_check = None
    if not args.no_extras:
        try:
            flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
            res = {}
            for cold in (False, True):
                ts = []
                for i in range(24):
                    if cold:
                        flush.fill_(i & 255)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); h.run_sweeps(step, 1); e1.record(); step += 1
                    torch.cuda.synchronize()
                    if i >= 4:
                        ts.append(e0.elapsed_time(e1))
                res["l2_flushed_ms" if cold else "l2_warm_ms"] = float(np.median(maxr(ts)) if world == 1 else np.median(ts))
            del flush
            touched = G * chains * 50 + G * 36
            res["touched_mb_per_sweep"] = touched / 1e6
            res["what"] = ("median of 20 isolated one-sweep launches of the timed kernel (each bracketed by events, a device synchronisation between launches, so the "
                           "per-launch ramp-up is included and the figures are larger than ms_per_step), without and with a 512 MB buffer rewritten before "
                           "every launch: the sweep is bound by instruction latency, its loads are staged one tile ahead, the L2 state makes no difference")
            l2_check = res
        except Exception as e:
            l2_check = {"error": repr(e)[:200]}

    # ---- e2e: the same sweep through the host-buffer entry point (mutable state in pinned host memory, H2D + D2H every step)
    e2e = None
    if chains_cfg == 1:
        f64 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).pin_memory()
        i32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).pin_memory()
        h.refresh_caches()
        cur = h.get_state()
        host = [f64(cur["Yg"]), f64(cur["variance_g"]), i32(cur["alloc_active_inactive"]), i32(cur["alloc_within_active"]),
                i32(cur["inactive_spike_allocation"]), f64(cur["XgLikelihood"]), f64(cur["variance_g_probability"]),
                f64(cur["tuningParam_yg"]), f64(cur["tuningParam_variance_g"])]
        ptrs = [x.data_ptr() for x in host]
        nb = (0, 0)
        h.set_stats_output_dev(0, allreduce=False)
        for i in range(3):
            nb = h.sweep_host(10_000_000 + i, False, *ptrs)
        torch.cuda.synchronize()
        e2e_blocks = []
        for b in range(3):
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for i in range(args.e2e_steps):
                nb = h.sweep_host(20_000_000 + b * 1000 + i, False, *ptrs)
            torch.cuda.synchronize()
            e2e_blocks.append(time.perf_counter() - t0)
        e2e_blocks = maxr(e2e_blocks)
        dt = float(np.median(e2e_blocks))
        nbs = torch.tensor([nb[0], nb[1]], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(nbs)
        e2e = {"value": G_tot * args.e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": int(nbs[0].item()), "d2h_bytes_per_step": int(nbs[1].item()),
               "steps": args.e2e_steps, "blocks_s": e2e_blocks,
               "what": "zz_sweep_host through the C ABI with HOST buffers: the mutable per-gene state (Yg, variance_g, tuning parameters, allocations) goes up "
                       "from pinned host arrays, fused sweep, updated state + both likelihood caches come back, every step (wall clock, max over ranks, "
                       "median of 3 blocks); the data matrix Xg stays resident on the device since zz_upload_data (like the data of the R object) and is NOT re-uploaded"}
        h.set_stats_output_dev(stats.data_ptr(), allreduce=use_p2p)

    # ---- informational: a WHOLE generation (sweep + all nine hyper-parameter moves) device-resident (zz_run_generations)
    chain = None
    if not args.no_extras and (world == 1 or use_p2p or shard_chains):
        try:
            thr_a = (1.0, 6.0)[:len(hd["active_means"])]
            h.chain_begin(_capi.Priors.defaults(-1.0, thr_a), [_capi.ChainScalars.initial(hd["active_means"], thr_a, R) for _ in range(chains)])
            chain = {"what": "zz_run_generations: fused sweep + gibbsMixtureWeights, mhInactiveMeans, mhActiveMeansDif, mhSharedVariance, gibbsSpikeProb, "
                             "mhTau, mhSg, mhS0Tau, mhP_x (+ tune_all every 400 generations in burn-in) on the device, the host only enqueues; "
                             "CUDA events around 200 generations, max over ranks; gene-updates/s = genes x generations / time"}
            for name, tune in (("burnin", True), ("sampling", False)):
                h.run_generations(20, tune=tune, sample_frequency=10)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                h.run_generations(2, tune=tune, sample_frequency=10)
                e0.record()
                h.run_generations(200, tune=tune, sample_frequency=10)
                e1.record()
                torch.cuda.synchronize()
                us = maxr([e0.elapsed_time(e1)])[0] * 1e3 / 200
                chain[name + "_generation_us"] = us
                chain[name + "_gene_updates_per_sec"] = updates_per_step / (us * 1e-6)
            chain["generation_us"] = chain["burnin_generation_us"]
            # the second end-to-end flavour: a sampling interval as the R drop-in's mcmc() runs it (r_overlay: gpu_run_generations, then
            # gpu_pull at the sample point, R/zigzagMCMC.R:194-217) — sample_frequency generations on the device, then the WHOLE mutable
            # state (52 B per gene), the hyper-parameters and the allocation-trace update, into pageable host arrays through the C ABI
            if chains_cfg == 1:
                try:
                    sf, n_int = 100, 3                    # the reference's default sample_frequency (R/zigzagBurnin.R:21)
                    h.run_generations(sf, tune=False, sample_frequency=sf); h.get_state(0); torch.cuda.synchronize()
                    if world > 1:
                        dist.barrier()
                    t0 = time.perf_counter()
                    for _ in range(n_int):
                        h.run_generations(sf, tune=False, sample_frequency=sf)
                        h.accumulate_alloc_trace()
                        pulled = h.get_state(0)
                        h.get_hyper(0)
                    dt = maxr([time.perf_counter() - t0])[0]
                    chain["e2e_sampling_interval"] = {
                        "value": G_tot * sf * n_int / dt, "unit": UNIT, "sample_frequency": sf, "intervals": n_int,
                        "us_per_generation": dt / (sf * n_int) * 1e6,
                        "d2h_bytes_per_interval": int(sum(v.nbytes for v in pulled.values())) * world,
                        "what": "wall clock (max over ranks) of 3 sampling intervals: 100 device-resident generations (zz_run_generations), then "
                                "zz_accumulate_alloc_trace + zz_get_state of every per-gene field + zz_get_hyper into pageable host arrays — what the "
                                "R drop-in does per sample; the state never goes up again because it lives on the device between samples"}
                    del pulled
                except Exception as e:
                    chain["e2e_sampling_interval"] = {"error": repr(e)[:200]}
            chain["nan_flag"] = h.nan_flag()
        except Exception as e:      # never let the informational leg break the contract line
            chain = {"error": repr(e)[:200]}

    # ---- informational, N > 1 in strong mode: the weak-scaling number of the same kernel (a full-size shard on every GPU)
    if not args.no_extras and world > 1 and args.scaling == "strong" and not shard_chains and persistent:
        try:
            h.close()
            hw, _ = make_handle(X_full, gl_full, st_full, 1, rank * G_cfg, seed)
            statsw = torch.zeros(nstat, dtype=torch.float64, device=dev)
            connect(hw)
            hw.set_stats_output_dev(statsw.data_ptr(), allreduce=True)
            _, _, s2 = timed_blocks(hw, 1, args.warmup, 0, True)
            msw, _, s2 = timed_blocks(hw, args.blocks, args.steps, s2, True)
            msw = float(np.median(maxr(msw)))
            assert abs(float(statsw[:3].sum().item()) - world * G_cfg) < 0.5
            extras["weak_scaling"] = {"value": world * G_cfg * args.steps / (msw * 1e-3), "ms_per_step": msw / args.steps, "genes_per_gpu": G_cfg,
                                      "what": "same protocol with a full-size gene shard on every GPU (the workload's data on every rank, distinct "
                                              "global gene offsets, so distinct Philox streams) and the same in-kernel all-reduce"}
            hw.close()
        except Exception as e:
            extras["weak_scaling"] = {"error": str(e)[:200]}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    balg = b_alg(R)
    achieved = G * chains * balg / (sweep_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp) and world == 1:   # the captures are of the one-GPU launch; a shard's launch moves less and was not captured
        try:
            with open(tp) as f:
                traffic = json.load(f).get(("k_generations:" if persistent else "k_sweep:") + args.workload)
            if isinstance(traffic, dict):        # persistent kernel: per launch of K sweeps, from the 2-sweep and 20-sweep captures
                traffic = traffic["first_sweep_bytes"] + traffic["per_additional_sweep_bytes"] * (args.steps - 1)
        except Exception:
            traffic = None
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, cores, n, per = cpu_sweep_rate(X_full, gl_full, hd, st_full, args.cpu_sample, args.cpu_steps, 1)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{args.cpu_steps} fused sweeps over the first {n} genes of the workload, oracle/zz_oracle.c + OpenMP ({cores} threads, {per:.3f} s/sweep)",
               "Rscript": shutil.which("Rscript")}
        try:
            nrate, nn, nper = numpy_sweep_rate(X_full, gl_full, hd, st_full, 100_000, 2)
            cpu["numpy_1core"] = {"value": nrate, "unit": UNIT, "cores": 1,
                                  "sample": f"2 passes of mhYg + mhVariance_g + allocation Gibbs over the first {nn} genes, oracle/oracle_np.py (the vectorised restatement "
                                            f"of the R code, the stand-in for single-threaded R), {nper:.3f} s/pass"}
        except Exception as e:
            cpu["numpy_1core"] = {"error": str(e)[:200]}

    cfg = config_dict(args.workload, args.scaling, world)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg,
            "detail": {"genes_per_gpu": G, "chains_per_gpu": chains, "mode": args.mode,
                       "sharding": ("chains over GPUs, no data-path collective" if shard_chains else "contiguous gene blocks, one all-reduce of %d doubles per step" % nstat) if world > 1 else "single GPU",
                       "collective": ("in-kernel exchange over NVLink peer memory (CUDA IPC inboxes)" if use_p2p else "NCCL all-reduce") if world > 1 and not shard_chains else None,
                       "timing": f"{args.blocks} blocks x {args.steps} steps, CUDA events per block on the launching stream, max over ranks per block, median block reported; "
                                 f"an untimed 2-step launch of the same kernel precedes every timed block on the same stream ({passes} pass(es) of the protocol, last one reported)",
                       "ms_blocks": ms_blocks,
                       "l2": f"inputs {(G * (8 * R + 8) + G * chains * 50) / 1e6:.0f} MB per GPU (data matrix + gene lengths + per-gene state) vs 126 MB L2; the sweep itself touches "
                             f"{(G * chains * 50 + G * 36) / 1e6:.0f} MB per step (the data matrix enters through per-gene sufficient statistics), which "
                             + ("exceeds L2" if G * chains * 50 + G * 36 > 126e6 else "fits in L2 and stays there between steps, as between the generations of a real chain")
                             + "; no flush between steps — l2_check times the kernel with and without an L2 flush before every launch",
                       "l2_check": l2_check},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": kernel, "kernel_ms": sweep_ms * (args.steps if persistent else 1), "sweeps_per_launch": args.steps if persistent else 1,
                         "bytes_per_gene_update": balg, "peak_source": peak_src},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "chain": chain}
    line.update(extras)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
