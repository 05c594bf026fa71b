"""Generates tests/golden/lung_long_run_oracle.json: the oracle's schedule-faithful chain on the bundled lung data at the tutorial's
full length (5 000 burn-in + 20 000 sampling generations, sample_frequency 10), three seeds, against the ten P(active) values the
reference's tutorial publishes (docs/tutorial.html:681-691).  ~75 s per seed on 8 host threads.  Run from the repository root:
    python tests/golden/make_lung_long_run.py"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle import oracle as O

d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "lung.npz"))
X = np.log(d["tpm"], where=d["tpm"] > 0, out=np.full(d["tpm"].shape, -np.inf))
gl = d["gene_length"] / d["gene_length"].mean()
pr = O.Priors.defaults(float(d["threshold_i"]), d["threshold_a"])
out = {"recipe": "oracle schedule-faithful chain (zzo_chain_burnin 5000 + zzo_chain_mcmc 20000, sample_frequency 10), bundled lung data, "
                 "threshold_i / threshold_a of the tutorial", "tutorial_first10": [float(x) for x in d["tutorial_prob_active_first10"]], "runs": []}
for seed in (1, 2, 3):
    t = time.time()
    c = O.Chain(X, gl, pr, 2, seed=seed, omp=True)
    c.burnin(5000); c.mcmc(20000, 10)
    pa, h = c.prob_active(), c.hyper
    out["runs"].append({"seed": seed, "prob_active_first10": [round(float(x), 4) for x in pa[:10]],
                        "max_abs_diff_vs_tutorial": float(np.abs(pa[:10] - d["tutorial_prob_active_first10"]).max()), "lnl": float(c.lnl()),
                        "s0": h.s0, "s1": h.s1, "tau": h.tau, "spike_probability": h.spike_probability, "inactive_means": h.inactive_means,
                        "weight_active": h.weight_active, "seconds": time.time() - t})
    print(out["runs"][-1], flush=True)
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "lung_long_run_oracle.json"), "w"), indent=1)
