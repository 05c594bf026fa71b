"""The host-side mirror of the reference's object API (zigzag_b200.zigzag) driving the GPU: north-star check (3) — posterior
probabilities of active expression on the bundled lung data agree with the reference's published values within Monte-Carlo
error — plus the output-file formats and the reference 15-move schedule."""
import os
import numpy as np
import pytest
from zigzag_b200.zigzag import zigzag

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def lung():
    d = np.load(os.path.join(HERE, "golden", "lung.npz"))
    return d, d["tpm"], d["gene_length"], [str(g) for g in d["gene_names"]]


def test_lung_prob_active_matches_tutorial_fused_schedule(tmp_path):
    d, tpm, gl, names = lung()
    mm = zigzag(tpm, gl, num_active_components=2, threshold_i=float(d["threshold_i"]), threshold_a=d["threshold_a"],
                output_directory=str(tmp_path), gene_names=names, seed=3)
    mm.burnin(sample_frequency=50, ngen=3000, burninprefix="lung", schedule="fused")
    mm.mcmc(sample_frequency=5, ngen=6000, mcmcprefix="lung_run1", schedule="fused")
    pa = mm.prob_active()
    pinned = d["tutorial_prob_active_first10"]
    assert np.abs(pa[:10] - pinned).max() < 0.08, (pa[:10], pinned)
    assert -10900 < mm.lnl_trace[-1] < -10100
    assert -2.4 < mm.s0 < -1.7 and -0.45 < mm.s1 < -0.27 and 0.7 < mm.tau < 1.05
    assert 0.05 < mm.spike_probability < 0.2 and -2.5 < mm.inactive_means < -1.4 and 2.2 < mm.active_means[0] < 3.1
    assert 0.64 < mm.weight_active < 0.8 and all(8 < a < 25 for a in mm.alpha_r)
    # tuned per-gene step sizes moved away from 0.5 and stayed inside the clamps (U:202-206)
    ty = mm.tuningParam_yg
    assert ty.min() >= 1e-4 and ty.max() <= 10 and np.std(ty) > 0.01
    # ---- output files keep the reference's formats (R/zigzagFileWriting.R)
    out = os.path.join(str(tmp_path), "lung_run1_mcmc_output")
    hdr = open(os.path.join(out, "lung_run1_model_parameters.log")).readline().rstrip("\n").split("\t")
    assert hdr == ["gen", "lnl", "s0", "s1", "tau", "alpha_r_library_1", "alpha_r_library_2", "alpha_r_library_3", "alpha_r_library_4",
                   "spike_probability", "inactive_mean", "inactive_variance", "active_mean_component1", "active_mean_component2",
                   "active_variance_component1", "active_variance_component2", "weight_active", "weight_within_active_component_1",
                   "weight_within_active_component_2"]
    rows = open(os.path.join(out, "lung_run1_model_parameters.log")).read().splitlines()[1:]
    assert len(rows) >= 1000 and all(len(r.split("\t")) == len(hdr) for r in rows[:20])
    pa_file = open(os.path.join(out, "lung_run1_probability_active.tab")).read().splitlines()
    assert pa_file[0] == "gene\tprob_active" and len(pa_file) == 5001 and pa_file[1].split("\t")[0] == names[0]
    assert abs(float(pa_file[1].split("\t")[1]) - pa[0]) < 1e-9
    yg_hdr = open(os.path.join(out, "lung_run1_yg_candidate_genes.log")).readline().split("\t")
    assert yg_hdr[0] == "gen" and len(yg_hdr) == 5001
    comp = open(os.path.join(out, "lung_run1_probability_in_component.tab")).readline().rstrip("\n")
    assert comp == "gene\tprob_0\tprob_1\tprob_2"
    assert os.path.exists(os.path.join(str(tmp_path), "hyperparameter_settings.txt"))


def test_reference_schedule_runs_the_15_move_table(tmp_path):
    d, tpm, gl, names = lung()
    mm = zigzag(tpm, gl, num_active_components=2, threshold_i=float(d["threshold_i"]), threshold_a=d["threshold_a"],
                output_directory=str(tmp_path), gene_names=names, seed=5)
    assert len(mm.proposal_list) == 15 and len(mm.proposal_probs) == 15
    assert list(mm.proposal_probs[:9]) == [4, 30, 5, 6, 0, 5, 5, 2, 5] and mm.proposal_probs[9] == 2 and mm.proposal_probs[14] == pytest.approx(4.4)
    lnl0 = mm.calculate_lnl()
    mm.burnin(sample_frequency=100, ngen=1200, write_to_files=False, schedule="reference")
    mm.mcmc(sample_frequency=10, ngen=300, write_to_files=False, schedule="reference")
    assert mm.gen >= 1500 and mm.n_samples >= 30
    assert mm.lnl_trace[-1] > lnl0 - 500 and mm.lnl_trace[-1] > -12500        # the chain climbed towards the posterior mode
    pa = mm.prob_active()
    assert pa[3] > 0.9 and pa[4] < 0.1                                          # clear-cut genes of the tutorial table
    # caches stayed consistent with a fresh recomputation through ~20 000 move calls
    st = mm.state()
    mm.h.refresh_caches()
    fresh = mm.h.get_state(0)
    fin = np.isfinite(fresh["XgLikelihood"])
    assert np.allclose(st["XgLikelihood"][fin], fresh["XgLikelihood"][fin], rtol=1e-9, atol=1e-9)
    out = st["inactive_spike_allocation"] == 0
    assert np.allclose(st["variance_g_probability"][out], fresh["variance_g_probability"][out], rtol=1e-9, atol=1e-9)


def test_lung_prob_active_matches_tutorial_device_schedule(tmp_path):
    """schedule="device": burnin() and mcmc() run their generations through zz_run_generations (sweep + every hyper move + tuning +
    allocation trace on the GPU, the host reads scalars back once per sampling point).  Same pins as the fused schedule."""
    d, tpm, gl, names = lung()
    mm = zigzag(tpm, gl, num_active_components=2, threshold_i=float(d["threshold_i"]), threshold_a=d["threshold_a"],
                output_directory=str(tmp_path), gene_names=names, seed=3)
    mm.burnin(sample_frequency=50, ngen=3000, burninprefix="lung", schedule="device")
    assert mm.gen == 3001 and mm.tuningParam_tau != 0.05
    mm.mcmc(sample_frequency=5, ngen=6000, mcmcprefix="lung_dev", schedule="device")
    pa = mm.prob_active()
    pinned = d["tutorial_prob_active_first10"]
    assert np.abs(pa[:10] - pinned).max() < 0.08, (pa[:10], pinned)
    assert -10900 < mm.lnl_trace[-1] < -10100
    assert -2.4 < mm.s0 < -1.7 and -0.45 < mm.s1 < -0.27 and 0.7 < mm.tau < 1.05
    assert 0.05 < mm.spike_probability < 0.2 and -2.5 < mm.inactive_means < -1.4 and 2.2 < mm.active_means[0] < 3.1
    assert 0.64 < mm.weight_active < 0.8 and all(8 < a < 25 for a in mm.alpha_r)
    rows = open(os.path.join(str(tmp_path), "lung_dev_mcmc_output", "lung_dev_model_parameters.log")).read().splitlines()[1:]
    assert len(rows) >= 1000
    assert mm.n_samples == 1 + sum(1 for g in range(3001, 9002) if g % 5 == 0)
    # the rows were written by the device during ONE segment (zz_set_sample_sink) and formatted afterwards
    assert len(rows) == sum(1 for g in range(3001, 9002) if g % 5 == 0) and rows[0].split("\t")[0] == "3005"
    yg = open(os.path.join(str(tmp_path), "lung_dev_mcmc_output", "lung_dev_yg_candidate_genes.log")).read().splitlines()
    assert len(yg[0].split("\t")) == 5001 and 400 <= len(yg) - 1 <= 501 and len(yg[1].split("\t")) == 5001


@pytest.mark.parametrize("schedule", ["fused", "device"])
def test_posterior_predictive_logs(tmp_path, schedule):
    """mcmc(run_posterior_predictive=TRUE): one simulated data set per (postpred_freq * sample_frequency) generations, realised
    discrepancies appended to the two log files with the reference's headers (FileWriting:44-55, PostPred:82-88)."""
    d, tpm, gl, names = lung()
    mm = zigzag(tpm, gl, num_active_components=2, threshold_i=float(d["threshold_i"]), threshold_a=d["threshold_a"],
                output_directory=str(tmp_path), gene_names=names, seed=4, write_settings=False)
    mm.burnin(sample_frequency=50, ngen=800, write_to_files=False, schedule="device")
    mm.mcmc(sample_frequency=5, ngen=200, mcmcprefix="pp", schedule=schedule, run_posterior_predictive=True, compute_probs=False)
    out = os.path.join(str(tmp_path), "pp_mcmc_output")
    a = open(os.path.join(out, "pp.post_predictive_output.log")).read().splitlines()
    b = open(os.path.join(out, "pp.library_specific_post_predictive_output.log")).read().splitlines()
    assert a[0].split("\t") == ["gen", "Wass_Lower", "Wass_Upper", "Rums"]
    assert b[0].split("\t") == ["gen"] + [f"Scaled_Wass_Lower_library_{r}" for r in (1, 2, 3, 4)] + [f"Rums_library_{r}" for r in (1, 2, 3, 4)]
    rows = np.array([[float(x) for x in l.split("\t")] for l in a[1:]])
    assert len(rows) == 40 and len(b) == 41 and np.all(rows[:, 0] % 5 == 0) and np.all(np.isfinite(rows))
    # realised minus simulated discrepancies of an (almost) converged chain scatter around zero, far from the scale of the data
    assert np.abs(rows[:, 1]).max() < 1.0 and np.abs(rows[:, 2]).max() < 1.0 and np.abs(rows[:, 3]).max() < 0.1
    assert np.std(rows[:, 1]) > 0


def test_lung_prob_active_full_length_device_schedule(tmp_path):
    """North-star check (3) at the tutorial's full length and beyond: 5 000 burn-in + 100 000 sampling generations (the tutorial:
    20 000) through the device-resident loop, 10 000 samples of the allocation — Monte-Carlo error ~0.006 per probability, so the
    ten published values must be met within 0.03 (the oracle's own long runs: tests/golden/lung_long_run_oracle.json)."""
    d, tpm, gl, names = lung()
    worst = []
    for seed in (3, 4):
        mm = zigzag(tpm, gl, num_active_components=2, threshold_i=float(d["threshold_i"]), threshold_a=d["threshold_a"],
                    output_directory=str(tmp_path), gene_names=names, seed=seed, write_settings=False)
        mm.burnin(sample_frequency=50, ngen=5000, write_to_files=False, schedule="device")
        mm.mcmc(sample_frequency=10, ngen=100000, write_to_files=False, compute_probs=False, schedule="device")
        pa = mm.prob_active()
        worst.append(float(np.abs(pa[:10] - d["tutorial_prob_active_first10"]).max()))
        assert mm.n_samples > 9990 and -10900 < mm.lnl_trace[-1] < -10100
    assert max(worst) < 0.03, worst
