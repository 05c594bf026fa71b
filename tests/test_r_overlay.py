"""The R-side binding (r_overlay/) cannot run here — R is not installed — so these tests pin what can be pinned statically:
the shim compiles against minimal stand-ins of the R headers (tests/r_stub/) with the real include/zigzag_gpu.h, its
registration table agrees with the definitions, every .Call in zigzagGPU.R names a registered routine with the right number
of arguments, every hyper-parameter move of the reference's proposal table is overridden, and the shim reaches every
entry point of the C ABI that an R caller can use."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "r_overlay", "src", "zz_shim.c")
RFILE = os.path.join(ROOT, "r_overlay", "R", "zigzagGPU.R")


def _table():
    src = open(SHIM).read()
    return {m.group(1): int(m.group(3)) for m in re.finditer(r'\{"(zzR_\w+)",\s*\(DL_FUNC\)&(\w+),\s*(\d+)\}', src)}, src


def test_shim_compiles_against_stub_r_headers():
    cmd = ["gcc", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-Wno-unused-parameter", "-Wno-cast-function-type",
           "-I", os.path.join(ROOT, "tests", "r_stub"), "-I", os.path.join(ROOT, "include"), SHIM]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_registration_table_matches_the_definitions():
    table, src = _table()
    defs = {}
    for m in re.finditer(r'^SEXP (zzR_\w+)\(([^)]*)\)', src, re.M):
        defs[m.group(1)] = len([a for a in m.group(2).split(",") if a.strip()])
    assert set(defs) == set(table), set(defs) ^ set(table)
    for name, n in table.items():
        assert defs[name] == n, (name, defs[name], n)
    for m in re.finditer(r'\{"(zzR_\w+)",\s*\(DL_FUNC\)&(\w+)', src):
        assert m.group(1) == m.group(2)


def _calls(text):
    """(.Call name, number of arguments after the name) for every .Call( in an R source, by bracket matching."""
    out = []
    for m in re.finditer(r'\.Call\("(\w+)"', text):
        i, depth, nargs, in_str = m.end(), 1, 0, None
        while depth:
            ch = text[i]
            if in_str:
                if ch == in_str:
                    in_str = None
            elif ch in "\"'":
                in_str = ch
            elif ch in "([{":
                depth += 1
            elif ch in ")]}":
                depth -= 1
            elif ch == "," and depth == 1:
                nargs += 1
            i += 1
        out.append((m.group(1), nargs))
    return out


def test_every_dot_call_is_registered_with_the_right_arity():
    table, _ = _table()
    rsrc = "\n".join(l.split("#")[0] for l in open(RFILE).read().splitlines())
    calls = _calls(rsrc)
    assert len(calls) > 20
    for name, n in calls:
        assert name in table, f"{name} is not registered in zz_shim.c"
        assert table[name] == n, f".Call(\"{name}\") passes {n} arguments, the shim takes {table[name]}"


def test_every_hyper_move_and_driver_hook_is_overridden():
    rsrc = open(RFILE).read()
    # the ten scalar entries of proposal_list with non-zero weight in the default model (R/zigzagInitialize.R:207-237) + the drivers' hooks
    for name in ("gibbsMixtureWeights", "mhInactiveMeans", "mhActiveMeansDif", "mhSharedVariance", "gibbsSpikeProb", "mhTau", "mhSg",
                 "mhS0Tau", "mhP_x", "calculate_lnl", "tune_all", "burnin", "mcmc", "gpu_pull", "gpu_after_move"):
        assert re.search(rf'^\s*{name} = function', rsrc, re.M), name
    # no hyper move may read a per-gene R field: those are stale between sampling points
    body = rsrc[rsrc.index("# ---- hyper-parameter moves"):rsrc.index("calculate_lnl = function")]
    for stale in ("Yg[", "variance_g[", "inactive_idx", "active_idx", "out_spike_idx", "in_spike_idx", "variance_g_probability", "XgLikelihood",
                  "allocation_within_active", "p_x"):
        code = "\n".join(l.split("#")[0] for l in body.splitlines())
        assert stale not in code, f"a hyper move reads the stale R field {stale}"
    for fld in ("gpu_hyper_ctr", "gpu_hyper_buf", "gpu_n_samples", "gpu_dirty", "gpu_sample_frequency"):
        assert re.search(rf'\b{fld} = "', rsrc), f"field {fld} is not declared"


def test_shim_reaches_the_c_abi():
    hdr = open(os.path.join(ROOT, "include", "zigzag_gpu.h")).read()
    exports = set(re.findall(r'^(?:int|void|const char \*|uint64_t|void \*)\s*\*?(zz_\w+)\(', hdr, re.M))
    src = open(SHIM).read()
    used = {e for e in exports if re.search(rf'\b{e}\(', src)}
    # not meaningful from R: raw CUDA stream / device pointers of another framework, diagnostics
    not_for_r = {"zz_set_stream", "zz_suffstats_dev", "zz_suffstats_allreduce_dev", "zz_set_stats_output_dev", "zz_debug_refresh_lps",
                 "zz_debug_gen_times", "zz_post_predictive_sharded"}      # (the last one takes a C callback for the caller's all-reduce)
    assert exports - used <= not_for_r, sorted(exports - used - not_for_r)
