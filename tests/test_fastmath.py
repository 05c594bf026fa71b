"""CPU test of the table-driven exp/log the production sweep uses (zz_fastmath.cuh is host+device code)."""
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fast_exp_log_within_two_ulp(tmp_path):
    exe = str(tmp_path / "fm")
    cc = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cc, "-O2", "-ffp-contract=off", "-I", os.path.join(ROOT, "zigzag_b200", "csrc"), "-o", exe,
                           os.path.join(ROOT, "tests", "fastmath_host.cpp")])
    out = dict(line.split(None, 1) for line in subprocess.check_output([exe, "3000000"], text=True).splitlines())
    assert float(out["exp_max_ulp"]) < 2.5            # table + degree-4 economised polynomial
    assert float(out["log_max_ulp"]) < 2.5
    assert float(out["log_max_abs_near_one"]) < 2e-17  # no cancellation blow-up around x = 1
    assert float(out["exp2_max_ulp"]) < 2.5           # base-2 detection exponential
    assert float(out["px_max_abs_err"]) < 2.3e-16      # 1 - 2^-b vs 1 - exp(-a): within one ulp of 1
    assert float(out["det_product_worst"]) < 1.0       # log of the product of the detection factors == sum of the per-library logs
    assert float(out["det_q_worst"]) <= 2.0            # each factor within 2^-52 of the per-library exp2 path (1 ulp of e, re-rounded by 1 - p)
    assert int(out["det_product_zero_bad"]) == 0       # a zero factor gives exactly -Inf, nothing else does
    assert float(out["log1"]) == 0.0
    assert out["log0"].strip() == "-inf" and "nan" in out["logneg"] and out["loginf"].strip() == "inf"
    a, b = map(float, out["logsub"].split())
    assert a == b                                      # subnormals take the libm path
    assert float(out["exp_m800"]) == 0.0 and out["exp_800"].strip() == "inf" and "nan" in out["exp_nan"]
    a, b = map(float, out["exp_m720"].split())
    assert a == b
    assert int(out["px_mismatch"]) == 0


def test_tables_are_reproducible():
    """zz_tables.h is exactly what tools/gen_tables.py emits."""
    path = os.path.join(ROOT, "zigzag_b200", "csrc", "zz_tables.h")
    before = open(path).read()
    subprocess.check_call(["python", os.path.join(ROOT, "tools", "gen_tables.py")], stdout=subprocess.DEVNULL)
    assert open(path).read() == before
