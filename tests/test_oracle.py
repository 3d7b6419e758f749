"""Pin the CPU oracle (oracle/nfft_oracle.c) before anything is compared against it.

(1) the reference's known-answer fixtures (tests/data/nfft_*.txt -> golden/ndft_fixtures.npz),
    with the reference's own error measure and bound (tests/nfft.c:217-284, 739-781);
(2) the bessel_i0 table (tests/bessel.c) to the reference's 4*eps;
(3) outputs of the reference itself (golden/ref_outputs.npz, made from oracle/_ref), and the
    live reference build when oracle/_ref is present.
"""
import numpy as np
import pytest

import common
from common import REF_CASES, make_case, oracle, rel_l2

FIX = np.load(common.GOLDEN + "/ndft_fixtures.npz")
REFOUT = np.load(common.GOLDEN + "/ref_outputs.npz")
NAMES = sorted({k.split("/")[0] for k in FIX.files})


def next_power_of_2(x):  # kernel/util/int.c: 1 -> 2, else smallest 2^k >= x
    if x < 2:
        return x + 1
    return 1 << (int(x) - 1).bit_length()


def test_bessel_i0_table():
    tab = np.load(common.GOLDEN + "/bessel_i0.npz")["i0"]
    o = oracle("double")
    eps = np.finfo(np.float64).eps
    for j, ref in enumerate(tab):
        y = o.bessel_i0(float(j))
        assert abs(y - ref) <= 4 * eps * abs(ref), (j, y, ref)


@pytest.mark.parametrize("precision", ["double", "float"])
@pytest.mark.parametrize("name", NAMES)
def test_oracle_vs_ndft_fixture(name, precision):
    """nfft_init-style plan (n = 2*next_pow2(N), default m) as in tests/nfft.c:init_."""
    N = [int(v) for v in FIX[name + "/N"]]
    x, fh, f = FIX[name + "/x"], FIX[name + "/f_hat"], FIX[name + "/f"]
    if any(v % 2 for v in N) and not all(v <= 8 for v in N):
        pytest.skip("odd N: nfft_check rejects (tests/nfft.c:313-324 reports OK-skipped)")
    o = oracle(precision)
    m = 8 if precision == "double" else 4   # WINDOW_HELP_ESTIMATE_m, include/infft.h:224-230
    n = [2 * next_power_of_2(v) for v in N]
    bound = common.kb_error_bound(m, 2.0, precision)
    if str(FIX[name + "/kind"]) == "trafo":
        out = o.trafo(N, n, m, x, fh)
        err = np.max(np.abs(out - f)) / np.sum(np.abs(fh))
    else:
        out = o.adjoint(N, n, m, x, f)
        err = np.max(np.abs(out - fh)) / np.sum(np.abs(f))
    assert err < bound, (name, err, bound)


@pytest.mark.parametrize("precision", ["double", "float"])
@pytest.mark.parametrize("case", sorted(REF_CASES))
def test_oracle_vs_reference_golden(case, precision):
    spec = REF_CASES[case]
    x, fh, f = make_case(spec, precision)
    o = oracle(precision)
    tol = 2e-14 if precision == "double" else 1e-5
    got_f = o.trafo(spec["N"], spec["n"], spec["m"], x, fh)
    assert rel_l2(got_f, REFOUT[f"{case}/{precision}/f"]) <= tol
    got_fh = o.adjoint(spec["N"], spec["n"], spec["m"], x, f, sorted_order=True)
    assert rel_l2(got_fh, REFOUT[f"{case}/{precision}/f_hat"]) <= tol
    perm = o.sort_nodes(spec["n"], spec["m"], x)[:, 1]
    assert np.array_equal(perm, REFOUT[f"{case}/{precision}/index_x"])  # bit-exact permutation


@pytest.mark.skipif(not common.have_ref(), reason="oracle/_ref not built (no /root/reference)")
@pytest.mark.parametrize("precision", ["double", "float"])
def test_oracle_vs_live_reference(precision):
    """Fresh random inputs through the reference build and the restatement."""
    from nfft_b200.plan import Plan
    spec = dict(d=3, N=[12, 16, 10], n=[32, 32, 24], m=5, M=777, seed=4242,
                flags=REF_CASES["3d_N16_M400"]["flags"])
    x, fh, f = make_case(spec, precision)
    p = Plan.init_guru(spec["d"], spec["N"], spec["M"], spec["n"], spec["m"], spec["flags"],
                       api=common.ref_api(precision))
    p.x[:] = x
    p.f_hat[:] = fh
    p.trafo()
    ref_f = p.f.copy()
    p.f[:] = f
    p.adjoint()
    ref_fh = p.f_hat.copy()
    ref_perm = p.index_x[:, 1].copy()
    p.finalize()
    o = oracle(precision)
    tol = 2e-14 if precision == "double" else 1e-5
    assert rel_l2(o.trafo(spec["N"], spec["n"], spec["m"], x, fh), ref_f) <= tol
    assert rel_l2(o.adjoint(spec["N"], spec["n"], spec["m"], x, f, True), ref_fh) <= tol
    assert np.array_equal(o.sort_nodes(spec["n"], spec["m"], x)[:, 1], ref_perm)


@pytest.mark.parametrize("precision", ["double", "float"])
def test_oracle_direct_vs_fixture(precision):
    name = "nfft_2d_10_10_20"
    N = [int(v) for v in FIX[name + "/N"]]
    o = oracle(precision)
    out = o.trafo_direct(N, FIX[name + "/x"], FIX[name + "/f_hat"])
    eps = np.finfo(o.real).eps
    err = np.max(np.abs(out - FIX[name + "/f"])) / np.sum(np.abs(FIX[name + "/f_hat"]))
    assert err < 48 * eps   # tests/nfft.c:211-215
