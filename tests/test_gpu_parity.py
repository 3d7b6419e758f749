"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through the drop-in
boundary: the reference-facing plan API of libnfft3_b200.so (nfft_b200.plan.Plan) or the C ABI of
libnfftcu.so (nfft_b200.cabi.Engine); the CPU oracle / reference build are only the checker.

Tolerances: fp64 rel-l2 <= 1e-12, fp32 rel-l2 <= 1e-5 against the reference's own NFFT output
(BASELINE.json north_star); sort permutation bit-exact; against the exact-NDFT fixtures the
reference's own bound (tests/nfft.c:217-284).
"""
import ctypes as C
import os

import numpy as np
import pytest

import common
from common import REF_CASES, make_case, oracle, rel_l2
from nfft_b200 import cabi, plan_abi as abi
from nfft_b200.plan import Plan

pytestmark = pytest.mark.gpu

TOL = {"double": 1e-12, "float": 1e-5}
FIX = np.load(common.GOLDEN + "/ndft_fixtures.npz")
REFOUT = np.load(common.GOLDEN + "/ref_outputs.npz")
NAMES = sorted({k.split("/")[0] for k in FIX.files})
BASE = abi.PRE_PHI_HUT | abi.MALLOC_X | abi.MALLOC_F_HAT | abi.MALLOC_F | abi.FFTW_INIT | abi.FFT_OUT_OF_PLACE


def next_power_of_2(x):
    return x + 1 if x < 2 else 1 << (int(x) - 1).bit_length()


def run_plan(spec, precision, x, fh, f, flags=None):
    """init_guru -> x -> precompute -> trafo -> adjoint, like tests/nfft.c:check_single."""
    p = Plan.init_guru(spec["d"], spec["N"], spec["M"], spec["n"], spec["m"],
                       spec["flags"] if flags is None else flags, precision=precision)
    p.x[:] = x
    if p.flags & abi.PRE_ONE_PSI:
        p.precompute_one_psi()
    p.f_hat[:] = fh
    p.trafo()
    out_f = p.f.copy()
    p.f[:] = f
    p.adjoint()
    out_fh = p.f_hat.copy()
    perm = p.index_x[:, 1].copy() if (p.flags & abi.NFFT_SORT_NODES) else None
    p.finalize()
    return out_f, out_fh, perm


# ---- (1) the reference's known-answer fixtures, with the reference's initialiser matrix -------------
@pytest.mark.parametrize("precision", ["double", "float"])
@pytest.mark.parametrize("initializer", ["init", "init_nd", "guru_pre_psi", "guru_pre_full_psi", "guru_no_psi"])
@pytest.mark.parametrize("name", NAMES)
def test_ndft_fixture(name, initializer, precision):
    N = [int(v) for v in FIX[name + "/N"]]
    x, fh, f = FIX[name + "/x"], FIX[name + "/f_hat"], FIX[name + "/f"]
    d, M = len(N), x.shape[0]
    m = 8 if precision == "double" else 4
    n = [2 * next_power_of_2(v) for v in N]
    if initializer == "init":
        p = Plan.init(d, N, M, precision=precision)
    elif initializer == "init_nd":
        p = Plan.init_nd(N, M, precision=precision)
    else:
        extra = {"guru_pre_psi": abi.PRE_PSI, "guru_pre_full_psi": abi.PRE_FULL_PSI,
                 "guru_no_psi": abi.NFFT_SORT_NODES}[initializer]
        p = Plan.init_guru(d, N, M, n, m, BASE | extra, precision=precision)
    p.x[:] = x
    if p.check() is not None:       # odd N etc.: the reference reports OK-skipped (tests/nfft.c:313-324)
        direct_ok = all(v <= p.m for v in N)
        if not direct_ok:
            p.finalize()
            pytest.skip(p.check() or "nfft_check")
    if p.flags & abi.PRE_ONE_PSI:
        p.precompute_one_psi()
    bound = common.kb_error_bound(p.m, 2.0, precision)
    if str(FIX[name + "/kind"]) == "trafo":
        p.f_hat[:] = fh
        p.trafo_nd() if initializer == "init_nd" else p.trafo()
        err = np.max(np.abs(p.f - f)) / np.sum(np.abs(fh))
    else:
        p.f[:] = f
        p.adjoint_nd() if initializer == "init_nd" else p.adjoint()
        err = np.max(np.abs(p.f_hat - fh)) / np.sum(np.abs(f))
    p.finalize()
    assert err < bound, (name, initializer, err, bound)


@pytest.mark.parametrize("precision", ["double", "float"])
@pytest.mark.parametrize("name", ["nfft_1d_50_50", "nfft_2d_20_20_50", "nfft_3d_10_10_10_10",
                                  "nfft_adjoint_1d_20_50", "nfft_adjoint_2d_10_20_20",
                                  "nfft_adjoint_3d_10_10_10_10"])
def test_direct_fixture(name, precision):
    """nfft_trafo_direct / nfft_adjoint_direct to 48 eps (tests/nfft.c:211-215)."""
    N = [int(v) for v in FIX[name + "/N"]]
    x, fh, f = FIX[name + "/x"], FIX[name + "/f_hat"], FIX[name + "/f"]
    p = Plan.init(len(N), N, x.shape[0], precision=precision)
    p.x[:] = x
    eps = np.finfo(p.api.real).eps
    if str(FIX[name + "/kind"]) == "trafo":
        p.f_hat[:] = fh
        p.trafo_direct()
        err = np.max(np.abs(p.f - f)) / np.sum(np.abs(fh))
    else:
        p.f[:] = f
        p.adjoint_direct()
        err = np.max(np.abs(p.f_hat - fh)) / np.sum(np.abs(f))
    p.finalize()
    assert err < 48 * eps


# ---- (2) outputs of the reference itself (committed golden vectors) -----------------------------------
@pytest.mark.parametrize("precision", ["double", "float"])
@pytest.mark.parametrize("case", sorted(REF_CASES))
def test_vs_reference_golden(case, precision):
    spec = REF_CASES[case]
    x, fh, f = make_case(spec, precision)
    out_f, out_fh, perm = run_plan(spec, precision, x, fh, f)
    assert rel_l2(out_f, REFOUT[f"{case}/{precision}/f"]) <= TOL[precision]
    assert rel_l2(out_fh, REFOUT[f"{case}/{precision}/f_hat"]) <= TOL[precision]
    assert np.array_equal(perm, REFOUT[f"{case}/{precision}/index_x"])


@pytest.mark.skipif(not common.have_ref(), reason="oracle/_ref not present")
@pytest.mark.parametrize("precision", ["double", "float"])
def test_vs_live_reference_build(precision):
    """Same plan run through the unmodified reference (oracle/_ref) and the product."""
    flags = BASE | abi.NFFT_SORT_NODES | abi.NFFT_OMP_BLOCKWISE_ADJOINT
    spec = dict(d=3, N=[32, 32, 32], n=[64, 64, 64], m=6, M=30000, seed=99, flags=flags)
    x, fh, f = make_case(spec, precision)
    ref = Plan.init_guru(3, spec["N"], spec["M"], spec["n"], 6, flags, api=common.ref_api(precision))
    ref.x[:] = x
    ref.f_hat[:] = fh
    ref.trafo()
    ref_f = ref.f.copy()
    ref.f[:] = f
    ref.adjoint()
    ref_fh, ref_perm = ref.f_hat.copy(), ref.index_x.copy()
    ref.finalize()
    out_f, out_fh, _ = run_plan(spec, precision, x, fh, f)
    assert rel_l2(out_f, ref_f) <= TOL[precision]
    assert rel_l2(out_fh, ref_fh) <= TOL[precision]
    eng = cabi.Engine(spec["N"], spec["n"], 6, spec["M"], precision=precision)
    eng.set_nodes(x)
    assert np.array_equal(eng.index_x(), ref_perm)   # keys and permutation, bit-exact
    eng.close()


# ---- (3) the oracle on fresh seeded inputs, sizes the oracle finishes in seconds ----------------------
ORACLE_CASES = {
    "cfg1_1d": dict(d=1, N=[1024], n=[2048], m=6, M=10000, seed=20260102, flags=BASE | abi.PRE_PSI),
    "2d_128": dict(d=2, N=[128, 128], n=[256, 256], m=6, M=40000, seed=21, flags=BASE | abi.PRE_PSI),
    "2d_m8_rect": dict(d=2, N=[64, 96], n=[128, 256], m=8, M=20000, seed=22, flags=BASE),
    "3d_32": dict(d=3, N=[32, 32, 32], n=[64, 64, 64], m=6, M=50000, seed=23,
                  flags=BASE | abi.NFFT_SORT_NODES | abi.NFFT_OMP_BLOCKWISE_ADJOINT),
    "3d_m2": dict(d=3, N=[16, 24, 32], n=[32, 48, 64], m=2, M=5000, seed=24, flags=BASE),
    "3d_m9": dict(d=3, N=[24, 24, 24], n=[48, 48, 48], m=9, M=3000, seed=25, flags=BASE),
    "4d": dict(d=4, N=[8, 8, 8, 8], n=[16, 16, 16, 16], m=2, M=1000, seed=26, flags=BASE),
    "1d_sigma_big": dict(d=1, N=[100], n=[512], m=5, M=3000, seed=27, flags=BASE),
}


@pytest.mark.parametrize("precision", ["double", "float"])
@pytest.mark.parametrize("case", sorted(ORACLE_CASES))
def test_vs_oracle(case, precision):
    spec = ORACLE_CASES[case]
    x, fh, f = make_case(spec, precision)
    o = oracle(precision)
    want_f = o.trafo(spec["N"], spec["n"], spec["m"], x, fh)
    if not np.all(np.isfinite(want_f)):
        # m = 9 in 3-D: psi0*psi1*psi2 ~ (1e17)^3 overflows float in the reference's arithmetic too
        pytest.skip("window product overflows single precision")
    out_f, out_fh, _ = run_plan(spec, precision, x, fh, f)
    assert rel_l2(out_f, want_f) <= TOL[precision]
    assert rel_l2(out_fh, o.adjoint(spec["N"], spec["n"], spec["m"], x, f, True)) <= TOL[precision]


# ---- (4) single stages through the C ABI ---------------------------------------------------------------
@pytest.mark.parametrize("precision", ["double", "float"])
@pytest.mark.parametrize("N,n,m", [([16, 16, 16], [32, 32, 32], 6), ([20, 30], [50, 72], 5),
                                   ([64], [256], 4), ([6, 10, 12], [16, 24, 28], 2)])
def test_stages(N, n, m, precision):
    rng = np.random.default_rng(5)
    d, M = len(N), 2000
    o = oracle(precision)
    x = (rng.random((M, d)) - 0.5).astype(o.real)
    NN, nn = int(np.prod(N)), int(np.prod(n))
    fh = (rng.random(NN) + 1j * rng.random(NN)).astype(o.cplx)
    f = (rng.random(M) + 1j * rng.random(M)).astype(o.cplx)
    g_in = (rng.random(nn) - 0.5 + 1j * (rng.random(nn) - 0.5)).astype(o.cplx)
    tol = 1e-13 if precision == "double" else 5e-6
    eng = cabi.Engine(N, n, m, M, precision=precision)
    eng.set_nodes(x)
    assert np.allclose(eng.c_phi_inv(0), o.c_phi_inv(N[0], n[0], m), rtol=4e-16 if precision == "double" else 3e-7)
    fh_d = cabi.DeviceBuffer(fh.nbytes).upload(fh)
    f_d = cabi.DeviceBuffer(f.nbytes).upload(f)
    # grid units of the single-stage calls: the device window carries an exact power-of-two factor S (1 in fp64)
    S = eng.window_scale()
    assert S == 1.0 or precision == "float"
    # D
    eng.stage_D(fh_d)
    assert rel_l2(eng.grid_to_host().astype(np.complex128) * S, o.stage_D(N, n, m, fh)) <= tol
    # F, both signs, on random data
    for sign in (-1, +1):
        eng.grid_from_host(g_in)
        eng.stage_F(sign)
        assert rel_l2(eng.grid_to_host(), o.stage_F(n, sign, g_in)) <= (1e-14 if precision == "double" else 2e-6)
    # B
    eng.grid_from_host(g_in)
    eng.stage_B(f_d)
    eng.sync()
    assert rel_l2(f_d.download(o.cplx, M).astype(np.complex128) / S, o.stage_B(N, n, m, x, g_in)) <= tol * 10
    # B^T
    f_d.upload(f)
    eng.stage_BT(f_d)
    assert rel_l2(eng.grid_to_host().astype(np.complex128) / S, o.stage_BT(N, n, m, x, f)) <= tol * 10
    # D^T
    eng.grid_from_host(g_in)
    eng.stage_DT(fh_d)
    eng.sync()
    assert rel_l2(fh_d.download(o.cplx, NN).astype(np.complex128) * S, o.stage_DT(N, n, m, g_in)) <= tol
    eng.close()


# ---- (5) node sort: bit-exact permutation --------------------------------------------------------------
@pytest.mark.parametrize("precision", ["double", "float"])
@pytest.mark.parametrize("n,m,M,mode", [([256, 256, 256], 6, 300000, "uniform"),
                                        ([64, 64], 4, 100000, "dups"),
                                        ([50, 72, 40], 5, 50000, "uniform"),
                                        ([2048], 6, 70000, "dups"),
                                        ([512, 512, 512], 6, 20000, "edges")])
def test_sort_bit_exact(n, m, M, mode, precision):
    rng = np.random.default_rng(77)
    d = len(n)
    o = oracle(precision)
    x = rng.random((M, d)) - 0.5
    if mode == "dups":      # heavy key duplication: stability decides the order
        x = np.round(x * 16) / 16
        x[x >= 0.5] = -0.5
    if mode == "edges":     # nodes on and next to cell boundaries, +-0.5
        x = np.round(x * n[0]) / n[0] + rng.choice([0.0, 1e-9, -1e-9], size=(M, d))
        x = np.clip(x, -0.5, np.nextafter(0.5, 0))
    x = x.astype(o.real)
    if precision == "float":
        x = np.minimum(x, np.nextafter(np.float32(0.5), np.float32(0)))
    N = [v // 2 for v in n]
    eng = cabi.Engine(N, n, m, M, precision=precision)
    eng.set_nodes(x)
    got = eng.index_x()
    eng.close()
    want = o.sort_nodes(n, m, x)
    assert np.array_equal(got[:, 0], want[:, 0])
    assert np.array_equal(got[:, 1], want[:, 1])


# ---- (6) edge cases ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["double", "float"])
def test_edge_nodes_and_sizes(precision):
    o = oracle(precision)
    N, n, m = [16, 16, 16], [32, 32, 32], 6
    rng = np.random.default_rng(3)
    NN = 16 ** 3
    fh = (rng.random(NN) + 1j * rng.random(NN)).astype(o.cplx)
    lo, hi = -0.5, np.nextafter(o.real(0.5), o.real(0))
    cases = {
        "single": np.array([[0.1, -0.2, 0.3]]),
        "corners": np.array([[lo, lo, lo], [hi, hi, hi], [lo, hi, 0.0], [0.0, 0.0, 0.0], [hi, lo, hi]]),
        "identical": np.tile(np.array([[0.25, -0.125, 0.0625]]), (257, 1)),
    }
    for name, x in cases.items():
        x = x.astype(o.real)
        M = x.shape[0]
        f = (rng.random(M) + 1j * rng.random(M)).astype(o.cplx)
        spec = dict(d=3, N=N, n=n, m=m, M=M, flags=BASE | abi.NFFT_SORT_NODES)
        out_f, out_fh, perm = run_plan(spec, precision, x, fh, f)
        assert rel_l2(out_f, o.trafo(N, n, m, x, fh)) <= TOL[precision], name
        assert rel_l2(out_fh, o.adjoint(N, n, m, x, f, True)) <= TOL[precision], name
        assert np.array_equal(perm, o.sort_nodes(n, m, x)[:, 1]), name


def test_empty_node_set():
    p = Plan.init_guru(2, [16, 16], 0, [32, 32], 4, abi.PRE_PHI_HUT | abi.MALLOC_F_HAT | abi.FFTW_INIT)
    dummy = np.zeros(2, dtype=np.float64)
    p.c.x = dummy.ctypes.data_as(C.POINTER(C.c_double))
    p.c.f = dummy.ctypes.data_as(C.POINTER(C.c_double))
    p.f_hat[:] = 1.0
    p.adjoint()          # g stays zero -> f_hat == 0 (nfft.c:5137 memset, no nodes)
    assert np.all(p.f_hat == 0)
    p.trafo()            # nothing to write
    p.finalize()


def test_small_N_falls_back_to_direct():
    """any N_t <= m -> exact NDFT (nfft.c:5658-5664): result must equal trafo_direct's."""
    o = oracle("double")
    rng = np.random.default_rng(8)
    N, n, m, M = [4, 32], [8, 64], 6, 100
    x = rng.random((M, 2)) - 0.5
    fh = rng.random(128) + 1j * rng.random(128)
    p = Plan.init_guru(2, N, M, n, m, BASE)
    p.x[:] = x
    p.f_hat[:] = fh
    p.trafo()
    assert rel_l2(p.f, o.trafo_direct(N, x, fh)) <= 1e-14
    p.f[:] = fh[:M]
    p.adjoint()
    assert rel_l2(p.f_hat, o.adjoint_direct(N, x, fh[:M])) <= 1e-14
    p.finalize()


# ---- (7) host-pointer semantics the callers rely on ---------------------------------------------------------
def test_pointer_swap_and_node_refresh():
    """solver.c swaps f/f_hat around each call (kernel/solver/solver.c:240-242,275-277); plans
    without a psi flag see new nodes on the next call without notification (nfft.c:4889)."""
    o = oracle("double")
    rng = np.random.default_rng(9)
    N, n, m, M = [32, 32], [64, 64], 6, 5000
    p = Plan.init_guru(2, N, M, n, m, BASE | abi.NFFT_SORT_NODES)
    x1, x2 = rng.random((M, 2)) - 0.5, rng.random((M, 2)) - 0.5
    fh = rng.random(1024) + 1j * rng.random(1024)
    other_f = np.zeros(M, dtype=np.complex128)
    p.x[:] = x1
    p.f_hat[:] = fh
    own_f = C.cast(p.c.f, C.c_void_p).value                    # the address: a ctypes field read aliases the field
    p.c.f = other_f.ctypes.data_as(C.POINTER(C.c_double))      # CSWAP
    p.trafo()
    p.c.f = C.cast(C.c_void_p(own_f), C.POINTER(C.c_double))
    assert rel_l2(other_f, o.trafo(N, n, m, x1, fh)) <= 1e-12
    v1 = np.array(p.index_x[:, 1])
    p.x[:] = x2                                                 # silent node change
    p.trafo()
    assert rel_l2(p.f, o.trafo(N, n, m, x2, fh)) <= 1e-12
    assert np.array_equal(p.index_x[:, 1], o.sort_nodes(n, m, x2)[:, 1])
    assert not np.array_equal(v1, p.index_x[:, 1])
    p.finalize()


# ---- (8) size-independent properties at the benchmark shape ---------------------------------------------------
@pytest.mark.parametrize("precision,M", [("double", 2_000_000), ("float", 2_000_000)])
def test_adjointness_and_linearity_cfg3_grid(precision, M):
    """3-D N=128^3, n=256^3, m=6 (BASELINE configs[2] grid): <A u, v> == <u, A^H v> and
    A(u1 + 2 u2) == A u1 + 2 A u2.  No oracle needed, so the node count can be large."""
    rng = np.random.default_rng(1234)
    N, n, m = [128] * 3, [256] * 3, 6
    real = np.float64 if precision == "double" else np.float32
    cplx = np.complex128 if precision == "double" else np.complex64
    x = (rng.random((M, 3)) - 0.5).astype(real)
    if precision == "float":
        x = np.minimum(x, np.nextafter(np.float32(0.5), np.float32(0)))
    NN = 128 ** 3
    u1 = (rng.random(NN) - 0.5 + 1j * (rng.random(NN) - 0.5)).astype(cplx)
    u2 = (rng.random(NN) - 0.5 + 1j * (rng.random(NN) - 0.5)).astype(cplx)
    v = (rng.random(M) - 0.5 + 1j * (rng.random(M) - 0.5)).astype(cplx)
    eng = cabi.Engine(N, n, m, M, precision=precision)
    eng.set_nodes(x)
    Au1, Au2 = eng.trafo(u1), eng.trafo(u2)
    Au12 = eng.trafo((u1 + 2 * u2).astype(cplx))
    AHv = eng.adjoint(v)
    eng.close()
    tol = 1e-12 if precision == "double" else 2e-5
    assert rel_l2(Au12, Au1.astype(np.complex128) + 2 * Au2.astype(np.complex128)) <= tol
    lhs = np.vdot(v.astype(np.complex128), Au1.astype(np.complex128))
    rhs = np.vdot(AHv.astype(np.complex128), u1.astype(np.complex128))
    scale = np.linalg.norm(v.astype(np.complex128)) * np.linalg.norm(Au1.astype(np.complex128))
    assert abs(lhs - rhs) / scale <= tol


def test_trafo_subsample_vs_oracle_cfg3_grid():
    """trafo outputs are per-node independent: check a 3000-node subsample of a 10^6-node run on
    the N=128^3 grid against the oracle's B step fed with the device grid."""
    rng = np.random.default_rng(4321)
    N, n, m, M = [128] * 3, [256] * 3, 6, 1_000_000
    x = rng.random((M, 3)) - 0.5
    NN = 128 ** 3
    fh = rng.random(NN) - 0.5 + 1j * (rng.random(NN) - 0.5)
    eng = cabi.Engine(N, n, m, M, precision="double")
    eng.set_nodes(x)
    f = eng.trafo(fh)
    g = eng.grid_to_host()
    eng.close()
    sel = rng.choice(M, 3000, replace=False)
    want = oracle("double").stage_B(N, n, m, x[sel], g)
    assert rel_l2(f[sel], want) <= 1e-12


# ---- (9) 3-D fast path (tile3d.cu) against the generic kernels and the oracle ----------------------------------
@pytest.mark.parametrize("precision", ["double", "float"])
@pytest.mark.parametrize("N,n,m,M", [
    ([16, 16, 16], [32, 32, 32], 6, 4000),       # power of two
    ([24, 16, 20], [50, 33, 41], 5, 3000),       # n not divisible by the tile / slab sizes, odd n2
    ([8, 8, 8], [16, 16, 16], 6, 500),           # footprint (16) == n: rows wrap onto the whole axis
    ([8, 10, 12], [15, 21, 26], 6, 700),         # footprint larger than n0: footprint rows alias
    ([32, 32, 32], [64, 64, 64], 2, 6000),
    ([20, 20, 20], [40, 40, 40], 8, 2000),
    ([64, 8, 8], [128, 16, 16], 4, 3000),
    ([20, 18, 21], [40, 36, 42], 6, 5000),       # DMMA path with n2 % 8 != 0 (RED flush), n % 16 != 0
    ([12, 12, 12], [24, 24, 24], 6, 20000),      # DMMA path, many nodes per tile, window wraps in z
    ([16, 16, 16], [32, 32, 40], 3, 4000),       # DMMA path, tile edge 9
])
def test_tile3d_vs_generic_and_oracle(N, n, m, M, precision):
    rng = np.random.default_rng(31)
    o = oracle(precision)
    x = (rng.random((M, 3)) - 0.5).astype(o.real)
    x[: M // 8] = np.round(x[: M // 8] * 4) / 4          # clustered + exact cell-boundary nodes
    x = np.clip(x, -0.5, np.nextafter(o.real(0.5), o.real(0)))
    NN = int(np.prod(N))
    fh = (rng.random(NN) - 0.5 + 1j * (rng.random(NN) - 0.5)).astype(o.cplx)
    f = (rng.random(M) - 0.5 + 1j * (rng.random(M) - 0.5)).astype(o.cplx)
    want_f = o.trafo(N, n, m, x, fh)
    want_fh = o.adjoint(N, n, m, x, f)
    if not (np.all(np.isfinite(want_f)) and np.all(np.isfinite(want_fh))):
        pytest.skip("psi0*psi1*psi2 overflows single precision at this m (in the reference too)")
    outs = {}
    variants = [("tile", 0, 0, 0), ("pencil", 2, 0, 0), ("pencil+table", 2, 1, 0), ("generic", 1, 0, 0),
                ("generic+table", 1, 1, 0)]
    if precision == "double":
        variants += [("mma", 3, 0, 0), ("mma+red", 3, 0, 1), ("mma+tma", 3, 0, 2)]
    for label, kernel, table, flush in variants:
        eng = cabi.Engine(N, n, m, M, precision=precision)
        eng.set_option(cabi.OPT_B_KERNEL, kernel)
        eng.set_option(cabi.OPT_PSI_TABLE, table)
        eng.set_option(cabi.OPT_B_FLUSH, flush)
        eng.set_nodes(x)
        outs[label] = (eng.trafo(fh), eng.adjoint(f))
        eng.close()
        assert rel_l2(outs[label][0], want_f) <= TOL[precision], label
        assert rel_l2(outs[label][1], want_fh) <= TOL[precision], label
    tight = 1e-13 if precision == "double" else 5e-6
    for label in outs:
        assert rel_l2(outs[label][0], outs["generic"][0]) <= tight, label
        assert rel_l2(outs[label][1], outs["generic"][1]) <= tight, label


@pytest.mark.parametrize("precision", ["double", "float"])
@pytest.mark.parametrize("N,n,m,M", [
    ([64], [128], 6, 500), ([20], [50], 4, 300),
    ([32, 16], [64, 32], 6, 2000), ([20, 30], [50, 72], 5, 1500),
    ([16, 16, 16], [32, 32, 32], 6, 3000), ([12, 10, 14], [30, 24, 36], 4, 2500), ([8, 6], [16, 18], 3, 200),
])
def test_pruned_fft_matches_full_passes(N, n, m, M, precision):
    """trafo/adjoint with the band-pruned FFT passes and the D that skips the zero padding (default) against
    full passes over a zero-padded grid, and both against the oracle."""
    rng = np.random.default_rng(77)
    o = oracle(precision)
    d = len(N)
    x = (rng.random((M, d)) - 0.5).astype(o.real)
    NN = int(np.prod(N))
    fh = (rng.random(NN) - 0.5 + 1j * (rng.random(NN) - 0.5)).astype(o.cplx)
    f = (rng.random(M) - 0.5 + 1j * (rng.random(M) - 0.5)).astype(o.cplx)
    outs = []
    for prune in (1, 0):
        eng = cabi.Engine(N, n, m, M, precision=precision)
        eng.set_option(cabi.OPT_FFT_PRUNE, prune)
        eng.set_nodes(x)
        outs.append((eng.trafo(fh), eng.adjoint(f), eng.trafo(fh)))   # trafo again: the grid holds adjoint leftovers
        eng.close()
    assert rel_l2(outs[0][0], o.trafo(N, n, m, x, fh)) <= TOL[precision]
    assert rel_l2(outs[0][1], o.adjoint(N, n, m, x, f)) <= TOL[precision]
    tight = 1e-14 if precision == "double" else 1e-6
    for a, b in zip(outs[0], outs[1]):
        assert rel_l2(a, b) <= tight
    assert np.array_equal(outs[0][0], outs[0][2])


@pytest.mark.parametrize("precision", ["double", "float"])
@pytest.mark.parametrize("prune", [1, 0])
@pytest.mark.parametrize("N,n,m,M", [
    ([32], [64], 4, 300), ([64], [128], 6, 500), ([100], [256], 6, 500), ([256], [512], 6, 800),
    ([512], [1024], 6, 800), ([1024], [2048], 6, 3000),                       # 1-D, contiguous lines: every radix split
    ([32, 64], [64, 128], 5, 3000), ([128, 30], [256, 64], 4, 3000), ([256, 512], [512, 1024], 4, 8000),
    ([60, 500], [128, 1024], 3, 4000), ([1000, 20], [2048, 64], 3, 4000),     # strided axis, partial bundles (n1 = 64 ...)
    ([32, 32, 32], [64, 64, 64], 6, 20000), ([20, 64, 128], [64, 128, 256], 4, 20000),
    ([128, 128, 128], [256, 256, 256], 6, 50000),                             # the benchmark grid
])
def test_register_fft_vs_shared_memory_fft(N, n, m, M, prune, precision):
    """F with the register-resident Stockham kernel (fft_reg_kernel, default for 2^k lengths 64..2048) against
    the shared-memory radix-4 Stockham (fft_stockham_kernel, NFFTCU_OPT_FFT_KERNEL = 1), pruned and full passes,
    both transform directions; the small cases also against the oracle."""
    rng = np.random.default_rng(78)
    o = oracle(precision)
    d = len(N)
    x = (rng.random((M, d)) - 0.5).astype(o.real)
    NN = int(np.prod(N))
    fh = (rng.random(NN) - 0.5 + 1j * (rng.random(NN) - 0.5)).astype(o.cplx)
    f = (rng.random(M) - 0.5 + 1j * (rng.random(M) - 0.5)).astype(o.cplx)
    outs = []
    for kernel in (0, 1):
        eng = cabi.Engine(N, n, m, M, precision=precision)
        eng.set_option(cabi.OPT_FFT_PRUNE, prune)
        eng.set_option(cabi.OPT_FFT_KERNEL, kernel)
        eng.set_nodes(x)
        outs.append((eng.trafo(fh), eng.adjoint(f)))
        eng.close()
    tight = 1e-14 if precision == "double" else 2e-6
    assert rel_l2(outs[0][0], outs[1][0]) <= tight
    assert rel_l2(outs[0][1], outs[1][1]) <= tight
    if int(np.prod(n)) <= 1 << 21:
        assert rel_l2(outs[0][0], o.trafo(N, n, m, x, fh)) <= TOL[precision]
        assert rel_l2(outs[0][1], o.adjoint(N, n, m, x, f)) <= TOL[precision]


@pytest.mark.parametrize("precision", ["double", "float"])
def test_tile3d_clustered_nodes_are_cut_into_chunks(precision):
    """2*10^5 nodes in a ball of radius 0.04 on a 64^3 grid: a handful of tiles hold thousands of batches each, so
    their work units are cut into many chunks (mma_chunk_fill_kernel); every chunk preloads / retires its own window."""
    rng = np.random.default_rng(33)
    o = oracle(precision)
    N, n, m, M = [32, 32, 32], [64, 64, 64], 6, 200_000
    v = rng.normal(size=(M, 3))
    x = (v / np.linalg.norm(v, axis=1, keepdims=True) * (0.04 * rng.random((M, 1)) ** (1 / 3)) + 0.21).astype(o.real)
    fh = (rng.random(32 ** 3) - 0.5 + 1j * (rng.random(32 ** 3) - 0.5)).astype(o.cplx)
    f = (rng.random(M) - 0.5 + 1j * (rng.random(M) - 0.5)).astype(o.cplx)
    eng = cabi.Engine(N, n, m, M, precision=precision)
    eng.set_nodes(x)
    got_f, got_fh = eng.trafo(fh), eng.adjoint(f)
    eng.close()
    assert rel_l2(got_f, o.trafo(N, n, m, x, fh)) <= TOL[precision]
    if precision == "double":
        assert rel_l2(got_fh, o.adjoint(N, n, m, x, f)) <= TOL[precision]
    else:
        # every grid cell of the cluster sums ~10^5 fp32 contributions: two fp32 summation orders (the reference's
        # and ours) differ by ~sqrt(10^5) eps = 2e-5 here, so the fp32 result is held against the fp64 oracle instead
        od = oracle("double")
        want = od.adjoint(N, n, m, x.astype(np.float64), f.astype(np.complex128))
        assert rel_l2(got_fh, want) <= 5e-5
        assert rel_l2(got_fh, want) <= 2 * rel_l2(o.adjoint(N, n, m, x, f), want) + 1e-5   # no worse than the fp32 reference


@pytest.mark.parametrize("precision", ["double", "float"])
@pytest.mark.parametrize("N,n,m,M", [([16, 16, 16], [32, 32, 32], 6, 20000), ([20, 18, 21], [40, 36, 42], 4, 5000),
                                     ([12, 12, 12], [24, 24, 24], 2, 3000)])
def test_tile3d_window_images_vs_evaluating_producers(N, n, m, M, precision):
    """3-D tensor kernels fed by plan-time window images (TMA bulk copy per batch, NFFTCU_OPT_WINDOW_IMAGES auto)
    against the same kernels with evaluating producer warps (option 1), and both against the oracle."""
    rng = np.random.default_rng(35)
    o = oracle(precision)
    x = (rng.random((M, 3)) - 0.5).astype(o.real)
    NN = int(np.prod(N))
    fh = (rng.random(NN) - 0.5 + 1j * (rng.random(NN) - 0.5)).astype(o.cplx)
    f = (rng.random(M) - 0.5 + 1j * (rng.random(M) - 0.5)).astype(o.cplx)
    outs = []
    for images in (0, 1, 2):
        eng = cabi.Engine(N, n, m, M, precision=precision)
        eng.set_option(cabi.OPT_WINDOW_IMAGES, images)
        eng.set_nodes(x)
        outs.append((eng.trafo(fh), eng.adjoint(f)))
        eng.close()
    assert rel_l2(outs[0][0], o.trafo(N, n, m, x, fh)) <= TOL[precision]
    assert rel_l2(outs[0][1], o.adjoint(N, n, m, x, f)) <= TOL[precision]
    tight = 1e-14 if precision == "double" else 2e-6
    for k in (1, 2):
        assert rel_l2(outs[k][0], outs[0][0]) <= tight
        assert rel_l2(outs[k][1], outs[0][1]) <= tight


def test_tile3d_z_segments_small_grid_many_nodes():
    """few tiles -> the sweep is split into z segments; every segment flushes / preloads its window."""
    rng = np.random.default_rng(32)
    N, n, m, M = [12, 12, 64], [24, 24, 128], 6, 60000
    o = oracle("double")
    x = rng.random((M, 3)) - 0.5
    fh = rng.random(12 * 12 * 64) + 1j * rng.random(12 * 12 * 64)
    f = rng.random(M) + 1j * rng.random(M)
    eng = cabi.Engine(N, n, m, M)
    eng.set_nodes(x)
    assert rel_l2(eng.trafo(fh), o.trafo(N, n, m, x, fh)) <= 1e-12
    assert rel_l2(eng.adjoint(f), o.adjoint(N, n, m, x, f)) <= 1e-12
    eng.close()


# ---- (10) the reference's own solver, unmodified, on top of the engine ------------------------------------
@pytest.mark.parametrize("d,N,n,M,solver", [
    (2, [32, 32], [64, 64], 3000, "cgnr_damp"),
    (2, [32, 32], [64, 64], 800, "cgne_weight"),
    (3, [16, 16, 16], [32, 32, 32], 6000, "cgnr_damp"),      # d = 3, m = 6: the DMMA kernels
    (1, [128], [256], 400, "cgnr"),
])
def test_reference_solver_runs_on_the_engine(d, N, n, M, solver):
    """kernel/solver/solver.c (CGNR 232-296, CGNE 298-344) compiled from the reference tree, unmodified, and
    linked against libnfft3_b200.so (oracle/refbuild: libsolver_b200.so) against the same driver on the
    reference's own nfft.c (libsolver_ref.so): 12 iterations, iterate and residual norms must agree."""
    ref_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
    so_ref, so_b200 = os.path.join(ref_dir, "libsolver_ref.so"), os.path.join(ref_dir, "libsolver_b200.so")
    if not (os.path.exists(so_ref) and os.path.exists(so_b200)):
        pytest.skip("oracle/_ref/libsolver_*.so not built (needs /root/reference at build time)")
    LANDWEBER, STEEPEST, CGNR, CGNE, NORMS, PRE_W, PRE_D = (1 << i for i in range(7))
    sflags = {"cgnr": CGNR, "cgnr_damp": CGNR | PRE_D, "cgne_weight": CGNE | PRE_W}[solver]
    rng = np.random.default_rng(5)
    m, iters = 6, 12
    x = np.ascontiguousarray(rng.random((M, d)) - 0.5)
    NN = int(np.prod(N))
    y = np.ascontiguousarray(rng.random(M) - 0.5 + 1j * (rng.random(M) - 0.5))
    w = np.ascontiguousarray(0.5 + rng.random(M))
    k = np.stack(np.meshgrid(*[np.arange(-v // 2, v // 2) / v for v in N], indexing="ij"), -1)
    w_hat = np.ascontiguousarray((np.sqrt((k ** 2).sum(-1)) <= 0.5).astype(np.float64).ravel())   # disc mask, mri2d style
    nfft_flags = (abi.PRE_PHI_HUT | abi.PRE_PSI | abi.MALLOC_X | abi.MALLOC_F_HAT | abi.MALLOC_F | abi.FFTW_INIT
                  | abi.FFT_OUT_OF_PLACE)
    outs = []
    for so in (so_ref, so_b200):
        L = C.CDLL(so, mode=os.RTLD_LOCAL)
        fn = L.solver_driver_run
        fn.restype = C.c_int
        f_hat = np.zeros(NN, dtype=np.complex128)
        dots = np.zeros(iters)
        ia = lambda a: (C.c_int * len(a))(*a)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        rc = fn(C.c_int(d), ia(N), C.c_int(M), ia(n), C.c_int(m), C.c_uint(nfft_flags), C.c_uint(sflags),
                p(x), p(y), p(w), p(w_hat), C.c_int(iters), p(f_hat), p(dots), C.c_double(0.0), None)
        assert rc == 0
        outs.append((f_hat, dots))
    assert np.all(np.isfinite(outs[1][0]))
    assert rel_l2(outs[1][0], outs[0][0]) <= 1e-9
    assert np.allclose(outs[1][1], outs[0][1], rtol=1e-8, atol=0)


# ---- (10b) the device-resident solver behind the reference's solver_* API ------------------------------------
@pytest.mark.parametrize("precision", ["double", "float"])
@pytest.mark.parametrize("d,N,n,M,solver", [
    (2, [32, 32], [64, 64], 3000, "cgnr_damp"),
    (2, [32, 32], [64, 64], 3000, "cgnr_weight_damp"),
    (2, [32, 32], [64, 64], 800, "cgne_weight"),
    (2, [32, 32], [64, 64], 800, "cgne_damp"),
    (2, [24, 40], [48, 80], 2500, "steepest_weight_damp"),
    (2, [24, 40], [48, 80], 2500, "landweber_norms_damp"),
    (2, [24, 40], [48, 80], 2500, "landweber"),
    (3, [16, 16, 16], [32, 32, 32], 6000, "cgnr_damp"),
    (1, [128], [256], 400, "cgnr"),
])
def test_device_solver_vs_reference(d, N, n, M, solver, precision):
    """solver_*_complex / solverf_*_complex of libnfft3_b200.so (device-resident iteration: solver.cu behind
    solver_host.c) against the reference's kernel/solver/solver.c on the reference's own nfft.c, same driver
    (oracle/refbuild/solver_driver.c), 10 iterations: iterate, residual vector and residual norms."""
    ref_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
    so_ref, so_dev = os.path.join(ref_dir, "libsolver_ref.so"), os.path.join(ref_dir, "libsolver_dev_b200.so")
    if not (os.path.exists(so_ref) and os.path.exists(so_dev)):
        pytest.skip("oracle/_ref/libsolver_*.so not built (needs /root/reference at build time)")
    LANDWEBER, STEEPEST, CGNR, CGNE, NORMS, PRE_W, PRE_D = (1 << i for i in range(7))
    sflags = {"cgnr": CGNR, "cgnr_damp": CGNR | PRE_D, "cgnr_weight_damp": CGNR | PRE_W | PRE_D,
              "cgne_weight": CGNE | PRE_W, "cgne_damp": CGNE | PRE_D,
              "steepest_weight_damp": STEEPEST | PRE_W | PRE_D,
              "landweber_norms_damp": LANDWEBER | NORMS | PRE_D, "landweber": LANDWEBER}[solver]
    real, cplx, creal = ((np.float64, np.complex128, C.c_double) if precision == "double"
                         else (np.float32, np.complex64, C.c_float))
    rng = np.random.default_rng(5)
    m, iters = 6, 10
    x = np.ascontiguousarray((rng.random((M, d)) - 0.5).astype(real))
    if precision == "float":
        x = np.minimum(x, np.nextafter(np.float32(0.5), np.float32(0)))
    NN = int(np.prod(N))
    y = np.ascontiguousarray((rng.random(M) - 0.5 + 1j * (rng.random(M) - 0.5)).astype(cplx))
    w = np.ascontiguousarray((0.5 + rng.random(M)).astype(real))
    k = np.stack(np.meshgrid(*[np.arange(-v // 2, v // 2) / v for v in N], indexing="ij"), -1)
    w_hat = np.ascontiguousarray((np.sqrt((k ** 2).sum(-1)) <= 0.5).astype(real).ravel())
    nfft_flags = (abi.PRE_PHI_HUT | abi.PRE_PSI | abi.MALLOC_X | abi.MALLOC_F_HAT | abi.MALLOC_F | abi.FFTW_INIT
                  | abi.FFT_OUT_OF_PLACE)
    alpha = 0.5 / M   # Landweber step well inside 2 / ||A||^2
    outs = []
    for so in (so_ref, so_dev):
        L = C.CDLL(so, mode=os.RTLD_LOCAL)
        fn = L.solver_driver_run if precision == "double" else L.solver_driver_run_f
        fn.restype = C.c_int
        f_hat = np.zeros(NN, dtype=cplx)
        r = np.zeros(M, dtype=cplx)
        dots = np.zeros(iters, dtype=real)
        ia = lambda a: (C.c_int * len(a))(*a)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        rc = fn(C.c_int(d), ia(N), C.c_int(M), ia(n), C.c_int(m), C.c_uint(nfft_flags), C.c_uint(sflags),
                p(x), p(y), p(w), p(w_hat), C.c_int(iters), p(f_hat), p(dots), creal(alpha), p(r))
        assert rc == 0
        outs.append((f_hat, dots, r))
    tol = 1e-9 if precision == "double" else 2e-3
    assert np.all(np.isfinite(outs[1][0]))
    assert np.linalg.norm(outs[0][0]) > 0
    assert rel_l2(outs[1][0], outs[0][0]) <= tol
    assert rel_l2(outs[1][2], outs[0][2]) <= tol
    if not (sflags & LANDWEBER) or (sflags & NORMS):
        assert np.allclose(outs[1][1], outs[0][1], rtol=10 * tol, atol=0)


# ---- (11) the reference's kernel/mri and applications/fastsum, unmodified, on top of the engine -----------
def _apps_libs():
    ref_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
    so_ref, so_b200 = os.path.join(ref_dir, "libapps_ref.so"), os.path.join(ref_dir, "libapps_b200.so")
    if not (os.path.exists(so_ref) and os.path.exists(so_b200)):
        pytest.skip("oracle/_ref/libapps_*.so not built (needs /root/reference at build time)")
    return [C.CDLL(so, mode=os.RTLD_LOCAL) for so in (so_ref, so_b200)]


_ia = lambda a: (C.c_int * len(a))(*a)            # noqa: E731
_vp = lambda a: a.ctypes.data_as(C.c_void_p)      # noqa: E731


@pytest.mark.parametrize("variant", ["2d1d", "3d"])
@pytest.mark.parametrize("N0,N3,m,sigma,M", [(32, 12, 2, 1.25, 900), (64, 16, 4, 2.0, 5000)])
def test_reference_mri_inh_runs_on_the_engine(variant, N0, N3, m, sigma, M):
    """kernel/mri/mri.c:58-330 (mri_inh_2d1d_* loops 2 N3/2+1 NFFTs with host-side PHI / PHI_HUT scaling and
    swaps the plan's f / f_hat buffers with nfft_free + its own; mri_inh_3d_* wraps one 3-D NFFT), compiled from
    the reference tree and linked against libnfft3_b200.so, vs the same code on the reference's nfft.c."""
    libs = _apps_libs()
    rng = np.random.default_rng(31)
    NN = N0 * N0
    x = np.ascontiguousarray(rng.random((M, 2)) - 0.5)
    t = np.ascontiguousarray((rng.random(M) - 0.5) * (1 - 2 * m / N3) * 0.99)
    w = np.ascontiguousarray((rng.random(NN) - 0.5) * 0.5)
    fh = np.ascontiguousarray(rng.random(NN) + 1j * rng.random(NN))
    f = np.ascontiguousarray(rng.random(M) + 1j * rng.random(M))
    flags = abi.PRE_PHI_HUT | abi.PRE_PSI | abi.MALLOC_X | abi.MALLOC_F_HAT | abi.MALLOC_F | abi.FFTW_INIT
    n2 = int(np.ceil(N0 * sigma))
    outs = []
    for L in libs:
        fo, fho = np.zeros(M, complex), np.zeros(NN, complex)
        if variant == "2d1d":
            rc = L.apps_mri_inh_2d1d(_ia([N0, N0, N3]), M, _ia([n2, n2, N3]), m, C.c_double(sigma), C.c_uint(flags),
                                     _vp(x), _vp(t), _vp(w), _vp(fh), _vp(f), _vp(fo), _vp(fho))
        else:
            x3 = np.ascontiguousarray(np.concatenate([x, t[:, None]], 1))
            n3 = int(np.ceil(N3 * sigma / 2)) * 2
            rc = L.apps_mri_inh_3d(_ia([N0, N0, N3]), M, _ia([n2, n2, n3]), m, C.c_double(sigma), C.c_uint(flags),
                                   _vp(x3), _vp(w), _vp(fh), _vp(f), _vp(fo), _vp(fho))
        assert rc == 0
        outs.append((fo, fho))
    assert np.all(np.isfinite(outs[1][0])) and np.all(np.isfinite(outs[1][1]))
    assert rel_l2(outs[1][0], outs[0][0]) <= 1e-12
    assert rel_l2(outs[1][1], outs[0][1]) <= 1e-12


@pytest.mark.parametrize("d,Ns,Mt,nn,m,p,kern,fs_flags", [
    (2, 4000, 3000, 64, 4, 3, 1, 0),         # multiquadric, search-tree near field
    (2, 4000, 3000, 64, 6, 5, 0, 2),         # gaussian, NEARFIELD_BOXES
    (3, 3000, 3000, 32, 4, 3, 3, 0),         # inverse multiquadric, d = 3: the DMMA kernels
    (1, 2000, 2000, 128, 5, 4, 2, 0),        # 1/x (complex-valued regularisation), d = 1
])
def test_reference_fastsum_runs_on_the_engine(d, Ns, Mt, nn, m, p, kern, fs_flags):
    """applications/fastsum/fastsum.c (fastsum_init_guru 1062-1068: two NFFT plans without MALLOC_* whose
    x / f / f_hat are assigned after init 919-921, 981-983; fastsum_trafo 1170-1260: adjoint -> multiply by b ->
    trafo -> near field) compiled from the reference tree and linked against libnfft3_b200.so, vs the same code
    on the reference's nfft.c; also against the direct sum to the accuracy the reference itself reaches."""
    libs = _apps_libs()
    rng = np.random.default_rng(41)
    c = 1.0 / np.sqrt(float(nn)) if kern != 2 else 0.0
    eps_I, eps_B = p / nn, 1.0 / 16

    def ball(K):
        r = 0.25 - eps_B / 2
        v = (rng.random((4 * K + 64, d)) * 2 - 1) * r
        v = v[(v ** 2).sum(1) < r * r][:K]
        assert len(v) == K
        return np.ascontiguousarray(v)

    xs, ys = ball(Ns), ball(Mt)
    al = np.ascontiguousarray(rng.random(Ns) + 1j * rng.random(Ns))
    outs = []
    for i, L in enumerate(libs):
        L.apps_fastsum.argtypes = [C.c_int] * 7 + [C.c_double] * 3 + [C.c_uint] + [C.c_void_p] * 5
        fo = np.zeros(Mt, complex)
        fe = np.zeros(Mt, complex)
        rc = L.apps_fastsum(d, Ns, Mt, nn, m, p, kern, c, eps_I, eps_B, fs_flags, _vp(xs), _vp(al), _vp(ys),
                            _vp(fo), _vp(fe) if i == 0 else None)
        assert rc == 0
        outs.append((fo, fe))
    ref_f, exact = outs[0]
    got = outs[1][0]
    assert np.all(np.isfinite(got))
    assert rel_l2(got, ref_f) <= 1e-11
    # the approximation error of the fast sum itself is the same on both NFFT back ends
    scale = np.abs(al).sum()
    assert abs(np.abs(got - exact).max() - np.abs(ref_f - exact).max()) <= 1e-10 * scale


# ---- (12) BASELINE configs[1]: 2-D N=512^2, M=512^2 MRI-style trajectories, PRE_PSI, full size vs the oracle ---
def _mri_knots(kind, M, N):
    """applications/mri/mri2d/construct_knots_spiral.m:20-33 (one arm) and construct_knots_radial.m:18-31."""
    if kind == "spiral":
        A, w = 0.5, N / 64 * 50
        t = np.sqrt(np.arange(M) / M)
        x = np.stack([A * t * np.cos(2 * np.pi * w * t), A * t * np.sin(2 * np.pi * w * t)], 1)
    else:
        Z = int(np.sqrt(M))
        i, j = np.meshgrid(np.arange(Z), np.arange(Z), indexing="ij")
        r = (j.ravel() - Z / 2) / Z
        phi = np.pi * i.ravel() / Z
        x = np.stack([r * np.cos(phi), r * np.sin(phi)], 1)
    return np.ascontiguousarray(np.clip(x, -0.5, np.nextafter(0.5, 0.0)))


@pytest.mark.parametrize("precision", ["double", "float"])
@pytest.mark.parametrize("kind", ["spiral", "radial"])
def test_cfg2_mri_trajectories_vs_oracle(kind, precision):
    N, n, m, M = [512, 512], [1024, 1024], 6, 512 * 512
    spec = dict(d=2, N=N, n=n, m=m, M=M, seed=20260102,
                flags=BASE | abi.PRE_PSI)   # reconstruct_data_2d.c:47-55
    _, fh, f = make_case(spec, precision)
    x = _mri_knots(kind, M, 512)
    if precision == "float":
        x = np.minimum(x.astype(np.float32), np.nextafter(np.float32(0.5), np.float32(0)))
    o = oracle(precision)
    out_f, out_fh, _ = run_plan(spec, precision, x, fh, f)
    assert rel_l2(out_f, o.trafo(N, n, m, x, fh)) <= TOL[precision]
    assert rel_l2(out_fh, o.adjoint(N, n, m, x, f, True)) <= TOL[precision]


# ---- (13) BASELINE configs[3] grid: 3-D N=256^3, n=512^3 (one GPU's share of the node-sharded run) --------------
def test_cfg4_grid_subsample_and_adjointness():
    """N=256^3, n=512^3, m=6, M=4*10^6 nodes: a 2000-node subsample of trafo against the oracle's B step fed with
    the device grid (D and F are covered at sizes the oracle can run in full), and <A u, v> == <u, A^H v>."""
    rng = np.random.default_rng(777)
    N, n, m, M = [256] * 3, [512] * 3, 6, 4_000_000
    NN = 256 ** 3
    x = rng.random((M, 3)) - 0.5
    u = rng.random(NN) - 0.5 + 1j * (rng.random(NN) - 0.5)
    v = rng.random(M) - 0.5 + 1j * (rng.random(M) - 0.5)
    eng = cabi.Engine(N, n, m, M, precision="double")
    eng.set_nodes(x)
    Au = eng.trafo(u)
    g = eng.grid_to_host()
    AHv = eng.adjoint(v)
    eng.close()
    sel = rng.choice(M, 2000, replace=False)
    assert rel_l2(Au[sel], oracle("double").stage_B(N, n, m, x[sel], g)) <= 1e-12
    lhs, rhs = np.vdot(v, Au), np.vdot(AHv, u)
    assert abs(lhs - rhs) / (np.linalg.norm(v) * np.linalg.norm(Au)) <= 1e-12


# ---- (14) F for general lengths: mixed radix, direct prime stages, table DFT, four-step split of long axes ------
@pytest.mark.parametrize("precision", ["double", "float"])
@pytest.mark.parametrize("N,n,m,M", [
    ([33], [66], 4, 500),                 # 2 * 3 * 11: radix 11 as a direct stage
    ([60], [122], 5, 500),                # 2 * 61
    ([64], [134], 5, 500),                # 2 * 67: prime factor > 61 -> Bluestein, P = 512
    ([100], [254], 5, 800),               # 2 * 127 -> Bluestein, P = 512
    ([500], [1009], 6, 2000),             # prime length -> Bluestein, P = 2048 (odd n: sigma not an integer)
    ([1000], [2062], 6, 3000),            # 2 * 1031: Bluestein in fp32 (P = 8192), O(len^2) table DFT in fp64
    ([40, 36], [134, 79], 4, 2000),       # Bluestein on both axes of a 2-D grid (strided and contiguous lines)
    ([16, 20, 12], [32, 67, 24], 3, 1500),   # a Bluestein axis in the middle of a 3-D grid
    ([30, 42], [105, 90], 4, 2000),       # 3 * 5 * 7 and 2 * 3^2 * 5
    ([1000], [2100], 6, 3000),            # 2^2 * 3 * 5^2 * 7
    ([8192], [16384], 6, 5000),           # four-step, 128 x 128
    ([1 << 19], [1 << 20], 6, 20000),     # four-step, 1024 x 1024
    ([50000], [100000], 6, 10000),        # four-step with mixed-radix factors (250 x 400)
    ([4096, 16], [8192, 32], 4, 5000),    # split axis with faster axes behind it (fp64 only: 8192 fits in fp32)
    ([16, 4096], [32, 8192], 4, 5000),    # split last axis with slower axes in front
])
def test_fft_general_lengths_vs_oracle(N, n, m, M, precision):
    spec = dict(d=len(N), N=N, n=n, m=m, M=M, seed=97, flags=BASE)
    x, fh, f = make_case(spec, precision)
    o = oracle(precision)
    out_f, out_fh, _ = run_plan(spec, precision, x, fh, f)
    if precision == "float" and n[0] == 100000:
        # n*x is not exact in float when n is not a power of two: at n = 10^5 the fp32 reference (and the fp32
        # oracle) lose ~3e-4 in the window argument alone (ulp(5e4) = 0.004 grid cells).  The engine evaluates
        # the window in double from the float node, so it is compared with the fp64 oracle on the same inputs.
        o = oracle("double")
        x, fh, f = x.astype(np.float64), fh.astype(np.complex128), f.astype(np.complex128)
    assert rel_l2(out_f, o.trafo(N, n, m, x, fh)) <= TOL[precision]
    assert rel_l2(out_fh, o.adjoint(N, n, m, x, f, True)) <= TOL[precision]


# ---- (15) 2-D shared-memory tile kernels (tile2d.cu) against the generic kernels and the oracle ------------------
@pytest.mark.parametrize("precision", ["double", "float"])
@pytest.mark.parametrize("N,n,m,M,mode", [
    ([64, 64], [128, 128], 6, 20000, "uniform"),
    ([32, 48], [64, 96], 6, 5000, "uniform"),
    ([20, 30], [50, 72], 5, 3000, "uniform"),          # n not a multiple of the tile edge: partial tiles wrap
    ([8, 8], [16, 16], 6, 500, "uniform"),             # footprint (45) larger than the grid: rows / columns alias
    ([16, 10], [33, 21], 2, 700, "uniform"),           # odd n, small window
    ([64, 64], [160, 144], 7, 4000, "uniform"),        # 2m+2 = 16 taps
    ([128, 128], [256, 256], 6, 60000, "cluster"),     # thousands of nodes in one tile: many chunks per tile
    ([128, 128], [256, 256], 4, 30000, "edges"),       # nodes on cell boundaries and at +-0.5
])
def test_tile2d_vs_generic_and_oracle(N, n, m, M, mode, precision):
    rng = np.random.default_rng(91)
    o = oracle(precision)
    x = rng.random((M, 2)) - 0.5
    if mode == "cluster":
        x = 0.02 * rng.normal(size=(M, 2)) + 0.3
        x = (x + 0.5) % 1.0 - 0.5
    if mode == "edges":
        x = np.round(x * 64) / 64
        x[: M // 20] = -0.5
    x = np.minimum(x, 0.5 - 2 ** -20).astype(o.real)
    NN = int(np.prod(N))
    fh = (rng.random(NN) - 0.5 + 1j * (rng.random(NN) - 0.5)).astype(o.cplx)
    f = (rng.random(M) - 0.5 + 1j * (rng.random(M) - 0.5)).astype(o.cplx)
    outs = []
    for kernel in (0, 1):   # 0: tile2d, 1: generic warp-per-node
        eng = cabi.Engine(N, n, m, M, precision=precision)
        eng.set_option(cabi.OPT_B_KERNEL, kernel)
        eng.set_nodes(x)
        outs.append((eng.trafo(fh), eng.adjoint(f)))
        eng.close()
    tol = TOL[precision] if mode != "cluster" or precision == "double" else 1e-4   # fp32 sums of ~10^4 terms per cell
    assert rel_l2(outs[0][0], o.trafo(N, n, m, x, fh)) <= TOL[precision]
    assert rel_l2(outs[0][1], o.adjoint(N, n, m, x, f)) <= tol
    assert rel_l2(outs[0][0], outs[1][0]) <= (1e-13 if precision == "double" else 1e-5)
    assert rel_l2(outs[0][1], outs[1][1]) <= (1e-13 if precision == "double" else tol)
