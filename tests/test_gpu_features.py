"""GPU parity of the SURVEY 8f rows beyond the plain Kaiser-Bessel transforms (run with -m gpu): Gaussian window,
batched "many vectors, one node set" transforms, device-side mri_inh scaling, fastsum far field.
Checker: the reference build under oracle/_ref (unmodified reference sources) or the CPU oracle."""
import ctypes as C
import os

import numpy as np
import pytest

import common
from common import make_case, rel_l2
from nfft_b200 import cabi, plan_abi as abi
from nfft_b200.plan import Api, Plan

pytestmark = pytest.mark.gpu
TOL = {"double": 1e-12, "float": 1e-5}
BASE = abi.PRE_PHI_HUT | abi.MALLOC_X | abi.MALLOC_F_HAT | abi.MALLOC_F | abi.FFTW_INIT | abi.FFT_OUT_OF_PLACE

_gauss = {}


def gauss_ref_api(precision):
    """the reference configured --with-window=gaussian (oracle/refbuild: -DORACLE_REF_GAUSSIAN)"""
    if precision not in _gauss:
        so = os.path.join(common.ROOT, "oracle", "_ref",
                          "libnfft3_ref_gauss.so" if precision == "double" else "libnfft3f_ref_gauss.so")
        mode = os.RTLD_LOCAL | os.RTLD_NOW | getattr(os, "RTLD_DEEPBIND", 0)
        _gauss[precision] = Api(C.CDLL(so, mode=mode), precision)
    return _gauss[precision]


def _pair(p, x, fh, f):
    p.x[:] = x
    if p.flags & abi.PRE_ONE_PSI:
        p.precompute_one_psi()
    p.f_hat[:] = fh
    p.trafo()
    out_f = p.f.copy()
    p.f[:] = f
    p.adjoint()
    out_fh = p.f_hat.copy()
    p.finalize()
    return out_f, out_fh


GAUSS_CASES = {
    "1d": dict(d=1, N=[256], n=[512], m=6, M=3000, seed=61, flags=BASE | abi.PRE_PSI),
    "1d_m13": dict(d=1, N=[64], n=[128], m=13, M=1000, seed=62, flags=BASE),
    "2d_fg": dict(d=2, N=[64, 48], n=[128, 96], m=6, M=20000, seed=63, flags=BASE | abi.PRE_FG_PSI | abi.FG_PSI),
    "2d_m8": dict(d=2, N=[32, 32], n=[80, 64], m=8, M=5000, seed=64, flags=BASE | abi.FG_PSI),
    "3d": dict(d=3, N=[32, 32, 32], n=[64, 64, 64], m=6, M=50000, seed=65,
               flags=BASE | abi.NFFT_SORT_NODES | abi.NFFT_OMP_BLOCKWISE_ADJOINT),
    "3d_m4_full": dict(d=3, N=[16, 24, 20], n=[32, 48, 40], m=4, M=4000, seed=66, flags=BASE | abi.PRE_FULL_PSI),
    "3d_m8": dict(d=3, N=[16, 16, 16], n=[40, 40, 40], m=8, M=3000, seed=67, flags=BASE),
}


@pytest.mark.skipif(not os.path.exists(os.path.join(common.ROOT, "oracle", "_ref", "libnfft3_ref_gauss.so")),
                    reason="oracle/_ref gaussian build not present")
@pytest.mark.parametrize("precision", ["double", "float"])
@pytest.mark.parametrize("case", sorted(GAUSS_CASES))
def test_gaussian_window_vs_reference(case, precision, monkeypatch):
    """NFFT_B200_WINDOW=gaussian: PHI / PHI_HUT of include/infft.h:154-173 on every kernel family (generic, 2-D tiles,
    tensor, pencil) against the reference built with the Gaussian window; FG_PSI / PRE_FG_PSI plans included."""
    monkeypatch.setenv("NFFT_B200_WINDOW", "gaussian")
    spec = GAUSS_CASES[case]
    x, fh, f = make_case(spec, precision)
    ref = Plan.init_guru(spec["d"], spec["N"], spec["M"], spec["n"], spec["m"], spec["flags"], api=gauss_ref_api(precision))
    ref_f, ref_fh = _pair(ref, x, fh, f)
    got = Plan.init_guru(spec["d"], spec["N"], spec["M"], spec["n"], spec["m"], spec["flags"], precision=precision)
    got_f, got_fh = _pair(got, x, fh, f)
    assert rel_l2(got_f, ref_f) <= TOL[precision]
    assert rel_l2(got_fh, ref_fh) <= TOL[precision]
    # and the window really is a different one: far from the Kaiser-Bessel result at the Gaussian's accuracy level
    monkeypatch.delenv("NFFT_B200_WINDOW")
    kb = Plan.init_guru(spec["d"], spec["N"], spec["M"], spec["n"], spec["m"], spec["flags"], precision=precision)
    kb_f, _ = _pair(kb, x, fh, f)
    assert rel_l2(kb_f, ref_f) > 1e-9 or precision == "float"
