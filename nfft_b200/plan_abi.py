"""ctypes mirror of the NFFT3 plan ABI (``nfft_plan`` / ``nfftf_plan``).

The struct layout restates ``include/nfft3.h:54-60`` (``MACRO_MV_PLAN``) and
``include/nfft3.h:109-161`` (``NFFT_DEFINE_API`` plan body) of the reference.  It is shared by

* the product's Python host mirror (:mod:`nfft_b200.plan`), which drives ``libnfft3_b200.so``
  -- the C host layer that exports the reference's own ``nfft_*``/``nfftf_*`` symbols on top
  of the CUDA C-ABI (``include/nfftcu.h``), and
* the parity tests, which load the *reference* build ``oracle/_ref/libnfft3_ref.so`` with the
  very same classes: both libraries speak the same ABI by construction, so a test can run one
  plan against both and compare.

Nothing here computes; it only describes memory.
"""
from __future__ import annotations

import ctypes as C

INT = C.c_ssize_t  # NFFT_INT == ptrdiff_t (include/nfft3.h:51)

# plan flags, include/nfft3.h:195-208
PRE_PHI_HUT = 1 << 0
FG_PSI = 1 << 1
PRE_LIN_PSI = 1 << 2
PRE_FG_PSI = 1 << 3
PRE_PSI = 1 << 4
PRE_FULL_PSI = 1 << 5
MALLOC_X = 1 << 6
MALLOC_F_HAT = 1 << 7
MALLOC_F = 1 << 8
FFT_OUT_OF_PLACE = 1 << 9
FFTW_INIT = 1 << 10
NFFT_SORT_NODES = 1 << 11
NFFT_OMP_BLOCKWISE_ADJOINT = 1 << 12
PRE_ONE_PSI = PRE_LIN_PSI | PRE_FG_PSI | PRE_PSI | PRE_FULL_PSI

# published FFTW3 flag values (kept for ABI fidelity of plan.fftw_flags)
FFTW_MEASURE = 0
FFTW_DESTROY_INPUT = 1 << 0
FFTW_ESTIMATE = 1 << 6


def _plan_struct(name: str, real):
    rp = C.POINTER(real)
    ip = C.POINTER(INT)
    fields = [
        # MACRO_MV_PLAN(C)
        ("N_total", INT),
        ("M_total", INT),
        ("f_hat", rp),          # C* viewed as interleaved reals
        ("f", rp),
        ("mv_trafo", C.c_void_p),
        ("mv_adjoint", C.c_void_p),
        # plan body
        ("d", INT),
        ("N", ip),
        ("sigma", rp),
        ("n", ip),
        ("n_total", INT),
        ("m", INT),
        ("b", rp),
        ("K", INT),
        ("flags", C.c_uint),
        ("fftw_flags", C.c_uint),
        ("x", rp),
        ("MEASURE_TIME_t", real * 3),
        ("my_fftw_plan1", C.c_void_p),
        ("my_fftw_plan2", C.c_void_p),
        ("c_phi_inv", C.POINTER(rp)),
        ("psi", rp),
        ("psi_index_g", ip),
        ("psi_index_f", ip),
        ("g", rp),
        ("g_hat", rp),
        ("g1", rp),
        ("g2", rp),
        ("spline_coeffs", rp),
        ("index_x", ip),
    ]
    return type(name, (C.Structure,), {"_fields_": fields})


NfftPlanD = _plan_struct("nfft_plan", C.c_double)
NfftPlanF = _plan_struct("nfftf_plan", C.c_float)

# byte offsets of every member on LP64 (x86-64 / aarch64 Linux); tests assert these against
# offsetof() in the C host layer and, when built, the reference header itself.
EXPECTED_SIZEOF = {"double": C.sizeof(NfftPlanD), "float": C.sizeof(NfftPlanF)}

API_VOID_FUNCS = (
    "trafo_direct", "adjoint_direct", "trafo", "trafo_1d", "trafo_2d", "trafo_3d",
    "adjoint", "adjoint_1d", "adjoint_2d", "adjoint_3d",
    "precompute_one_psi", "precompute_psi", "precompute_full_psi", "precompute_fg_psi",
    "precompute_lin_psi", "finalize",
)
API_ALL_FUNCS = API_VOID_FUNCS + (
    "init_1d", "init_2d", "init_3d", "init", "init_guru", "init_lin", "check",
    "malloc", "free", "die",
)


def bind_api(lib: C.CDLL, prefix: str):
    """Attach argtypes/restype for every plan-API symbol (include/nfft3.h:163-187, 69-82)."""
    struct = NfftPlanD if prefix == "nfft_" else NfftPlanF
    pp = C.POINTER(struct)
    ns = {}
    for fn in API_VOID_FUNCS:
        f = getattr(lib, prefix + fn)
        f.argtypes = [pp]
        f.restype = None
        ns[fn] = f
    intp = C.POINTER(C.c_int)
    sig = {
        "init_1d": [pp, C.c_int, C.c_int],
        "init_2d": [pp, C.c_int, C.c_int, C.c_int],
        "init_3d": [pp, C.c_int, C.c_int, C.c_int, C.c_int],
        "init": [pp, C.c_int, intp, C.c_int],
        "init_guru": [pp, C.c_int, intp, C.c_int, intp, C.c_int, C.c_uint, C.c_uint],
        "init_lin": [pp, C.c_int, intp, C.c_int, intp, C.c_int, C.c_int, C.c_uint, C.c_uint],
    }
    for fn, at in sig.items():
        f = getattr(lib, prefix + fn)
        f.argtypes = at
        f.restype = None
        ns[fn] = f
    f = getattr(lib, prefix + "check")
    f.argtypes = [pp]
    f.restype = C.c_char_p
    ns["check"] = f
    f = getattr(lib, prefix + "malloc")
    f.argtypes = [C.c_size_t]
    f.restype = C.c_void_p
    ns["malloc"] = f
    f = getattr(lib, prefix + "free")
    f.argtypes = [C.c_void_p]
    f.restype = None
    ns["free"] = f
    # split-phase extensions of libnfft3_b200.so (absent from a reference build)
    for fn in ("b200_trafo_begin", "b200_adjoint_begin", "b200_wait"):
        if hasattr(lib, prefix + fn):
            f = getattr(lib, prefix + fn)
            f.argtypes = [pp]
            f.restype = None
            ns[fn] = f
    return struct, ns
