"""Python host mirror of the NFFT3 plan API (same names, argument meaning, error behaviour).

``Plan`` is a thin object over a C ``nfft_plan``/``nfftf_plan`` living in ctypes memory; every
method forwards to the identically named C entry point (``include/nfft3.h:163-187`` of the
reference) of a library that speaks that ABI.  By default that library is the product,
``nfft_b200/lib/libnfft3_b200.so`` (C host layer -> ``libnfftcu.so`` -> sm_100a kernels).  There
is no CPU fallback: if the product library is missing, loading raises.

The parity tests construct the same class over the reference build (``oracle/_ref``) by passing
``api=``; that is the only way a non-product library gets in here.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

from . import plan_abi as abi
from .plan_abi import (  # noqa: F401  (re-exported, reference flag names)
    PRE_PHI_HUT, FG_PSI, PRE_LIN_PSI, PRE_FG_PSI, PRE_PSI, PRE_FULL_PSI, MALLOC_X, MALLOC_F_HAT,
    MALLOC_F, FFT_OUT_OF_PLACE, FFTW_INIT, NFFT_SORT_NODES, NFFT_OMP_BLOCKWISE_ADJOINT,
    PRE_ONE_PSI, FFTW_MEASURE, FFTW_DESTROY_INPUT, FFTW_ESTIMATE,
)

_LIBDIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib")
_PRODUCT = {}


class Api:
    """A loaded library + bound plan API for one precision ('double' -> nfft_, 'float' -> nfftf_)."""

    def __init__(self, lib: C.CDLL, precision: str):
        if precision not in ("double", "float"):
            raise ValueError("precision must be 'double' or 'float' (nfftl_ is not provided)")
        self.lib = lib
        self.precision = precision
        self.prefix = "nfft_" if precision == "double" else "nfftf_"
        self.struct, self.fn = abi.bind_api(lib, self.prefix)
        self.real = np.float64 if precision == "double" else np.float32
        self.cplx = np.complex128 if precision == "double" else np.complex64
        self.creal = C.c_double if precision == "double" else C.c_float


def product_api(precision: str = "double") -> Api:
    """The product library. Raises (never falls back) when the CUDA build is absent."""
    if precision not in _PRODUCT:
        cu = os.path.join(_LIBDIR, "libnfftcu.so")
        host = os.path.join(_LIBDIR, "libnfft3_b200.so")
        for p in (cu, host):
            if not os.path.exists(p):
                raise RuntimeError(
                    f"{p} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(nfft_b200 has no CPU fallback)")
        # local scope: libnfft3_b200.so finds libnfftcu.so through its rpath, and nothing else in
        # the process (e.g. a reference build loaded by the tests) may bind to these nfft_* names
        mode = os.RTLD_LOCAL | os.RTLD_NOW | getattr(os, "RTLD_DEEPBIND", 0)
        _PRODUCT[precision] = Api(C.CDLL(host, mode=mode), precision)
    return _PRODUCT[precision]


class Plan:
    """One NFFT plan. Mirrors ``nfft_plan`` usage in the reference (examples/nfft/simple_test.c.in)."""

    def __init__(self, api: Optional[Api] = None, precision: str = "double"):
        self.api = api if api is not None else product_api(precision)
        self.c = self.api.struct()
        self._live = False

    # -- initialisers (kernel/nfft/nfft.c:6048-6167) ------------------------------------------
    @classmethod
    def init_guru(cls, d: int, N: Sequence[int], M: int, n: Sequence[int], m: int, flags: int,
                  fftw_flags: int = FFTW_ESTIMATE | FFTW_DESTROY_INPUT, *,
                  api: Optional[Api] = None, precision: str = "double") -> "Plan":
        p = cls(api, precision)
        Na = (C.c_int * d)(*[int(v) for v in N])
        na = (C.c_int * d)(*[int(v) for v in n])
        p.api.fn["init_guru"](C.byref(p.c), d, Na, int(M), na, int(m), flags, fftw_flags)
        p._live = True
        return p

    @classmethod
    def init(cls, d: int, N: Sequence[int], M: int, *, api: Optional[Api] = None,
             precision: str = "double") -> "Plan":
        p = cls(api, precision)
        Na = (C.c_int * d)(*[int(v) for v in N])
        p.api.fn["init"](C.byref(p.c), d, Na, int(M))
        p._live = True
        return p

    @classmethod
    def init_nd(cls, N: Sequence[int], M: int, *, api: Optional[Api] = None,
                precision: str = "double") -> "Plan":
        """nfft_init_1d / _2d / _3d."""
        p = cls(api, precision)
        d = len(N)
        if d not in (1, 2, 3):
            raise ValueError("init_1d/2d/3d need 1..3 bandwidths")
        p.api.fn[f"init_{d}d"](C.byref(p.c), *[int(v) for v in N], int(M))
        p._live = True
        return p

    # -- plan members --------------------------------------------------------------------------
    @property
    def d(self) -> int:
        return int(self.c.d)

    @property
    def M_total(self) -> int:
        return int(self.c.M_total)

    @property
    def N_total(self) -> int:
        return int(self.c.N_total)

    @property
    def m(self) -> int:
        return int(self.c.m)

    @property
    def flags(self) -> int:
        return int(self.c.flags)

    @property
    def N(self):
        return [int(self.c.N[t]) for t in range(self.d)]

    @property
    def n(self):
        return [int(self.c.n[t]) for t in range(self.d)]

    def _view(self, ptr, count, dtype):
        if not ptr:
            raise ValueError("plan member is NULL (plan initialised without the MALLOC_ flag?)")
        buf = (C.c_char * (count * np.dtype(dtype).itemsize)).from_address(
            C.addressof(ptr.contents))
        return np.frombuffer(buf, dtype=dtype, count=count)

    @property
    def x(self) -> np.ndarray:
        return self._view(self.c.x, self.M_total * self.d, self.api.real).reshape(self.M_total, self.d)

    @property
    def f(self) -> np.ndarray:
        return self._view(self.c.f, self.M_total, self.api.cplx)

    @property
    def f_hat(self) -> np.ndarray:
        return self._view(self.c.f_hat, self.N_total, self.api.cplx)

    @property
    def index_x(self) -> np.ndarray:
        """[M,2] (key, original index) pairs, filled when NFFT_SORT_NODES (nfft.c:75-123)."""
        return self._view(self.c.index_x, 2 * self.M_total, np.int64).reshape(self.M_total, 2)

    @property
    def stage_times(self):
        return [float(self.c.MEASURE_TIME_t[i]) for i in range(3)]

    # -- operations ----------------------------------------------------------------------------
    def _call(self, name):
        if not self._live:
            raise RuntimeError("plan is finalized")
        self.api.fn[name](C.byref(self.c))

    def precompute_one_psi(self): self._call("precompute_one_psi")
    def precompute_psi(self): self._call("precompute_psi")
    def precompute_full_psi(self): self._call("precompute_full_psi")
    def precompute_lin_psi(self): self._call("precompute_lin_psi")
    def trafo(self): self._call("trafo")
    def adjoint(self): self._call("adjoint")
    def trafo_direct(self): self._call("trafo_direct")
    def adjoint_direct(self): self._call("adjoint_direct")

    # split-phase extensions (product library only): begin returns at once, wait blocks until the plan's work is done
    def trafo_begin(self): self._call("b200_trafo_begin")
    def adjoint_begin(self): self._call("b200_adjoint_begin")
    def wait(self): self._call("b200_wait")

    def trafo_nd(self):
        self._call(f"trafo_{self.d}d")

    def adjoint_nd(self):
        self._call(f"adjoint_{self.d}d")

    def check(self) -> Optional[str]:
        r = self.api.fn["check"](C.byref(self.c))
        return None if r is None else r.decode()

    def finalize(self):
        if self._live:
            self.api.fn["finalize"](C.byref(self.c))
            self._live = False

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.finalize()
